import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def ctx():
    import mgm_b200
    c = mgm_b200.Context()
    yield c
    c.close()


def synth_pair(nx, ny, L, seed=0, nch=1):
    """SURVEY.md 8d synthetic stereo pair: blurred integer noise, sinusoidal disparity field."""
    rng = np.random.default_rng(seed)
    base = rng.random((nch, ny, nx + L))
    for k in (2, 4, 8):   # 3 octaves of box blur
        ker = np.ones(k) / k
        base = base + np.apply_along_axis(lambda r: np.convolve(r, ker, mode="same"), 2, base)
    base = np.floor(255 * (base - base.min()) / (base.max() - base.min() + 1e-9))
    v = base[:, :, :nx].astype(np.float32)
    yy, xx = np.mgrid[0:ny, 0:nx]
    d = -np.round(L / 4 + (L / 8) * np.sin(0.01 * xx) * np.cos(0.013 * yy)).astype(int)
    xs = np.clip(xx + d, 0, nx - 1)
    u = v[:, yy, xs] + rng.integers(-2, 3, (nch, ny, nx))
    return np.ascontiguousarray(u, np.float32), np.ascontiguousarray(v, np.float32)


def synth_volume(nx, ny, L, seed=0, real=False, inf_border=True):
    """Aggregator-only input (mirrors matlab/mgm_o): integer costs in [0,64), optional INF wedge."""
    rng = np.random.default_rng(seed)
    cc = rng.integers(0, 64, (ny, nx, L)).astype(np.float32)
    if real:
        cc = (cc + 3 * rng.random(cc.shape)).astype(np.float32)
    if inf_border:   # what "-r -(L-1) -R 0" produces near the left border: match outside the image
        for x in range(min(nx, L - 1)):
            cc[:, x, : L - 1 - x] = np.inf
    return cc


def synth_weights(nx, ny, seed=0, p=0.3, val=4.0):
    rng = np.random.default_rng(seed + 1000)
    return np.where(rng.random((8, ny, nx)) < p, val, 1.0).astype(np.float32)
