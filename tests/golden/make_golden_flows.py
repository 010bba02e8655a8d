"""Generates tests/golden/flow_*.npz from the UNMODIFIED reference BINARY (oracle/_ref/mgm, built from
/root/reference/mgm.cc by oracle/Makefile): whole command-line flows on small seeded image pairs written as 8-bit
PNM files.  Run in the build container:  python tests/golden/make_golden_flows.py

  flow_lr      both directions, MEDIAN, left-right tests, -l map, back-projection            (mgm.cc:372-443)
  flow_ranges  -m/-M range images and TSGM_ITER iterations, one direction (TESTLRRL=0)        (mgm.cc:342-395)

Every fixture stores the inputs (the float images the 8-bit files decode to, the range images as the CLI repairs
them, mgm.cc:346-352) and the reference's output files.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.conftest import synth_pair  # noqa: E402

REF_MGM = os.path.join(ROOT, "oracle", "_ref", "mgm")


def pnm(path, a):
    a = a.astype(np.uint8)
    if a.shape[0] == 1:
        open(path, "wb").write(b"P5\n%d %d\n255\n" % (a.shape[2], a.shape[1]) + a[0].tobytes())
    else:
        open(path, "wb").write(b"P6\n%d %d\n255\n" % (a.shape[2], a.shape[1]) + np.transpose(a, (1, 2, 0)).tobytes())


def images(nx, ny, L, seed, nch):
    u, v = synth_pair(nx, ny, L, seed=seed, nch=nch)
    u = np.round(np.clip(u, 0, 255)).reshape(nch, ny, nx).astype(np.float32)
    v = np.round(np.clip(v, 0, 255)).reshape(nch, ny, nx).astype(np.float32)
    return u, v


def run_cli(tmp, u, v, args, env, outputs):
    pnm(os.path.join(tmp, "u.pnm"), u)
    pnm(os.path.join(tmp, "v.pnm"), v)
    names = [os.path.join(tmp, k + ".npy") for k in outputs]
    r = subprocess.run([REF_MGM] + args + [os.path.join(tmp, "u.pnm"), os.path.join(tmp, "v.pnm")] + names,
                       env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    return [np.load(n) for n in names]


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print("%-40s %7.1f KB" % (name, os.path.getsize(path) / 1024))


def flow_lr(name, nx, ny, L, seed, nch, p):
    u, v = images(nx, ny, L, seed, nch)
    with tempfile.TemporaryDirectory() as tmp:
        nolr = os.path.join(tmp, "nolr.npy")
        args = ["-r", str(p["dmin"]), "-R", str(p["dmax"]), "-t", p["distance"], "-O", str(p["NDIR"]), "-P1", repr(p["P1"]),
                "-P2", repr(p["P2"]), "-s", p["refinement"], "-aP2", repr(p["aP"]), "-aThresh", repr(p["aThresh"]), "-l", nolr]
        env = dict(TSGM=str(p["K"]), MEDIAN=str(p["median"]), TESTLRRL="1", TESTLRRL_TAU=repr(p["tau"]),
                   USE_TRUNCATED_LINEAR_POTENTIALS=str(p["felz"]), CENSUS_NCC_WIN=str(p["win"]))
        disp, cost, back = run_cli(tmp, u, v, args, env, ["disp", "cost", "back"])
        nolr = np.load(nolr)
    back = np.moveaxis(back, -1, 0) if back.ndim == 3 else back[None]
    save(name, kind="flow_lr", u=u, v=v, params=np.array(repr(p)), disp=np.squeeze(disp), cost=np.squeeze(cost),
         back=np.ascontiguousarray(back), nolr=np.squeeze(nolr))


def flow_ranges(name, nx, ny, L, seed, p, ragged):
    u, v = images(nx, ny, L, seed, 1)
    with tempfile.TemporaryDirectory() as tmp:
        args = ["-r", str(p["dmin"]), "-R", str(p["dmax"]), "-t", p["distance"], "-O", str(p["NDIR"]), "-P1", repr(p["P1"]),
                "-P2", repr(p["P2"]), "-s", p["refinement"], "-truncDist", repr(p["truncDist"])]
        if ragged:
            rng = np.random.default_rng(seed)
            lo = (p["dmin"] + 9 + rng.integers(-6, 3, (ny, nx))).astype(np.float32)
            hi = (lo + rng.integers(0, 14, (ny, nx))).astype(np.float32)
            np.save(os.path.join(tmp, "dmin.npy"), lo)
            np.save(os.path.join(tmp, "dmax.npy"), hi)
            args += ["-m", os.path.join(tmp, "dmin.npy"), "-M", os.path.join(tmp, "dmax.npy")]
            hi = np.where(hi < lo + 1, np.ceil(lo + 1), hi).astype(np.float32)   # mgm.cc:349-352
        else:
            lo = np.full((ny, nx), p["dmin"], np.float32)
            hi = np.full((ny, nx), p["dmax"], np.float32)
        env = dict(TSGM=str(p["K"]), TSGM_ITER=str(p["iters"]), TESTLRRL="0", USE_TRUNCATED_LINEAR_POTENTIALS=str(p["felz"]),
                   CENSUS_NCC_WIN=str(p["win"]))
        disp, cost = run_cli(tmp, u, v, args, env, ["disp", "cost"])
    save(name, kind="flow_ranges", u=u, v=v, lo=lo, hi=hi, params=np.array(repr(p)), disp=np.squeeze(disp),
         cost=np.squeeze(cost))


def main():
    assert os.path.exists(REF_MGM), "build oracle/_ref first (make -C oracle)"
    base = dict(P1=8.0, P2=32.0, NDIR=8, K=2, felz=0, distance="ad", win=3, refinement="none", aP=1.0, aThresh=5.0)
    flow_lr("flow_lr_census_trunc_tsgm3_median_vfit", 96, 56, 20, 21, 1,
            dict(base, dmin=-19, dmax=3, P1=2.0, P2=20000.0, K=3, felz=1, distance="census", refinement="vfit", median=1, tau=1.0))
    flow_lr("flow_lr_sd_rgb_weights_tsgm4_cubic", 80, 48, 16, 22, 3,
            dict(base, dmin=-15, dmax=2, K=4, distance="sd", refinement="cubic", aP=4.0, aThresh=9.0, median=2, tau=2.0))
    flow_ranges("flow_ranges_ad_sgm_tsgm2_iter2", 96, 56, 20, 23,
                dict(base, dmin=-23, dmax=4, truncDist=40.0, refinement="vfit", iters=2), ragged=True)
    flow_ranges("flow_ranges_ad_trunc_tsgm3_iter3", 96, 56, 20, 24,
                dict(base, dmin=-23, dmax=4, P1=2.0, P2=20000.0, K=3, felz=1, truncDist=40.0, refinement="parabola", iters=3),
                ragged=True)
    flow_ranges("flow_ranges_census_uniform_tsgm2_iter2", 96, 56, 20, 25,
                dict(base, dmin=-21, dmax=2, P1=2.0, P2=20000.0, felz=1, distance="census", truncDist=20.0, refinement="vfit",
                     iters=2), ragged=False)


if __name__ == "__main__":
    main()
