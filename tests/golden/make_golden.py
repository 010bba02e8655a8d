"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libmgmref.so, compiled from
/root/reference by oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py

Every fixture stores the inputs and the reference's own outputs, so that the oracle port and the CUDA path
can be checked where /root/reference is not mounted (the GPU box).  Inputs are crops of the image pairs the
reference ships (data/fountain23-im{L,R}.png, data/im{L,R}.png = tsukuba, rectified_{ref,sec}.tif) and
seeded synthetic volumes.  Kept small: the whole directory is < 1.5 MB.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

DATA = "/root/reference/matlab/data"


def load_pair(name):
    from PIL import Image
    if name == "fountain":
        a, b = "fountain23-imL.png", "fountain23-imR.png"
    elif name == "tsukuba":
        a, b = "imL.png", "imR.png"
    else:
        a, b = "rectified_ref.tif", "rectified_sec.tif"
    out = []
    for f in (a, b):
        im = np.asarray(Image.open(os.path.join(DATA, f))).astype(np.float32)
        if im.ndim == 3:
            im = np.transpose(im[:, :, :3], (2, 0, 1))
        else:
            im = im[None]
        im = np.where(np.isfinite(im), im, 0).astype(np.float32)   # remove_nonfinite_values_Img, mgm.cc:335
        out.append(np.ascontiguousarray(im))
    return out


def crop(pair, x0, y0, w, h):
    return [np.ascontiguousarray(p[:, y0:y0 + h, x0:x0 + w]) for p in pair]


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print("%-40s %7.1f KB" % (name, os.path.getsize(path) / 1024))


def pipeline_fixture(name, u, v, **p):
    """whole hot path (mgm.cc:356-385) on an image pair"""
    r = O.ref_pipeline(u, v, **p)
    # also keep the intermediate cost volume digest and the weights for finer-grained checks
    cc = O.ref_costvolume(u, v, p["dmin"], p["dmax"], p.get("prefilter", "none"), p.get("distance", "ad"),
                          p.get("truncDist", np.inf), p.get("win", 3))
    w = O.ref_weights(u, p.get("aP", 1.0), p.get("aThresh", 5.0))
    save(name, kind="pipeline", u=u, v=v, params=np.array(repr(p)), out=r["out"], outcost=r["outcost"],
         cc_sum=np.array([np.nansum(np.where(np.isfinite(cc), cc, 0), dtype=np.float64)]),
         cc_ninf=np.array([np.isinf(cc).sum()]), w_not_one=np.array([(w != 1).sum()]))


def volume_fixture(name, cc, w, **p):
    """mgm() alone on a cost volume (the matlab/mgm_o entry)"""
    r = O.ref_mgm(cc, w, p["dmin"], p["P1"], p["P2"], p["NDIR"], p["K"], p.get("felz", 0), p.get("fix", 1))
    S = r["S"]
    save(name, kind="volume", cc=cc.astype(np.float16) if p.get("half_ok") else cc,
         w=(w if w is not None else np.zeros(0, np.float32)), params=np.array(repr(p)), out=r["out"],
         outcost=r["outcost"], S_row=S[S.shape[0] // 2].copy(),
         S_sum=np.array([np.sum(np.where(np.isfinite(S), S, 0), dtype=np.float64)]))


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle)"
    f = load_pair("fountain")
    t = load_pair("tsukuba")
    s = load_pair("rect")
    # config 1 of BASELINE.json (fountain23, -t ad -O 4 -r -120 -R 30, TSGM=2) on a crop, range scaled to the crop
    u, v = crop(f, 300, 200, 96, 64)
    pipeline_fixture("fountain_ad_O4_tsgm2", u, v, dmin=-40, dmax=10, P1=8.0, P2=32.0, NDIR=4, K=2, felz=0,
                     distance="ad", refinement="none")
    # Makefile:17 (make test line 1): census 3x3, trunc-linear, TSGM=3, vfit, -O 8, P1 2 P2 20000
    pipeline_fixture("fountain_census_trunc_tsgm3_vfit", u, v, dmin=-40, dmax=10, P1=2.0, P2=20000.0, NDIR=8, K=3,
                     felz=1, distance="census", win=3, refinement="vfit")
    # Makefile:18 (line 2): "-p sobel_x" is an unknown name -> none; AD with truncDist 63, P1 4
    pipeline_fixture("fountain_ad_trunc63_tsgm3_vfit", u, v, dmin=-40, dmax=10, P1=4.0, P2=20000.0, NDIR=8, K=3,
                     felz=1, distance="ad", prefilter="sobel_x", truncDist=63.0, refinement="vfit")
    # tsukuba gray-ish crop, CLI defaults (TSGM=4), census 5x5, weights on
    u, v = crop(t, 120, 90, 80, 56)
    pipeline_fixture("tsukuba_census5_tsgm4_weights", u, v, dmin=-16, dmax=0, P1=8.0, P2=32.0, NDIR=8, K=4, felz=0,
                     aP=4.0, aThresh=12.0, distance="census", win=5, refinement="parabola")
    pipeline_fixture("tsukuba_ncc_tsgm1_cubic", u[:1], v[:1], dmin=-16, dmax=0, P1=8.0, P2=32.0, NDIR=2, K=1,
                     felz=0, distance="ncc", win=3, refinement="cubic")
    # the float tif pair (contains NaNs in the source, cleaned like the CLI does), BT distance, sobelx prefilter
    u, v = crop(s, 60, 60, 72, 48)
    pipeline_fixture("rectified_btad_sobelx_tsgm2", u, v, dmin=-12, dmax=12, P1=8.0, P2=32.0, NDIR=8, K=2, felz=0,
                     distance="btad", prefilter="sobelx", refinement="parabolaOCV")
    pipeline_fixture("rectified_sd_trunc_weights_tsgm2", u, v, dmin=-12, dmax=12, P1=3.0, P2=40.0, NDIR=8, K=2,
                     felz=1, aP=0.5, aThresh=30.0, distance="sd", truncDist=400.0, refinement="vfit")
    # synthetic volumes for the aggregator alone
    rng = np.random.default_rng(2024)
    for name, K, felz, wt, (P1, P2) in [("vol_sgm_k2", 2, 0, 0, (8.0, 32.0)), ("vol_sgm_k3_w", 3, 0, 1, (8.0, 32.0)),
                                        ("vol_trunc_k3", 3, 1, 0, (2.0, 20000.0)),
                                        ("vol_trunc_k4_w", 4, 1, 1, (1.5, 11.0)), ("vol_trunc_k2", 2, 1, 0, (2.0, 9.0))]:
        nx, ny, L = 45, 29, 21
        cc = rng.integers(0, 64, (ny, nx, L)).astype(np.float32)
        for x in range(min(nx, L - 1)):
            cc[:, x, : L - 1 - x] = np.inf
        w = np.where(rng.random((8, ny, nx)) < 0.3, 4.0, 1.0).astype(np.float32) if wt else None
        volume_fixture(name, cc, w, dmin=-(L - 1), P1=P1, P2=P2, NDIR=8, K=K, felz=felz, fix=1, half_ok=True)


if __name__ == "__main__":
    main()
