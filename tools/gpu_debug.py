"""Not a test: verbose first-contact script for the GPU box (prints mismatch statistics)."""
import sys, time, itertools
import numpy as np
sys.path.insert(0, ".")
import oracle as O
import mgm_b200
from tests.conftest import synth_volume, synth_weights, synth_pair

ctx = mgm_b200.Context()
def cmp(name, a, b):
    eq = np.array_equal(a, b, equal_nan=True)
    if eq:
        print("  ok   ", name); return True
    fin = np.isfinite(a) & np.isfinite(b)
    d = np.abs(a[fin] - b[fin])
    print("  FAIL ", name, "mismatch frac %.5f" % np.mean(~((a == b) | (np.isnan(a) & np.isnan(b)))),
          "max abs diff %.4g" % (d.max() if d.size else -1), "nonfinite pattern equal:", np.array_equal(np.isfinite(a), np.isfinite(b)))
    return False

mode = sys.argv[1] if len(sys.argv) > 1 else "all"
u, v = synth_pair(37, 23, 12, seed=3, nch=3)
print("weights"); cmp("w", ctx.compute_mgm_weights(u, 4.0, 12.0), O.orc_weights(u, 4.0, 12.0))
for dist in ["ad", "sd", "census", "ncc", "btad", "btsd"]:
    for win in ([3, 5] if dist in ("census", "ncc") else [3]):
        for uu, vv in [(u, v), (u[:1], v[:1])]:
            a = ctx.allocate_and_fill_sgm_costvolume(uu, vv, -11, 2, "none", dist, np.inf, win)
            cmp("cc %s win%d nch%d" % (dist, win, uu.shape[0]), a, O.orc_costvolume(uu, vv, -11, 2, "none", dist, np.inf, win))
a = ctx.allocate_and_fill_sgm_costvolume(u, v, -11, 2, "sobelx", "ad", 20.0, 3)
cmp("cc sobelx trunc20", a, O.orc_costvolume(u, v, -11, 2, "sobelx", "ad", 20.0, 3))
if mode == "cc": sys.exit(0)

shapes = [(23, 17, 9), (67, 41, 19)]
bad = 0
for (nx, ny, L) in shapes:
    for real in [0, 1]:
        cc = synth_volume(nx, ny, L, seed=nx, real=bool(real))
        for rows in [0, 5]:
            ctx.set_rows_per_band(rows)
            for wt in [0, 1]:
                w = synth_weights(nx, ny, seed=nx) if wt else None
                for felz in [0, 1]:
                    for K in [1, 2, 3, 4]:
                        for NDIR in ([8] if rows == 0 else [4, 8]):
                            P1, P2 = (8, 32) if not felz else (2, 20000)
                            t0 = time.time()
                            r = ctx.mgm(cc, w, -(L - 1), P1, P2, NDIR, K, felz, 1)
                            o = O.orc_mgm(cc, w, -(L - 1), P1, P2, NDIR, K, felz, 1)
                            name = "mgm %dx%dx%d real%d rows%d w%d felz%d K%d NDIR%d" % (nx, ny, L, real, rows, wt, felz, K, NDIR)
                            ok = np.array_equal(r["out"], o["out"], equal_nan=True) and np.array_equal(r["S"], o["S"], equal_nan=True) and np.array_equal(r["outcost"], o["outcost"])
                            if not ok:
                                bad += 1
                                print(name, "info", ctx.last_launch_info())
                                cmp("   out", r["out"], o["out"]); cmp("   S", r["S"], o["S"]); cmp("   outcost", r["outcost"], o["outcost"])
                                if bad > 12: print("too many failures"); sys.exit(1)
print("mgm done, bad =", bad)
ctx.set_rows_per_band(0)
cc = synth_volume(67, 41, 19, seed=5, real=True)
r = ctx.mgm(cc, None, -18, 8, 32, 8, 2, 0, 1)
for m in mgm_b200.REFINEMENTS:
    a = ctx.subpixel_refinement_sgm(r["S"], -18, r["out"], r["outcost"], m)
    b = O.orc_refine(r["S"], -18, r["out"], r["outcost"], m)
    cmp("refine " + m + " out", a[0], b[0]); cmp("refine " + m + " cost", a[1], b[1])
u, v = synth_pair(97, 55, 24, seed=1, nch=1)
for kw in [dict(distance="census", census_ncc_win=5, NDIR=8, MGM=2, refinement="vfit"),
           dict(distance="census", census_ncc_win=3, NDIR=8, MGM=3, use_felzenszwalb_potentials=1, P1=2, P2=20000, refinement="vfit"),
           dict(distance="ad", NDIR=4, MGM=4, aP=4.0, aThresh=6.0, refinement="cubic")]:
    a = ctx.stereo(u, v, dmin=-23, dmax=0, **kw)
    kk = dict(kw); kk["win"] = kk.pop("census_ncc_win", 3); kk["K"] = kk.pop("MGM"); kk["felz"] = kk.pop("use_felzenszwalb_potentials", 0)
    b = O.orc_pipeline(u, v, -23, 0, **kk)
    cmp("stereo out " + str(kw), a[0], b["out"]); cmp("stereo cost", a[1], b["outcost"])
print("DONE")
