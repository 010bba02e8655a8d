"""Not a test: time the two directions of the CLI flow separately and together."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from mgm_b200 import Context
wl = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
W, H, L = wl["W"], wl["H"], wl["L"]
u, v = bench.synth_pair(W, H, L, 0)
ctx = Context(0)
kw = dict(P1=wl["P1"], P2=wl["P2"], NDIR=wl["NDIR"], MGM=wl["K"], use_felzenszwalb_potentials=wl["felz"], distance="census",
          census_ncc_win=wl["win"], refinement=wl["refine"])
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n
print("L->R stereo       %.2f ms" % t(lambda: ctx.stereo(u, v, dmin=-(L - 1), dmax=0, **kw)))
print("R->L stereo       %.2f ms" % t(lambda: ctx.stereo(v, u, dmin=0, dmax=L - 1, **kw)))
print("stereo_lr         %.2f ms" % t(lambda: ctx.stereo_lr(u, v, dmin=-(L - 1), dmax=0, testlrrl=1, median=1, **kw)))
print("stereo_lr no med  %.2f ms" % t(lambda: ctx.stereo_lr(u, v, dmin=-(L - 1), dmax=0, testlrrl=1, median=0, **kw)))
print("stereo_lr no lr   %.2f ms" % t(lambda: ctx.stereo_lr(u, v, dmin=-(L - 1), dmax=0, testlrrl=0, median=1, **kw)))
