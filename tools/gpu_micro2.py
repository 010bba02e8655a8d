"""Not a test: per-step latency of one band as a function of the rows per band."""
import sys
sys.path.insert(0, ".")
import torch
import mgm_b200
from tools.gpu_micro import run, ctx
for felz, K in [(1, 3), (0, 3)]:
    for T in ([int(a) for a in sys.argv[1:]] or [43, 32, 24, 16, 8, 4]):
        for mask, sig in [(0x01, 1), (0x10, 2)]:
            if mask == 0x10 and T > 35: continue
            W, H = 2048, T
            ms, info = run(W, H, 256, K, felz, mask, rows=T, reps=2)
            steps = W + sig * (T - 1)
            print("felz=%d K=%d T=%2d sweep=%s: %.3f ms  %.2f us/step" % (felz, K, T, "axis" if mask == 1 else "diag", ms, ms * 1e3 / steps), flush=True)
