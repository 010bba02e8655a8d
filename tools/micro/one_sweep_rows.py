"""Not a test: one sweep (or two) alone on the GPU -- the per-GPU share of the sweep-sharded layout -- vs rows per band."""
import sys
sys.path.insert(0, ".")
from tools.gpu_micro import run
for mask, name in [(0x01, "sweep 0"), (0x04, "sweep 2"), (0x10, "sweep 4"), (0x03, "sweeps 0,1"), (0x11, "sweeps 0,4")]:
    for rows in (56, 40, 28, 20, 14, 10, 7):
        ms, info = run(2048, 1536, 256, 3, 1, mask, rows=rows, reps=2)
        print("%-10s rows=%2d/%2d thr=%3d: %.2f ms" % (name, info["rows_axis"], info["rows_diag"], info["threads_per_cta"], ms), flush=True)
