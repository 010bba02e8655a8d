"""Not a test: the automatic rows-per-band choice for launches with few sweeps."""
import sys
sys.path.insert(0, ".")
from tools.gpu_micro import run
for mask, name in [(0x01, "sweep 0"), (0x10, "sweep 4"), (0x11, "sweeps 0,4"), (0x33, "sweeps 0,1,4,5"), (0xFF, "all")]:
    ms, info = run(2048, 1536, 256, 3, 1, mask, rows=0, reps=2)
    print("%-16s rows=%2d/%2d thr=%3d: %.2f ms" % (name, info["rows_axis"], info["rows_diag"], info["threads_per_cta"], ms), flush=True)
ms, info = run(1242, 375, 192, 4, 0, 0xFF, rows=0, reps=2); print("kitti sgm K4 all  rows=%d/%d: %.2f ms" % (info["rows_axis"], info["rows_diag"], ms))
ms, info = run(640, 480, 64, 2, 0, 0xFF, rows=0, reps=2); print("vga sgm K2 all    rows=%d/%d: %.2f ms" % (info["rows_axis"], info["rows_diag"], ms))
ms, info = run(640, 480, 64, 2, 0, 0xFF, rows=56, reps=2); print("vga sgm K2 all    rows=%d/%d: %.2f ms" % (info["rows_axis"], info["rows_diag"], ms))
