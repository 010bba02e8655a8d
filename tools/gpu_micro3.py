"""Not a test: full-size sweeps subsets vs rows per band."""
import sys
sys.path.insert(0, ".")
from tools.gpu_micro import run
for mask, name in [(0x0F, "axis 0-3"), (0xF0, "diag 4-7"), (0x01, "sweep0"), (0x10, "sweep4"), (0xFF, "all")]:
    for T in [43, 32, 24, 16, 12, 8]:
        ms, info = run(2048, 1536, 256, 3, 1, mask, rows=T, reps=2)
        print("%-9s rows=%2d/%2d: %.2f ms" % (name, info["rows_axis"], info["rows_diag"], ms), flush=True)
