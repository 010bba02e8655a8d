"""Not a test: incremental cost of chained bands (H = n*T rows, sweep 0; W = n*T for the diagonal sweep 4)."""
import sys
sys.path.insert(0, ".")
from tools.gpu_micro import run
T = int(sys.argv[1]) if len(sys.argv) > 1 else 56
for mask, name in [(0x01, "sweep0"), (0x10, "sweep4")]:
    base = None
    for n in [1, 2, 3, 4, 8]:
        ms, info = run(2048, n * T, 256, 3, 1, mask, rows=T, reps=2)
        if base is None:
            base = ms
        us_step = base / (2048 + T - 1) * 1e3 if mask == 1 else None
        extra = (ms - base) / max(n - 1, 1)
        print("%s T=%d bands=%d: %.3f ms  (+%.3f ms per extra band%s)" % (
            name, info["rows_axis"], n, ms, extra, " = %.1f steps of %.2f us" % (extra * 1e3 / us_step, us_step) if us_step else ""), flush=True)
