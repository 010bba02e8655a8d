"""Not a test: incremental cost of chained bands (H = n*T rows, sweep 0)."""
import sys
sys.path.insert(0, ".")
from tests.gpu_micro import run
for T in [43, 16]:
    base = None
    for n in [1, 2, 3, 4, 8]:
        ms, info = run(2048, n * T, 256, 3, 1, 0x01, rows=T, reps=2)
        if base is None: base = ms
        print("T=%d bands=%d: %.3f ms  (+%.3f ms per extra band = %.1f steps of %.2f us)" % (T, n, ms, (ms - base) / max(n - 1, 1), (ms - base) / max(n - 1, 1) / (base / (2048 + T - 1)), base / (2048 + T - 1) * 1e3), flush=True)
