"""Not a test: whole-workload aggregation time vs rows per band (several CTAs per SM when the bands are small)."""
import os, sys
sys.path.insert(0, ".")
pairs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(43, 35), (20, 16), (13, 10), (9, 7)]
from tools.gpu_micro import run
for (ta, td) in pairs:
    os.environ["MGMB200_ROWS_AXIS"] = str(ta)
    os.environ["MGMB200_ROWS_DIAG"] = str(td)
    for mask, name in [(0xFF, "all"), (0x0F, "axis"), (0xF0, "diag")]:
        ms, info = run(2048, 1536, 256, 3, 1, mask, reps=2)
        print("%-5s rows=%2d/%2d thr=%d smem=%d: %.2f ms" % (name, info["rows_axis"], info["rows_diag"], info["threads_per_cta"], info["smem_bytes"], ms), flush=True)
