import sys
sys.path.insert(0, ".")
from tools.gpu_micro import run
for K in (3,):
    ms, info = run(2048, 56, 256, K, 0, 0x01, rows=56, reps=2)
    print("sgm K%d 1 band 56 rows: %.3f ms -> %.2f us/step" % (K, ms, ms * 1e3 / (2048 + 55)))
