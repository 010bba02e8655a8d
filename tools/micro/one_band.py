import sys
sys.path.insert(0, ".")
from tools.gpu_micro import run
print(run(2048, 56, 256, 3, 1, 0x01, rows=56, reps=1))
