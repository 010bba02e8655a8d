(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
timeout 300 python tools/gpu_micro4.py 56 2>&1 | grep "bands=1:"
timeout 300 python tools/micro/rows_sweep.py 56,56 2>&1 | tail -3
timeout 300 python tools/gpu_micro.py "cfg2" 2>&1 | grep -v "phase timing"
timeout 300 python tools/gpu_micro.py "full all sgm" 2>&1 | grep -v "phase timing"
