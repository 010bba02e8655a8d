mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/pytest.log 2>&1
cat gpurun_out/pytest.log
timeout 300 python tools/micro/rows_sweep.py 56,56 48,56 40,56 2>&1 | tail -9
echo "== static order"
MGMB200_STATIC_ORDER=1 timeout 300 python tools/micro/rows_sweep.py 56,56 2>&1 | tail -3
timeout 300 python tools/gpu_micro.py "full all" 2>&1 | grep -v "phase timing"
timeout 300 python tools/gpu_micro.py "cfg2" 2>&1 | grep -v "phase timing"
