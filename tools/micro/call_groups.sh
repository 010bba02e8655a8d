mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/pytest.log 2>&1
cat gpurun_out/pytest.log
for g in 1 2 3; do
  echo "== MGMB200_GROUPS=$g"
  MGMB200_GROUPS=$g timeout 300 python tools/micro/rows_sweep.py 56,56 2>&1 | tail -3
  MGMB200_GROUPS=$g timeout 300 python tools/gpu_micro4.py 56 2>&1 | grep "bands=1:"
done
