mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/pytest.log 2>&1
MGMB200_PHASE_TIMING=1 timeout 300 python tools/gpu_micro.py "1 band" 2>&1 | awk '/phase timing/{l[$4]=$0} /1 band|us\/step/{for(k in l)print l[k]; delete l; print}' > gpurun_out/phase.log
timeout 300 python tools/gpu_micro.py "full" > gpurun_out/full.log 2>&1
cat gpurun_out/pytest.log gpurun_out/phase.log gpurun_out/full.log
