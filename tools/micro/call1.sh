set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/c1_pytest.log 2>&1
./tools/micro/chain_bench 43 > gpurun_out/c1_chain.log 2>&1
./tools/micro/chain_bench 32 >> gpurun_out/c1_chain.log 2>&1
MGMB200_PHASE_TIMING=1 timeout 300 python tools/gpu_micro.py "1 band" > gpurun_out/c1_phase.log 2>&1
timeout 600 python tools/micro/rows_sweep.py > gpurun_out/c1_rows.log 2>&1
cat gpurun_out/c1_pytest.log gpurun_out/c1_chain.log gpurun_out/c1_phase.log gpurun_out/c1_rows.log
