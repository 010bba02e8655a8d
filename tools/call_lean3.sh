#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lean or alternative or midsize or matches_oracle" > gpurun_out/lean3_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/lean3_pytest.log
tail -15 gpurun_out/lean3_pytest.log
FELZ=1 timeout 600 python tools/exp_lean.py 2048,1536,256,3 1920,1080,128,2 640,480,64,3 > gpurun_out/lean3_exp.log 2>&1
cat gpurun_out/lean3_exp.log
