#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mgm_matches_oracle or alternative or label_counts or midsize or golden or ranges or 16_sweeps or batch or slab" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
tail -12 gpurun_out/r2d_pytest.log
timeout 300 python tools/exp_ab.py 2048,1536,256,3,1 2048,1536,256,2,1 640,480,64,3,1 > gpurun_out/r2d_ab.log 2>&1
MGMB200_PAIR_CHAINS=1 timeout 300 python tools/exp_ab.py 2048,1536,256,3,1 >> gpurun_out/r2d_ab.log 2>&1
cat gpurun_out/r2d_ab.log
