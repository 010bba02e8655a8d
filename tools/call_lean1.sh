#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lean or alternative or midsize or matches_oracle" > gpurun_out/lean1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/lean1_pytest.log
tail -15 gpurun_out/lean1_pytest.log
timeout 600 python tools/exp_lean.py > gpurun_out/lean1_exp.log 2>&1
cat gpurun_out/lean1_exp.log
