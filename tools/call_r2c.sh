#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exp_ab.py > gpurun_out/r2c_ab.log 2>&1; cat gpurun_out/r2c_ab.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
