#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab.log
for lib in "" $PWD/mgm_b200/variants/*.so; do
  MGMB200_LIBRARY=$lib timeout 300 python tools/exp_ab.py >> gpurun_out/ab.log 2>&1
done
cat gpurun_out/ab.log
