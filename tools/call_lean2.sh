#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/lean2_exp.log
for lib in "" $PWD/mgm_b200/variants/*.so; do
  echo "== lib ${lib##*/}" >> gpurun_out/lean2_exp.log
  MGMB200_LIBRARY=$lib timeout 300 python tools/exp_lean.py $CASES >> gpurun_out/lean2_exp.log 2>&1
done
grep -v "cc_pf" gpurun_out/lean2_exp.log
