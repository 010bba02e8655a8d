#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/lean8_exp.log
: > $L
for lib in "" $PWD/mgm_b200/variants/*.so; do
  echo "== lib ${lib##*/}" >> $L
  MGMB200_LIBRARY=$lib FELZ=1 OPTS='[{}]' timeout 300 python tools/exp_lean.py 2048,1536,256,3 2>&1 | grep -v "^  \|Traceback\|\^" >> $L
done
cat $L
