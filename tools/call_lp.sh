#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/lp_exp.log; : > $L
for lib in "" $PWD/mgm_b200/variants/libmgmb200_lp16.so; do
  echo "== lib ${lib##*/}" >> $L
  MGMB200_LIBRARY=$lib FELZ=0 OPTS='[{}, {"no_fused_finish":1}]' timeout 300 python tools/exp_lean.py 1920,1080,128,2 640,480,100,2 2>&1 | grep -v "^  \|Traceback\|\^" >> $L
done
cat $L
