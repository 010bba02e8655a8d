#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/lean5_exp.log
: > $L
C="2048,1536,256,3"
for v in "pf2 56" "pf4 56" "pf2lb2 28" "pf4lb2 28"; do set -- $v
  echo "== $1 rows $2" >> $L
  MGMB200_LIBRARY=$PWD/mgm_b200/variants/libmgmb200_$1.so FELZ=1 OPTS="[{\"rows_axis\":$2,\"rows_diag\":$2}]" timeout 300 python tools/exp_lean.py $C 2>&1 | grep -v "^  \|Traceback\|\^" >> $L
done
cat $L
