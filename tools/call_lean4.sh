#!/bin/bash
# multi-CTA-per-SM variants of the lean truncated-linear kernels
mkdir -p gpurun_out
L=gpurun_out/lean4_exp.log
: > $L
C="2048,1536,256,3"
FELZ=1 OPTS='[{}, {"rows_axis":40,"rows_diag":40}, {"rows_axis":28,"rows_diag":28}]' timeout 300 python tools/exp_lean.py $C >> $L 2>&1
for v in "2 28" "3 18" "4 14" "2 24" "3 16" "4 12"; do set -- $v
  echo "== lb$1 rows $2" >> $L
  MGMB200_LIBRARY=$PWD/mgm_b200/variants/libmgmb200_lb$1.so FELZ=1 OPTS="[{\"rows_axis\":$2,\"rows_diag\":$2}]" timeout 300 python tools/exp_lean.py $C >> $L 2>&1
done
cat $L
