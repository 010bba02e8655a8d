#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/lean6_exp.log
: > $L
C="2048,1536,256,3 640,480,64,3"
FELZ=1 OPTS='[{}]' timeout 300 python tools/exp_lean.py $C >> $L 2>&1
echo "== lay0" >> $L
MGMB200_LIBRARY=$PWD/mgm_b200/variants/libmgmb200_lay0.so FELZ=1 OPTS='[{}, {"rows_axis":48,"rows_diag":48}]' timeout 300 python tools/exp_lean.py $C 2>&1 | grep -v "^  \|Traceback\|\^" >> $L
MGMB200_LIBRARY=$PWD/mgm_b200/variants/libmgmb200_lay0.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lean_trunc or midsize" 2>&1 | tail -3 >> $L
cat $L
