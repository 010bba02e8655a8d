#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/lean7_exp.log
: > $L
FELZ=1 OPTS='[{}, {"full_block":1,"rows_axis":48,"rows_diag":48}, {"full_block":1,"rows_axis":40,"rows_diag":40}, {"full_block":1,"rows_axis":48}, {"full_block":1,"rows_axis":40,"rows_diag":48}]' timeout 300 python tools/exp_lean.py 2048,1536,256,3 >> $L 2>&1
FELZ=0 OPTS='[{}, {"full_block":1,"rows_axis":48,"rows_diag":48}, {"full_block":1,"rows_axis":40,"rows_diag":40}, {"full_block":1,"rows_axis":32,"rows_diag":32}]' timeout 300 python tools/exp_lean.py 1920,1080,128,2 >> $L 2>&1
cat $L
bash tools/call_ncu.sh r02b "cfg3_2048x1536x256_census3_O8_TSGM3_trunclinear cfg2_1920x1080x128_census5_O8_TSGM2" 2>&1 | tail -3
