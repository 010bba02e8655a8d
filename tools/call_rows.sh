#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/rows_exp2.log; : > $L
FELZ=0 OPTS='[{}, {"full_block":1,"rows_axis":52,"rows_diag":52}, {"full_block":1,"rows_axis":48,"rows_diag":48}, {"full_block":1,"rows_axis":44,"rows_diag":44}, {"full_block":1,"rows_axis":40,"rows_diag":40}, {"full_block":1,"rows_axis":36,"rows_diag":36}, {"full_block":1,"rows_axis":32,"rows_diag":32}, {"full_block":1,"rows_axis":40,"rows_diag":48}]' timeout 300 python tools/exp_lean.py 1920,1080,128,2 2>&1 | grep -v "^  \|Traceback\|\^" >> $L
FELZ=0 OPTS='[{}, {"full_block":1,"rows_axis":48,"rows_diag":48}, {"full_block":1,"rows_axis":40,"rows_diag":40}]' timeout 300 python tools/exp_lean.py 2048,1536,256,3 2>&1 | grep -v "^  \|Traceback\|\^" >> $L
FELZ=0 OPTS='[{}, {"rows_axis":48,"rows_diag":48}, {"rows_axis":40,"rows_diag":40}, {"rows_axis":32,"rows_diag":32}]' timeout 300 python tools/exp_lean.py 4096,4096,64,4 4096,4096,64,2 2>&1 | grep -v "^  \|Traceback\|\^" >> $L
FELZ=1 OPTS='[{}, {"full_block":1,"rows_axis":52,"rows_diag":52}]' timeout 300 python tools/exp_lean.py 2048,1536,256,3 2>&1 | grep -v "^  \|Traceback\|\^" >> $L
cat $L
