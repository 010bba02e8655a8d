#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_nj.log
timeout 300 python tools/exp_ab.py 1920,1080,128,2,0 2048,1536,256,3,1 2048,1536,256,3,0 4096,4096,64,2,0 >> gpurun_out/r2_nj.log 2>&1
MGMB200_LIBRARY=$PWD/mgm_b200/variants/libmgmb200_nj4.so timeout 300 python tools/exp_ab.py 1920,1080,128,2,0 4096,4096,64,2,0 >> gpurun_out/r2_nj.log 2>&1
MGMB200_LIBRARY=$PWD/mgm_b200/variants/libmgmb200_nj8.so timeout 300 python tools/exp_ab.py 2048,1536,256,3,1 2048,1536,256,3,0 >> gpurun_out/r2_nj.log 2>&1
cat gpurun_out/r2_nj.log
