#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/rows45.log; : > $L
run() { echo "== $WL $*" >> $L; env "$@" timeout 300 python bench.py --workload $WL --steps 3 --warmup 2 --no-parity 2>>gpurun_out/rows45.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['launch_info'], 'e2e', d['e2e']['ms_per_step'])" >> $L; }
WL=cfg4_32x1242x375x192_ad_O8_TSGM4
run A=1; run MGMB200_ROWS_AXIS=48 MGMB200_ROWS_DIAG=48; run MGMB200_ROWS_AXIS=40 MGMB200_ROWS_DIAG=40; run MGMB200_ROWS_AXIS=32 MGMB200_ROWS_DIAG=32
WL=cfg5_4096x4096x64_ncc5_O16_TSGM4
run A=1; run MGMB200_ROWS_AXIS=48 MGMB200_ROWS_DIAG=48; run MGMB200_ROWS_AXIS=40 MGMB200_ROWS_DIAG=40; run MGMB200_LANES4=1
cat $L
