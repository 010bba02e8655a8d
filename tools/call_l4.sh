#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/l4.log; : > $L
run() { echo "== $WL $*" >> $L; env "$@" timeout 300 python bench.py --workload $WL --steps 3 --warmup 2 2>>gpurun_out/l4.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['launch_info'], 'e2e', d['e2e']['ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'))" >> $L; }
WL=cfg5_4096x4096x64_ncc5_O16_TSGM4
run A=1; run MGMB200_LANES4=1; run MGMB200_LANES4=1 MGMB200_ROWS_AXIS=96 MGMB200_ROWS_DIAG=96; run MGMB200_LANES4=1 MGMB200_ROWS_AXIS=80 MGMB200_ROWS_DIAG=80
cat $L
