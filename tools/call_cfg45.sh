#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_cfg45b.log
run() { echo "== $*" >> gpurun_out/r2_cfg45b.log; env "$@" timeout 300 python bench.py --workload $WL --steps 3 --warmup 2 --no-parity 2>>gpurun_out/r2_cfg45.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['launch_info'], d['e2e']['value'])" >> gpurun_out/r2_cfg45b.log; }
WL=cfg4_32x1242x375x192_ad_O8_TSGM4
run MGMB200_BATCH=4; run MGMB200_BATCH=8; run MGMB200_BATCH=16; run MGMB200_BATCH=32; run MGMB200_BATCH=8 MGMB200_ROWS_AXIS=40 MGMB200_ROWS_DIAG=40
WL=cfg5_4096x4096x64_ncc5_O16_TSGM4
run MGMB200_BATCH=8
WL=cfg2_1920x1080x128_census5_O8_TSGM2
run MGMB200_BATCH=8
cat gpurun_out/r2_cfg45b.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
