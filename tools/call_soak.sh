#!/bin/bash
# soak: the GPU suite three times and long bench loops with the parity gate (rare races in the band hand-off would show as a
# hang -- every command sits under its own timeout -- or as a mismatch)
mkdir -p gpurun_out
L=gpurun_out/soak.log; : > $L
for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1 >> $L; done
for spec in "cfg3_2048x1536x256_census3_O8_TSGM3_trunclinear 40" "cfg2_1920x1080x128_census5_O8_TSGM2 120" "cfg5_4096x4096x64_ncc5_O16_TSGM4 10" "small_640x480x64_census3_O8_TSGM3_trunclinear 400" "cfg4_32x1242x375x192_ad_O8_TSGM4 6"; do set -- $spec
  timeout 600 python bench.py --workload $1 --steps $2 --warmup 3 2>>gpurun_out/soak.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'][:5], 'steps', d['steps'], d['ms_per_step'], 'parity', d['parity']['ok'], d['parity']['wta_mismatch'])" >> $L 2>&1 || echo "FAILED $1" >> $L
done
cat $L
