#!/bin/bash
# round-2 record run on one GPU: GPU suite, bench lines of every workload (parity gate on), reference arm, ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r02_gpu.txt; nproc >> gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log
tail -3 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n1.json 2> gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench.err
for wl in cfg2_1920x1080x128_census5_O8_TSGM2 cfg4_32x1242x375x192_ad_O8_TSGM4 cfg5_4096x4096x64_ncc5_O16_TSGM4 small_640x480x64_census3_O8_TSGM3_trunclinear; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/r02_bench_${wl%%_*}_n1.json 2>> gpurun_out/r02_bench.err
done
for f in gpurun_out/r02_bench_*_n1.json gpurun_out/r02_bench_reference_arm.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', d['value'], d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), d['e2e']['value'], (d.get('parity') or {}).get('ok'), (d.get('parity') or {}).get('wta_mismatch'), (d.get('cpu_baseline') or {}).get('value'))"; done
tail -3 gpurun_out/r02_bench.err
