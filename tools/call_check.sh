#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/check_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/check_pytest.log
tail -4 gpurun_out/check_pytest.log
for wl in cfg2_1920x1080x128_census5_O8_TSGM2 cfg3_2048x1536x256_census3_O8_TSGM3_trunclinear; do
timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-parity 2>>gpurun_out/check.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'], d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['launch_info'])"
done
