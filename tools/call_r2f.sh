#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
tail -6 gpurun_out/r2f_pytest.log
for WL in cfg5_4096x4096x64_ncc5_O16_TSGM4 cfg4_32x1242x375x192_ad_O8_TSGM4 small_640x480x64_census3_O8_TSGM3_trunclinear; do
timeout 300 python bench.py --workload $WL --steps 3 --warmup 2 --no-parity 2>>gpurun_out/r2f.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'], d['value'], d['ms_per_step'], d['launch_info'], d['e2e']['value'], d['roofline']['frac'])"
done
