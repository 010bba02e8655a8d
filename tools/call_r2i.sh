#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest.log
tail -6 gpurun_out/r2i_pytest.log
timeout 300 python bench.py --workload cfg4_32x1242x375x192_ad_O8_TSGM4 --steps 3 --warmup 2 --no-parity 2>>gpurun_out/r2i.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'], d['value'], d['ms_per_step'], 'e2e', d['e2e'])"
tail -3 gpurun_out/r2i.err
