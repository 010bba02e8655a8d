#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "costvolume or golden or ranges or midsize or stereo" > gpurun_out/ncc_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/ncc_pytest.log
tail -6 gpurun_out/ncc_pytest.log
python tools/micro/cv_time.py 2>&1 | tail -5
timeout 600 python bench.py --workload cfg5_4096x4096x64_ncc5_O16_TSGM4 --steps 3 --warmup 2 2>gpurun_out/ncc_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e'], 'parity', d.get('parity'))"
