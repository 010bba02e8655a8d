#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "costvolume or ranges or golden or midsize or stereo or 16_sweeps or cli" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest.log
tail -6 gpurun_out/r2h_pytest.log
timeout 300 python bench.py --workload cfg5_4096x4096x64_ncc5_O16_TSGM4 --steps 3 --warmup 2 2>>gpurun_out/r2h.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'], d['value'], d['ms_per_step'], 'e2e', d['e2e'], d['parity'])"
