#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest.log
tail -5 gpurun_out/r2j_pytest.log
for e in "" MGMB200_LR_SEQUENTIAL=1 MGMB200_FUSED_FINISH=1; do
env $e timeout 300 python bench.py --steps 3 --warmup 2 --no-parity 2>>gpurun_out/r2j.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$e', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'cli', d['e2e_cli_flow'])"
done
env timeout 300 python bench.py --workload cfg2_1920x1080x128_census5_O8_TSGM2 --steps 3 --warmup 2 --no-parity 2>>gpurun_out/r2j.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'cli', d['e2e_cli_flow'])"
