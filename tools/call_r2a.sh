#!/bin/bash
# GPU call A of round 2: full GPU suite, bench with the full-size parity gate, reference arm, experiment sweeps
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt; nproc >> gpurun_out/r2a_gpu.txt; free -g >> gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -3 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err; tail -c 1500 gpurun_out/r2a_bench_cfg3.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_bench_ref.json 2>> gpurun_out/r2a_bench_cfg3.err; tail -c 600 gpurun_out/r2a_bench_ref.json
timeout 600 python bench.py --workload cfg2_1920x1080x128_census5_O8_TSGM2 --steps 5 --warmup 3 > gpurun_out/r2a_bench_cfg2.json 2>> gpurun_out/r2a_bench_cfg3.err; tail -c 1200 gpurun_out/r2a_bench_cfg2.json
timeout 600 python bench.py --workload cfg4_32x1242x375x192_ad_O8_TSGM4 --steps 3 --warmup 3 > gpurun_out/r2a_bench_cfg4.json 2>> gpurun_out/r2a_bench_cfg3.err; tail -c 1200 gpurun_out/r2a_bench_cfg4.json
timeout 900 python tools/exp_r2a.py sgm > gpurun_out/r2a_exp_sgm.log 2>&1
timeout 600 python tools/exp_r2a.py axis > gpurun_out/r2a_exp_axis.log 2>&1
timeout 300 python tools/exp_r2a.py chain > gpurun_out/r2a_exp_chain.log 2>&1
MGMB200_LIBRARY=$PWD/mgm_b200/variants/libmgmb200_pf2.so timeout 300 python tools/exp_r2a.py chain >> gpurun_out/r2a_exp_chain.log 2>&1
MGMB200_LIBRARY=$PWD/mgm_b200/variants/libmgmb200_pf4.so timeout 300 python tools/exp_r2a.py chain >> gpurun_out/r2a_exp_chain.log 2>&1
tail -5 gpurun_out/r2a_exp_chain.log
