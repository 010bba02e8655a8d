#!/bin/bash
# GPU call B of round 2: table-driven launch (batches, 16 sweeps, slab stores), full GPU suite, benches
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_cfg3.json 2> gpurun_out/r2b_bench.err; tail -c 1800 gpurun_out/r2b_bench_cfg3.json
timeout 600 python bench.py --workload cfg2_1920x1080x128_census5_O8_TSGM2 --steps 5 --warmup 3 > gpurun_out/r2b_bench_cfg2.json 2>> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench_cfg2.json
timeout 600 python bench.py --workload cfg4_32x1242x375x192_ad_O8_TSGM4 --steps 3 --warmup 3 > gpurun_out/r2b_bench_cfg4.json 2>> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench_cfg4.json
timeout 600 python bench.py --workload cfg5_4096x4096x64_ncc5_O16_TSGM4 --steps 3 --warmup 3 > gpurun_out/r2b_bench_cfg5.json 2>> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench_cfg5.json
timeout 600 python tools/exp_r2b.py > gpurun_out/r2b_exp.log 2>&1; cat gpurun_out/r2b_exp.log
tail -5 gpurun_out/r2b_bench.err
