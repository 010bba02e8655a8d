#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg3_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 3000 gpurun_out/r2_bench_cfg3_n$N.json; tail -5 gpurun_out/r2_bench_n$N.err
