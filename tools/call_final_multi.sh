#!/bin/bash
# round-2 record run on N GPUs of one box: bench lines (batch layout = value, sweep-sharded layout beside it) + hardware parity test
N=$1; shift
mkdir -p gpurun_out
for wl in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl > gpurun_out/r02_bench_${wl%%_*}_n$N.json 2>> gpurun_out/r02_bench_n$N.err
  python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_${wl%%_*}_n$N.json').read().strip().splitlines()[-1])
print('${wl%%_*}', 'N=$N value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
for k,v in (d.get('sweep_sharded') or {}).items(): print('   sweep_sharded', k, {a:b for a,b in v.items() if a!='parallelism'})
"
done
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -s 2>&1 | tail -5 | tee gpurun_out/r02_multigpu_test_n$N.log
tail -3 gpurun_out/r02_bench_n$N.err
