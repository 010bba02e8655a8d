mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5)
python bench.py --steps 5 --warmup 3 --no-cpu-baseline | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'agg', d['roofline']['ms_per_launch'], 'frac', d['roofline']['frac'], 'wta', d['roofline']['finish_kernel']['ms_per_launch'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
