for v in "" "MGMB200_ROWS_AXIS=48" "MGMB200_ROWS_AXIS=40" "MGMB200_ROWS_AXIS=32" "MGMB200_ROWS_AXIS=40 MGMB200_NO_FUSED_FINISH=1"; do
  echo "== $v"
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline'].get('unfused_split', {}).get('aggregation_only', d['roofline']); print('ms_per_step', d['ms_per_step'], 'agg-only', r['ms_per_launch'], 'e2e', d['e2e']['ms_per_step'], d['launch_info']['rows_axis'], d['launch_info']['rows_diag'])
    elif 'Error' in l or 'error' in l: print(l.strip())
"
done
