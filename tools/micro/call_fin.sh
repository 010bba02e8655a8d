(timeout 900 python -m pytest tests -m gpu -x -q -k "layouts or full or golden" 2>&1 | tail -3)
for v in "" "MGMB200_FIN_AXIS_MODEL=1" "MGMB200_FIN_TILE=128x8" ; do
  echo "== $v"
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], 'lr', d['e2e_cli_flow']['ms_per_pair'])
    elif 'Error' in l or 'error' in l: print(l.strip())
"
done
for v in "" "MGMB200_FIN_AXIS_MODEL=1"; do
env $v timeout 300 python bench.py --workload cfg2_1920x1080x128_census5_O8_TSGM2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
done
