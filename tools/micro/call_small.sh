(timeout 900 python -m pytest tests -m gpu -x -q -k "weights or stereo or golden or cli" 2>&1 | tail -2)
ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"mgm_weights|mgm_census|mgm_costvolume" --log-file gpurun_out/small.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
for r in csv.reader(open('gpurun_out/small.csv')):
    if len(r) > 5 and 'gpu__time_duration' in r[-3]: print(r[4][:40], r[-1], r[-2])
PY
