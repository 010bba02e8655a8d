#!/bin/bash
# ncu captures of a round: launch list + one --set full capture of the aggregation launch of ONE workload per call
# (gpurun merges at most 64 MiB back)
mkdir -p gpurun_out
R=${1:-r02}
for wl in ${2:-cfg3_2048x1536x256_census3_O8_TSGM3_trunclinear}; do
  short=${wl%%_*}
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_bench_${short}.csv python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu_launch_${short}.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mgm_aggregate" -s 1 -c 1 -f -o gpurun_out/${R}_full_${short} python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu_full_${short}.log 2>&1
  ls -la gpurun_out/${R}_full_${short}.ncu-rep
done
