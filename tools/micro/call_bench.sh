mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
tail -1 gpurun_out/bench_cfg3.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_cfg3.err
tail -1 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mgm_aggregate|mgm_wta|mgm_costvolume" -s 2 -c 3 -f -o gpurun_out/r01b_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python bench.py --workload cfg2_1920x1080x128_census5_O8_TSGM2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2>> gpurun_out/bench_cfg3.err
tail -1 gpurun_out/bench_cfg2.json
ls -la gpurun_out
