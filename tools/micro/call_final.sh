mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -1 gpurun_out/bench_cfg3.json | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_cfg3.err; tail -1 gpurun_out/bench_ref.json | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python bench.py --workload cfg2_1920x1080x128_census5_O8_TSGM2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2>> gpurun_out/bench_cfg3.err; tail -1 gpurun_out/bench_cfg2.json | cut -c1-300
