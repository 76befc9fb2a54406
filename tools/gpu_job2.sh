set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --no-cpu-baseline > gpurun_out/bench_tail.json 2> gpurun_out/bench_tail.err; tail -c 300 gpurun_out/bench_tail.json; tail -3 gpurun_out/bench_tail.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/tail_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
