set -x
python -m pytest tests -q -m gpu 2>&1 | tail -15
python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 3000 gpurun_out/bench_a.json; tail -5 gpurun_out/bench_a.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1.csv python tools/prof_step.py 8192 3 > gpurun_out/ncu_launch.log 2>&1; tail -3 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:seq_.*_r1 -s 2 -c 2 -o gpurun_out/prof_r1 -f python tools/prof_step.py 8192 2 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
