set -x
python -m pytest tests -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 3500 gpurun_out/bench_r1.json; tail -5 gpurun_out/bench_r1.err
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_r1_ref.json 2>gpurun_out/bench_r1_ref.err; tail -c 1200 gpurun_out/bench_r1_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:"seq_fwd_mma|seq_bwd_mma|grad_rows" -s 3 -c 3 -o gpurun_out/r01_mma_full -f python tools/prof_step.py 8192 2 > gpurun_out/ncu_full5.log 2>&1; tail -2 gpurun_out/ncu_full5.log
