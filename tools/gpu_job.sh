set -x
python -m pytest tests -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 600 gpurun_out/bench_r1.json; tail -3 gpurun_out/bench_r1.err
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_r1_ref.json 2>gpurun_out/bench_r1_ref.err; tail -c 300 gpurun_out/bench_r1_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:"seq_fwd_mma|seq_bwd_fused|dux_rows|xproj_small" -s 4 -c 4 -o gpurun_out/r01_mma_full -f python tools/prof_step.py 9472 2 > gpurun_out/ncu_full5.log 2>&1; tail -2 gpurun_out/ncu_full5.log
python tools/config_sweep.py gpurun_out/r01_configs.json > /dev/null 2> gpurun_out/configs.err; tail -2 gpurun_out/configs.err
