python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -400
python bench.py --no-cpu-baseline --steps 100 > gpurun_out/bench_tail.json 2> gpurun_out/bench_tail.err; tail -3 gpurun_out/bench_tail.err
