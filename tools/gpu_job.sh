set -x
python -m pytest tests -q -m gpu 2>&1 | tail -15
build/ubench > gpurun_out/ubench.txt 2>&1; cat gpurun_out/ubench.txt
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -c 2500 gpurun_out/bench_b.json; tail -5 gpurun_out/bench_b.err
