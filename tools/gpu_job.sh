# round-2 single-GPU job: tests, smoke, bench (ours + reference arm)
set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | tail -3
tail -c 400 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
( time python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/r02_bench_n1_reference.json 2> gpurun_out/r02_bench_ref.err ) 2>&1 | tail -3
tail -c 300 gpurun_out/r02_bench_n1_reference.json
