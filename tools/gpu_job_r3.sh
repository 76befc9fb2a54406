set -x
timeout 900 python tools/fuzz_regimes.py 40 1 > gpurun_out/fuzz.log 2>&1; tail -45 gpurun_out/fuzz.log
timeout 600 python -m pytest tests/test_gpu_tail.py -x -q -m gpu 2>&1 | tail -2
