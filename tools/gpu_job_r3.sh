set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r3_tests.log 2>&1; tail -3 gpurun_out/r3_tests.log
timeout 300 python tools/time_r2.py 2048 128 9 1024 64 64 3 > gpurun_out/r3_time.log 2>&1
timeout 300 python tools/time_r2.py 512 35 650 650 300 300 >> gpurun_out/r3_time.log 2>&1
timeout 300 python tools/time_r2.py 8192 24 77 256 32 32 >> gpurun_out/r3_time.log 2>&1
timeout 300 python tools/time_r2.py 20 35 650 650 300 300 >> gpurun_out/r3_time.log 2>&1
grep "shape\|train" gpurun_out/r3_time.log
