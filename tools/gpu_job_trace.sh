timeout 600 python tools/trace_r2.py 512 16 650 650 300 300 > gpurun_out/trace_lm512.log 2>&1
timeout 600 python tools/trace_r2.py 2048 16 9 1024 64 64 > gpurun_out/trace_cfg5.log 2>&1
tail -2 gpurun_out/trace_cfg5.log
