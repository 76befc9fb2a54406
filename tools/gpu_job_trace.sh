timeout 600 python tools/trace_r2.py 2048 16 9 1024 64 64 bwd > gpurun_out/trace_cfg5_bwd.log 2>&1
tail -2 gpurun_out/trace_cfg5_bwd.log
