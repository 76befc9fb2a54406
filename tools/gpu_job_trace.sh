timeout 600 python tools/trace_r2.py 20 8 650 650 300 300 > gpurun_out/trace_lm20_fwd.log 2>&1
timeout 600 python tools/trace_r2.py 20 8 650 650 300 300 bwd > gpurun_out/trace_lm20_bwd.log 2>&1
timeout 600 python tools/trace_r2.py 512 8 650 650 300 300 > gpurun_out/trace_lm512_fwd.log 2>&1
tail -2 gpurun_out/trace_lm20_fwd.log gpurun_out/trace_lm20_bwd.log
