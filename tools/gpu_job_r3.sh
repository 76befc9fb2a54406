set -x
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "r3_ or generic_regime or lm_model_cfg4" 2>&1 | tail -1; done
timeout 300 python tools/time_r2.py 20 35 650 650 300 300 10 > gpurun_out/r3_time.log 2>&1; cat gpurun_out/r3_time.log
timeout 600 python tools/trace_r2.py 20 12 650 650 300 300 > gpurun_out/trace_r3_lm20_fwd.log 2>&1
timeout 600 python tools/trace_r2.py 20 12 650 650 300 300 bwd > gpurun_out/trace_r3_lm20_bwd.log 2>&1
