# regime R3 (small-batch, weight-stationary) development job: parity, one-layer timings R3 vs R2, cycle traces of one CTA
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "r3_ or generic_regime or lm_model_cfg4 or randomised" 2>&1 | tail -2
timeout 300 python tools/time_r2.py 20 35 650 650 300 300 10 > gpurun_out/r3_time.log 2>&1
VMLMF_NO_R3=1 timeout 300 python tools/time_r2.py 20 35 650 650 300 300 10 >> gpurun_out/r3_time.log 2>&1
grep "shape\|ms:" gpurun_out/r3_time.log
timeout 600 python tools/trace_r2.py 20 12 650 650 300 300 > gpurun_out/trace_r3_lm20_fwd.log 2>&1
timeout 600 python tools/trace_r2.py 20 12 650 650 300 300 bwd > gpurun_out/trace_r3_lm20_bwd.log 2>&1
