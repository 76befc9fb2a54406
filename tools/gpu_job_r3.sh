# regime R3: cycle trace of one CTA at the LM shape (debug build), then the whole GPU suite
set -x
timeout 600 python tools/trace_r2.py 20 12 650 650 300 300 > gpurun_out/trace_r3_lm20_fwd.log 2>&1
tail -3 gpurun_out/trace_r3_lm20_fwd.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r3_tests.log 2>&1
tail -15 gpurun_out/r3_tests.log
