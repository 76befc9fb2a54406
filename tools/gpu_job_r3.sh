set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "r3_ or generic_regime or lm_model_cfg4" > gpurun_out/r3_tests.log 2>&1
tail -2 gpurun_out/r3_tests.log
for tool in synccheck racecheck; do
timeout 600 compute-sanitizer --tool $tool --print-limit 3 python tools/sanitize_cases.py r3_small 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|case .* done|Barrier error|at vmlmf" | head -8
done
timeout 300 python tools/time_r2.py 20 35 650 650 300 300 > gpurun_out/r3_time.log 2>&1; cat gpurun_out/r3_time.log
