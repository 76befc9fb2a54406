set -x
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "generic_regime or lm_model_cfg4 or r2_ or lm_" 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r2b_pytest.log
for a in "2048 128 9 1024 64 64" "512 35 650 650 300 300" "8192 24 77 256 32 32"; do
timeout 300 python tools/time_r2.py $a 3 2>&1 | tail -2
done
