set -x
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "generic_regime or lm_model_cfg4 or r2_ or lm_layer or lm_model" 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r2b_pytest.log
timeout 300 python tools/time_r2.py 512 35 650 650 300 300 3 2>&1 | tail -3
python bench.py --config cfg4 --steps 30 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['step_share'])"
