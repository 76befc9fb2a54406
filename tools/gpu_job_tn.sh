timeout 300 python tools/test_tn.py 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tensor_core or generic_regime or lm_model" 2>&1 | tail -4
timeout 300 python tools/time_r2.py 512 35 650 650 300 300 3 2>&1 | tail -2
python bench.py --config cfg4 --steps 30 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')})"
