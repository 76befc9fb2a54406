set -x
( timeout 900 python -m pytest tests -q -m gpu -k "lm_ or embed or tail or golden" 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r2b_pytest.log
python bench.py --config cfg4 --steps 30 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['step_share'])"
