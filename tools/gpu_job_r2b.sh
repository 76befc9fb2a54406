set -x
( timeout 900 python -m pytest tests -q -m gpu -k "fast_tf32 or tensor_core or lm_model or generic_regime" 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r2b_pytest.log
for f in 0 1; do
VMLMF_FAST_TF32=$f python bench.py --config cfg4 --steps 30 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fast=$f', {k:d[k] for k in ('value','ms_per_step')})"
done
