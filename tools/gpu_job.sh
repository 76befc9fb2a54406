python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python bench.py --steps 100 --warmup 10 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'inf',d['inference']['value'])
print(d['roofline']['step_share'], d['recurrence_latency'])
"
