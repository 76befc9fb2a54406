set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r3_tests.log 2>&1
tail -3 gpurun_out/r3_tests.log
timeout 600 python bench.py --config cfg4_b20 --no-configs --no-cpu-baseline --steps 30 > gpurun_out/b20.json 2> gpurun_out/b20.err
timeout 600 python bench.py --config cfg4 --no-configs --no-cpu-baseline --steps 20 > gpurun_out/b512.json 2> gpurun_out/b512.err
python - <<'PY'
import json
for f in ("gpurun_out/b20.json","gpurun_out/b512.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['roofline']['regime'], d['roofline']['step_share'], d['inference'])
PY
