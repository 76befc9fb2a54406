set -x
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_try.json 2> gpurun_out/bench_try.err
tail -5 gpurun_out/bench_try.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_try.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","inference","gpu_launches")})
print("roofline", {k:d["roofline"].get(k) for k in ("kernel","bound","achieved","frac","regime")})
for c in d["configs"] or []:
    if "error" in c: print(c); continue
    print(c["name"], c["regime"], "train %.4g seq/s %.3f ms | infer %.4g | bwd frac %s fwd frac %s inf frac %s" % (c["train"]["value"], c["train"]["ms_per_step"], c["inference"]["value"],
          c["roofline"]["bwd"] and (c["roofline"]["bwd"]["bound"], round(c["roofline"]["bwd"]["frac"],3)), c["roofline"]["fwd"] and round(c["roofline"]["fwd"]["frac"],3), c["roofline"]["inference"] and round(c["roofline"]["inference"]["frac"],3)))
print(d["cpu_baseline"])
PY
