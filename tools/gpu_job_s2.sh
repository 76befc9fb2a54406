export PYTHONUNBUFFERED=1
for spec in "synccheck r1m_fused" "racecheck gemm" "racecheck r2_group" "racecheck r2_single"; do
  set -- $spec
  echo "=== $1 $2"
  timeout 900 compute-sanitizer --tool $1 --print-limit 3 python tools/sanitize_cases.py $2 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|case .* done|Error|Barrier error|located" | head -8
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 --no-configs > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -4 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n2.json"))
print({k:d[k] for k in ("value","ms_per_step","replicas_in_sync","grad_allreduce_bytes")}, d["config"]["collective"], d["config"]["collective_in_graph"])
PY
