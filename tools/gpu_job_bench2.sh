set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -8 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n2.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","replicas_in_sync","grad_allreduce_bytes")})
print(d["config"])
print(d["lm_data_parallel"])
PY
