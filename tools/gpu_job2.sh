python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][0])
print('n',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'graph',d['config']['cuda_graph'])
"; wc -l gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('n',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
