python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n8.json') if l.startswith('{')][0])
print('n',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'graph',d['config']['cuda_graph'], d['clocks'])
"; wc -l gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err
