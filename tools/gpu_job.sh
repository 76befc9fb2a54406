for b in 8192 9472; do
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --batch $b 2>gpurun_out/bench_g.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('B',d['config']['per_gpu_batch'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'inf',d['inference']['value'])
print(d['roofline']['step_share'], d['roofline']['frac'], d['roofline']['other_kernels'][0]['frac'])
"; done
