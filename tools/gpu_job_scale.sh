# round-2 multi-GPU job (one 8-GPU box): weak scaling of the headline + the LM data-parallel line
set -x
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err
tail -c 900 gpurun_out/r02_bench_n$n.json | head -c 700; echo; tail -2 gpurun_out/r02_bench_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --config cfg4 --steps 30 --no-configs > gpurun_out/r02_bench_cfg4_n8.json 2> gpurun_out/r02_bench_cfg4_n8.err
tail -c 600 gpurun_out/r02_bench_cfg4_n8.json
