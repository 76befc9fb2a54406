for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; tail -2 gpurun_out/bench_n$n.err | cut -c1-200; head -c 250 gpurun_out/bench_n$n.json; echo
done
