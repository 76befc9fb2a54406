timeout 900 ncu --set full --clock-control none --import-source on -k regex:r2_bwd_kernel -c 1 -o gpurun_out/r2bwd_cfg5 -f python tools/time_r2.py 2048 16 9 1024 64 64 1 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
