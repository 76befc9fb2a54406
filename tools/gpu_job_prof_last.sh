set -x
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:"r3_bwd_kernel" -s 1 -c 1 -o gpurun_out/r02_r3_lm20_bwd python tools/time_r2.py 20 35 650 650 300 300 1 > gpurun_out/prof_h.log 2>&1
timeout 600 $NCU -k regex:"gemm_tn_kernel" -c 2 -o gpurun_out/r02_tp_cfg5 python tools/time_r2.py 2048 16 9 1024 64 64 1 > gpurun_out/prof_i.log 2>&1
ls -la gpurun_out/r02_r3_lm20_bwd.ncu-rep gpurun_out/r02_tp_cfg5.ncu-rep
