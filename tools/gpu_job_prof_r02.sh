# round-2 evidence captures (one B200).  Outputs under gpurun_out/, summarised into profiles/ by tools/ncu_summary.py.
set -x
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:"r2_fwd_kernel|r2_bwd_kernel|gemm_tn_kernel" -c 4 -o gpurun_out/r02_r2_cfg5 python tools/time_r2.py 2048 16 9 1024 64 64 1 > gpurun_out/prof_a.log 2>&1
timeout 900 $NCU -k regex:"r2_fwd_kernel|r2_bwd_kernel" -c 3 -o gpurun_out/r02_r2_lm512 python tools/time_r2.py 512 35 650 650 300 300 1 > gpurun_out/prof_b.log 2>&1
timeout 900 $NCU -k regex:"seq_fwd_r1_kernel|seq_bwd_r1_kernel" -s 4 -c 2 -o gpurun_out/r02_seq_r1_cfg1 python tools/prof_cfg1.py 64 2 > gpurun_out/prof_c.log 2>&1
timeout 900 $NCU -k regex:"seq_fwd_mma|seq_bwd_fused|dux_rows|xproj_small" -s 4 -c 4 -o gpurun_out/r02_mma_full python tools/prof_step.py 9472 2 > gpurun_out/prof_d.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs > gpurun_out/prof_e.log 2>&1
tail -2 gpurun_out/prof_?.log
ls -la gpurun_out/r02_*
