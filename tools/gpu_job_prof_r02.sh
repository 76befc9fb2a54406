# round-2 evidence captures (one B200).  Outputs under gpurun_out/, summarised into profiles/ by tools/ncu_summary.py.
set -x
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:"r2_fwd_kernel" -s 2 -c 2 -o gpurun_out/r02_r2_cfg5_fwd python tools/time_r2.py 2048 16 9 1024 64 64 1 > gpurun_out/prof_a.log 2>&1
timeout 900 $NCU -k regex:"r2_bwd_kernel|gemm_tn_kernel|colreduce4" -c 4 -o gpurun_out/r02_r2_cfg5_bwd python tools/time_r2.py 2048 16 9 1024 64 64 1 > gpurun_out/prof_a2.log 2>&1
timeout 900 $NCU -k regex:"r2_fwd_kernel|r2_bwd_kernel" -s 2 -c 2 -o gpurun_out/r02_r2_lm512 python tools/time_r2.py 512 35 650 650 300 300 1 > gpurun_out/prof_b.log 2>&1
timeout 900 $NCU -k regex:"r3_fwd_kernel|r3_bwd_kernel" -s 2 -c 2 -o gpurun_out/r02_r3_lm20 python tools/time_r2.py 20 35 650 650 300 300 1 > gpurun_out/prof_g.log 2>&1
timeout 900 $NCU -k regex:"seq_fwd_r1_kernel|seq_bwd_r1_kernel" -s 4 -c 2 -o gpurun_out/r02_seq_r1_cfg1 python tools/prof_cfg1.py 64 2 > gpurun_out/prof_c.log 2>&1
timeout 900 $NCU -k regex:"gemm_tc_kernel|gemm_tn_kernel" -s 3 -c 3 -o gpurun_out/r02_gemm_lmhead python -c "
import torch, sys
sys.path.insert(0, '.')
from vmlmf_b200.functional import linear_tc
x = torch.randn(17920, 650, device='cuda', requires_grad=True); w = (torch.randn(10000, 650, device='cuda') * 0.05).requires_grad_(True); b = torch.zeros(10000, device='cuda', requires_grad=True)
for _ in range(2):
    y = linear_tc(x, w, b); y.sum().backward()
torch.cuda.synchronize()" > gpurun_out/prof_f.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs > gpurun_out/prof_e.log 2>&1
for f in gpurun_out/prof_?.log gpurun_out/prof_a2.log; do tail -n 1 $f | cut -c1-200; done
ls -la gpurun_out/r02_*.ncu-rep
