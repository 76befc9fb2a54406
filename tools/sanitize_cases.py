#!/usr/bin/env python
"""Small forward+backward cases, one per kernel family, for compute-sanitizer runs (memcheck / racecheck / synccheck):
   SIMT R1, warp-MMA forward + fused backward (TMEM accumulators), split backward, regime R2 (cooperative tcgen05 recurrence,
   two group sizes), regime R3 (small batch), the tcgen05 GEMMs (NT, TN), softmax-NLL / head / optimizers.   usage: sanitize_cases.py [case ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vmlmf_b200 as vb
from vmlmf_b200.functional import vmlmf_sequence, gemm_nt, gemm_tn

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)


def seq(T, B, I, H, RX, RH, sc=0.1):
    r = lambda *s: (torch.randn(*s, device=dev, generator=g) * sc).requires_grad_(True)
    canon = [r(I, RX), r(4 * H, RX), r(4, I), r(H, RH), r(4 * H, RH), r(4, H), r(4 * H)]
    x = torch.randn(T, B, I, device=dev, generator=g).requires_grad_(True)
    h0 = (torch.randn(B, H, device=dev, generator=g) * 0.3).requires_grad_(True)
    c0 = (torch.randn(B, H, device=dev, generator=g) * 0.3).requires_grad_(True)
    y, hT, cT = vmlmf_sequence(x, h0, c0, canon, batch_first=False)
    (y.sum() + hT.sum() + cT.sum()).backward()
    torch.cuda.synchronize()


CASES = {
    "r1_simt": lambda: seq(4, 10, 9, 32, 8, 6),
    "r1m_fused": lambda: (os.environ.__setitem__("VMLMF_MMA_MIN_BATCH", "1"), seq(3, 40, 9, 64, 8, 6)),
    "r1m_split": lambda: (os.environ.__setitem__("VMLMF_MMA_MIN_BATCH", "1"), os.environ.__setitem__("VMLMF_BWD_SPLIT", "1"), seq(3, 40, 9, 64, 8, 6)),
    "r2_group": lambda: seq(3, 140, 12, 96, 20, 24, 0.08),
    "r3_small": lambda: seq(3, 20, 12, 96, 20, 40, 0.08),          # B <= 32: weight-stationary regime, 12 CTAs, named-barrier hand-over
    "r2_single": lambda: (os.environ.__setitem__("VMLMF_R2_CLUSTER", "1"), seq(2, 40, 8, 64, 20, 20, 0.08)),
    "gemm": lambda: (gemm_nt(torch.randn(200, 72, device=dev), torch.randn(150, 72, device=dev)),
                     gemm_tn(torch.randn(300, 136, device=dev), torch.randn(300, 40, device=dev)), torch.cuda.synchronize()),
    "tail": lambda: _tail(),
}


def _tail():
    torch.manual_seed(3)
    net = vb.Net(9, [32], w_rank=8, u_rank=[6], cell=vb.MyVMLMFCell).to(dev)
    opt = vb.FlatAdam(net, lr=0.002)
    x = torch.randn(12, 6, 9, device=dev)
    y = torch.randint(0, 6, (12,), device=dev)
    for _ in range(2):
        opt.zero_grad()
        vb.cross_entropy(net(x), y).backward()
        opt.step()
    torch.cuda.synchronize()


for name in (sys.argv[1:] or list(CASES)):
    CASES[name]()
    print("case", name, "done")
