#!/usr/bin/env python
"""Error of the fused kernels against the fp64 numpy spec at a cfg2-shaped problem (warp-MMA regime forced).
usage: python tools/precision_probe.py [B [RX RH [T I H]]]   (ranks beyond 16 or H > 256 plan the generic regime;
       the LM layer is `512 300 300 35 650 650`)"""
import os, sys
import numpy as np
import torch
os.environ.setdefault("VMLMF_MMA_MIN_BATCH", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import canonical_numpy as cn
from vmlmf_b200.functional import vmlmf_sequence

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T, I, H = (int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else (24, 77, 256)
RX = int(sys.argv[2]) if len(sys.argv) > 3 else 8
RH = int(sys.argv[3]) if len(sys.argv) > 3 else 6
rng = np.random.default_rng(0)
scale = 0.03 if H > 256 else 0.1          # the LM initialises U(-0.05, 0.05)
f = lambda *s: (rng.standard_normal(s) * scale).astype(np.float32)
cp = dict(Ux=f(I, RX), Vx=f(4 * H, RX), Dx=f(4, I), A=f(H, RH), Bm=f(4 * H, RH), Dh=f(4, H), bias=f(4 * H))
x = rng.standard_normal((T, B, I)).astype(np.float32)
dy = rng.standard_normal((T, B, H)).astype(np.float32)
dhT = rng.standard_normal((B, H)).astype(np.float32)
cp64 = {k: v.astype(np.float64) for k, v in cp.items()}
y64, hT64, cT64, saved = cn.forward(cp64, x.astype(np.float64))
g64 = cn.backward(cp64, x.astype(np.float64), y64, saved, dy.astype(np.float64), dhT.astype(np.float64), None)
names = ("Ux", "Vx", "Dx", "A", "Bm", "Dh", "bias")
tp = [torch.from_numpy(cp[k]).cuda().requires_grad_(True) for k in names]
xt = torch.from_numpy(np.ascontiguousarray(x.transpose(1, 0, 2))).cuda().requires_grad_(True)
y, hT, cT = vmlmf_sequence(xt, None, None, tp, batch_first=True)
torch.autograd.backward([y, hT], [torch.from_numpy(np.ascontiguousarray(dy.transpose(1, 0, 2))).cuda(), torch.from_numpy(dhT).cuda()])
def err(a, b):
    a = a.astype(np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30), np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
print("y   ", err(y.detach().cpu().numpy().transpose(1, 0, 2), y64))
print("cT  ", err(cT.detach().cpu().numpy(), cT64))
for k, t in zip(names, tp):
    print(f"d{k:5s}", err(t.grad.cpu().numpy(), g64[k]))
print("dx  ", err(xt.grad.cpu().numpy().transpose(1, 0, 2), g64["dx"]))
