#!/usr/bin/env python
"""Times one layer of the large-H regimes (CUDA events around the C-ABI calls): training forward, inference forward,
backward.  Development aid; never a bench value.   usage: time_r2.py B T I H RX RH [n]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmlmf_b200 import functional as F, _lib
from vmlmf_b200.functional import vmlmf_sequence

B, T, I, H, RX, RH = [int(v) for v in sys.argv[1:7]]
n = int(sys.argv[7]) if len(sys.argv) > 7 else 5
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
sc = 0.05 if H >= 300 else 0.1
r = lambda *s: torch.randn(*s, device=dev, generator=g) * sc
canon = [r(I, RX), r(4 * H, RX), r(4, I), r(H, RH), r(4 * H, RH), r(4, H), r(4 * H)]
for p in canon:
    p.requires_grad_(True)
x = torch.randn(T, B, I, device=dev, generator=g)
print("shape B,T,I,H,RX,RH =", (B, T, I, H, RX, RH), "path", _lib.plan(T, B, I, H, RX, RH).path,
      "NO_R2=", os.environ.get("VMLMF_NO_R2"), "CLUSTER=", os.environ.get("VMLMF_R2_CLUSTER"))


def run(mode):
    F.EVENT_LOG = []
    for i in range(n + 2):
        if mode == "infer":
            with torch.no_grad():
                vmlmf_sequence(x, None, None, canon, False)
        else:
            y, hT, cT = vmlmf_sequence(x, None, None, canon, False)
            if mode == "train":
                (hT.sum() + y.sum()).backward()
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in F.EVENT_LOG:
        out.setdefault(name, []).append(e0.elapsed_time(e1))
    F.EVENT_LOG = None
    return {k: round(sorted(v[2:])[len(v[2:]) // 2], 3) for k, v in out.items()}


print("infer ms:", run("infer"))
print("train ms:", run("train"))
