#!/usr/bin/env python
"""One LM layer (H=650, ranks 300) forward+backward at B=512, T=35 -- the command profiled under ncu for the
tcgen05 GEMM (regime G).  Never a source of timing numbers."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vmlmf_b200 as vb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = "cuda:0"
torch.manual_seed(3)
layer = vb.MyVMLSTM(650, 650, w_rank=300, u_ranks=300).to(dev)
for p in layer.parameters():
    torch.nn.init.uniform_(p, -0.05, 0.05)
x = torch.randn(35, B, 650, device=dev, requires_grad=True)
h = torch.zeros(B, 650, device=dev); c = torch.zeros(B, 650, device=dev)
for _ in range(2):
    out, (h1, c1) = layer(x, (h, c))
    out.sum().backward()
torch.cuda.synchronize()
print("done")
