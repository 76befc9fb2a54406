#!/usr/bin/env python
"""A few cfg2 training steps at the bench batch size -- the command profiled under ncu
(see profiles/README.md).  Never a source of timing numbers."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vmlmf_b200 as vb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 9472
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.manual_seed(3)
net = vb.Net(77, [256], w_rank=8, u_rank=[6], cell=vb.MyVMLMFCell).to(dev)
opt = torch.optim.Adam(net.parameters(), lr=0.002, fused=True)
x = torch.randn(B, 24, 77, device=dev)
y = torch.randint(0, 18, (B,), device=dev)
for _ in range(steps):
    net.zero_grad()
    torch.nn.functional.cross_entropy(net(x), y).backward()
    opt.step()
torch.cuda.synchronize()
print("done")
