#!/usr/bin/env python
"""cfg1-shaped eager train steps for a kernel launch list: python tools/prof_cfg1.py B [steps]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vmlmf_b200 as vb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(3)
net = vb.Net(9, [128], w_rank=8, u_rank=[6], cell=vb.MyVMLMFCell).cuda()
opt = torch.optim.Adam(net.parameters(), lr=0.002, fused=True)
x = torch.randn(B, 128, 9, device="cuda")
y = torch.randint(0, 6, (B,), device="cuda")
def step():
    net.zero_grad()
    torch.nn.functional.cross_entropy(net(x), y).backward()
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / n * 1e3)
