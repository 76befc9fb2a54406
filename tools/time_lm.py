#!/usr/bin/env python
"""cfg4: LM training step Model(10000,650,2,dropout 0,0.05,300,[300],"vmlmf"), x[35,B] -- times the fused layer calls
(CUDA events) and the whole step with the tcgen05 GEMM and with the SIMT GEMM.  Development aid."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vmlmf_b200 as vb
from vmlmf_b200 import functional as F

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = "cuda:0"
torch.manual_seed(3)
model = vb.Model(10000, 650, 2, 0.0, 0.05, 300, [300], "vmlmf").to(dev)
x = torch.randint(0, 10000, (35, B), device=dev)
y = torch.randint(0, 10000, (35 * B,), device=dev)
states = model.state_init(B)

def step():
    global states
    model.zero_grad()
    states = model.detach(states)
    scores, states = model(x, states)
    loss = torch.nn.functional.cross_entropy(scores, y)
    loss.backward()
    return loss

for _ in range(3):
    step()
torch.cuda.synchronize()
F.EVENT_LOG = []
t0 = time.perf_counter()
n = 10
for _ in range(n):
    step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
ev = {}
for name, a, b in F.EVENT_LOG:
    ev.setdefault(name, []).append(a.elapsed_time(b))
F.EVENT_LOG = None
print(f"B={B} G_SIMT={os.environ.get('VMLMF_G_SIMT')} step {dt*1e3:.2f} ms  {35*B/dt:.0f} tokens/s  {B/dt:.0f} seq/s ",
      {k: round(sum(v) / n, 3) for k, v in ev.items()}, "(ms per step, both layers)")
