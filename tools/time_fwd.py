#!/usr/bin/env python
"""Times the fused sequence kernels alone (CUDA events) at one shape: training forward (saves state),
inference forward (no saves) and backward.  Development aid; never a bench value.
usage: time_fwd.py [B T I H RX RH]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmlmf_b200 import functional as F
from vmlmf_b200.functional import vmlmf_sequence

a = [int(v) for v in sys.argv[1:7]] if len(sys.argv) >= 7 else [8192, 24, 77, 256, 8, 6]
B, T, I, H, RX, RH = a
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
r = lambda *s: torch.randn(*s, device=dev, generator=g) * 0.1
canon = [r(I, RX), r(4 * H, RX), r(4, I), r(H, RH), r(4 * H, RH), r(4, H), r(4 * H)]
for p in canon:
    p.requires_grad_(True)
xs = [torch.randn(B, T, I, device=dev, generator=g) for _ in range(3)]


def run(mode, n=20):
    F.EVENT_LOG = []
    for i in range(n + 3):
        x = xs[i % 3]
        if mode == "infer":
            with torch.no_grad():
                vmlmf_sequence(x, None, None, canon, True)
        else:
            y, hT, cT = vmlmf_sequence(x, None, None, canon, True)
            if mode == "train":
                hT.sum().backward()
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in F.EVENT_LOG:
        out.setdefault(name, []).append(e0.elapsed_time(e1))
    F.EVENT_LOG = None
    return {k: sorted(v[3:])[len(v[3:]) // 2] for k, v in out.items()}


print("shape", a, "env SIMT=", os.environ.get("VMLMF_R1_SIMT"), "DBG=", os.environ.get("VMLMF_DBG"))
def loop_ms(mode, n=30):
    """GPU-bound timing: n calls back to back between two events (includes xproj ~0.055 ms per call)"""
    for _ in range(3):
        run(mode, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        x = xs[i % 3]
        if mode == "infer":
            with torch.no_grad():
                vmlmf_sequence(x, None, None, canon, True)
        else:
            vmlmf_sequence(x, None, None, canon, True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


import subprocess
def clk():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    except OSError:
        return "n/a"
run("fwd", 5)
print("clocks under load:", clk())
print("loop infer ms/call", loop_ms("infer"), " loop fwd-train ms/call", loop_ms("fwd"))
print("infer", run("infer"))
print("fwd-train", run("fwd"))
print("train", run("train"))
print("clocks after:", clk())
