#!/usr/bin/env python
"""Randomised parity sweep of the large-shape regimes (R2, R3) against the fp64 numpy spec of the canonical recurrence:
random (T, B, I, H, RX, RH), with / without carried state, batch-first or time-major.  Development aid (the fixed cases live
in tests/test_gpu_parity.py).   usage: fuzz_regimes.py [n_cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import canonical_numpy as cn
from vmlmf_b200 import _lib
from vmlmf_b200.functional import vmlmf_sequence

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = "cuda:0"
names = ("Ux", "Vx", "Dx", "A", "Bm", "Dh", "bias")
worst_all, seen = 0.0, {}
for case in range(n_cases):
    small = case % 2 == 0
    B = int(rng.integers(1, 33)) if small else int(rng.integers(33, 300))
    H = int(rng.choice([17, 40, 96, 130, 257, 300, 520, 650, 1000]))
    RH = int(rng.choice([17, 24, 33, 64, 100, 130, 260])) if H >= 96 else int(rng.choice([17, 20, 40]))
    RX = int(rng.choice([3, 8, 20, 64]))
    I = int(rng.integers(1, min(H, 80) + 1))
    T = int(rng.integers(1, 6))
    bf, state = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    path = _lib.plan(T, B, I, H, RX, RH).path
    sc = 0.05 if H >= 300 else 0.12
    f = lambda *s: (rng.standard_normal(s) * sc).astype(np.float32)
    cp = dict(Ux=f(I, RX), Vx=f(4 * H, RX), Dx=f(4, I), A=f(H, RH), Bm=f(4 * H, RH), Dh=f(4, H), bias=f(4 * H))
    x = rng.standard_normal((T, B, I)).astype(np.float32)
    h0 = (rng.standard_normal((B, H)) * .5).astype(np.float32) if state else None
    c0 = (rng.standard_normal((B, H)) * .5).astype(np.float32) if state else None
    dy = rng.standard_normal((T, B, H)).astype(np.float32)
    dhT = rng.standard_normal((B, H)).astype(np.float32)
    dcT = rng.standard_normal((B, H)).astype(np.float32)
    d = lambda a: None if a is None else a.astype(np.float64)
    cp64 = {k: v.astype(np.float64) for k, v in cp.items()}
    y64, hT64, cT64, saved = cn.forward(cp64, d(x), d(h0), d(c0))
    g64 = cn.backward(cp64, d(x), y64, saved, d(dy), d(dhT), d(dcT), d(h0), d(c0))
    tp = [torch.from_numpy(cp[k]).to(dev).requires_grad_(True) for k in names]
    xt = torch.from_numpy(x if not bf else np.ascontiguousarray(x.transpose(1, 0, 2))).to(dev).requires_grad_(True)
    h0t = None if h0 is None else torch.from_numpy(h0).to(dev).requires_grad_(True)
    c0t = None if c0 is None else torch.from_numpy(c0).to(dev).requires_grad_(True)
    y, hT, cT = vmlmf_sequence(xt, h0t, c0t, tp, batch_first=bf)
    dyt = torch.from_numpy(dy if not bf else np.ascontiguousarray(dy.transpose(1, 0, 2))).to(dev)
    torch.autograd.backward([y, hT, cT], [dyt, torch.from_numpy(dhT).to(dev), torch.from_numpy(dcT).to(dev)])
    torch.cuda.synchronize()
    un = (lambda a: a.transpose(1, 0, 2)) if bf else (lambda a: a)
    def err(a, b):
        a = np.asarray(a, dtype=np.float64)
        return max(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30), np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
    errs = {"y": err(un(y.detach().cpu().numpy()), y64), "hT": err(hT.detach().cpu().numpy(), hT64), "cT": err(cT.detach().cpu().numpy(), cT64),
            "dx": err(un(xt.grad.cpu().numpy()), g64["dx"])}
    for k, t in zip(names, tp):
        errs["d" + k] = err(t.grad.cpu().numpy(), g64[k])
    if state:
        errs["dh0"] = err(h0t.grad.cpu().numpy(), g64["dh0"])
        errs["dc0"] = err(c0t.grad.cpu().numpy(), g64["dc0"])
    w = max(errs.values())
    worst_all = max(worst_all, w)
    seen[path] = seen.get(path, 0) + 1
    flag = "" if w < 1e-5 else "   <-- ABOVE 1e-5: " + max(errs, key=errs.get)
    print(f"case {case:2d} path {path} T={T} B={B} I={I} H={H} RX={RX} RH={RH} bf={int(bf)} state={int(state)}  worst {w:.2e}{flag}", flush=True)
print("paths seen", seen, "worst", f"{worst_all:.2e}")
sys.exit(0 if worst_all < 1e-5 else 1)
