#!/usr/bin/env python
"""Train-step and inference throughput of the five BASELINE.json configurations on one GPU (synthetic data, random
init, fp32) -- secondary numbers next to bench.py's headline line.  Prints one JSON object.
usage: python tools/config_sweep.py [out.json]"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vmlmf_b200 as vb

dev = "cuda:0"
ce = torch.nn.functional.cross_entropy


def timeit(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def har(name, I, H, wr, ur, cell, B, T, classes, n=20):
    torch.manual_seed(3)
    net = vb.Net(I, [H], w_rank=wr, u_rank=ur, cell=cell).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=0.002, fused=True)
    x = torch.randn(B, T, I, device=dev)
    y = torch.randint(0, classes, (B,), device=dev)

    def train():
        net.zero_grad()
        ce(net(x), y).backward()
        opt.step()

    def infer():
        with torch.no_grad():
            net(x)

    tr, inf = timeit(train, n), timeit(infer, n)
    return {"config": name, "batch": B, "seq_len": T, "train_ms": tr, "train_seq_per_s": B / tr * 1e3,
            "infer_ms": inf, "infer_seq_per_s": B / inf * 1e3}


def lm(B, n=10):
    torch.manual_seed(3)
    model = vb.Model(10000, 650, 2, 0.5, 0.05, 300, [300], "vmlmf").to(dev)
    x = torch.randint(0, 10000, (35, B), device=dev)
    y = torch.randint(0, 10000, (35 * B,), device=dev)
    st = [model.state_init(B)]

    def train():
        model.zero_grad()
        st[0] = model.detach(st[0])
        scores, st[0] = model(x, st[0])
        ce(scores, y).backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5)
        with torch.no_grad():
            for p in model.parameters():
                p -= 1.0 * p.grad

    tr = timeit(train, n)
    return {"config": "cfg4 LM Model(10000,650,2,0.5,0.05,300,[300],'vmlmf') bptt 35", "batch": B, "seq_len": 35,
            "train_ms": tr, "train_seq_per_s": B / tr * 1e3, "train_tokens_per_s": 35 * B / tr * 1e3}


out = {"gpu": torch.cuda.get_device_name(0), "results": []}
R = out["results"]
R.append(har("cfg1 UCI Net(9,[128],8,[6]) B=64 (reference batch)", 9, 128, 8, [6], vb.MyVMLMFCell, 64, 128, 6))
R.append(har("cfg1 UCI shape, B=8192", 9, 128, 8, [6], vb.MyVMLMFCell, 8192, 128, 6))
R.append(har("cfg2 OPP Net(77,[256],8,[6]) B=81 (reference batch)", 77, 256, 8, [6], vb.MyVMLMFCell, 81, 24, 18))
R.append(har("cfg2 OPP Net(77,[256],8,[6]) B=8192", 77, 256, 8, [6], vb.MyVMLMFCell, 8192, 24, 18))
R.append(har("cfg2 OPP Net(77,[256],32,[32]) B=8192 (generic regime)", 77, 256, 32, [32], vb.MyVMLMFCell, 8192, 24, 18, n=5))
R.append(har("cfg3 group Net(9,[128],8,[2,4],MyVMLMFCellg2) B=8192", 9, 128, 8, [2, 4], vb.MyVMLMFCellg2, 8192, 128, 6, n=10))
R.append(lm(20))
R.append(lm(512))
R.append(har("cfg5 Net(9,[1024],64,[64]) B=2048 T=128 (generic regime)", 9, 1024, 64, [64], vb.MyVMLMFCell, 2048, 128, 6, n=3))
s = json.dumps(out, indent=1)
print(s)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(s)
