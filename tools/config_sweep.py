#!/usr/bin/env python
"""Train-step and inference throughput of the five BASELINE.json configurations on one GPU (synthetic data, random
init, fp32) -- secondary numbers next to bench.py's headline line.  Prints one JSON object.
usage: python tools/config_sweep.py [out.json]"""
import json, os, sys
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


def har(name, I, H, wr, ur, cell, B, T, classes, n=20, graph=False):
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
    out = {"config": name, "batch": B, "seq_len": T, "train_ms": tr, "train_seq_per_s": B / tr * 1e3,
           "infer_ms": inf, "infer_seq_per_s": B / inf * 1e3}
    if graph:                                  # same step replayed as one CUDA graph (vmlmf_b200.graphs)
        from vmlmf_b200.graphs import GraphedTrainStep
        opt_g = torch.optim.Adam(net.parameters(), lr=0.002, capturable=True, foreach=True)
        step = GraphedTrainStep(net, opt_g, ce, x, y)
        trg = timeit(lambda: step(x, y), n)
        out.update({"train_graph_ms": trg, "train_graph_seq_per_s": B / trg * 1e3})
    return out


def lm(B, n=10, graph=False):
    """cfg4 train step as V/train_test/lm_test.py:196-209 runs it: carried + detached state, nll_loss, clip 5, SGD lr 1 --
    loss and update on the library's kernels (vb.nll_loss, vb.FlatClipSGD)."""
    torch.manual_seed(3)
    model = vb.Model(10000, 650, 2, 0.5, 0.05, 300, [300], "vmlmf").to(dev)
    x = torch.randint(0, 10000, (35, B), device=dev)
    y = torch.randint(0, 10000, (35, B), device=dev)
    sgd = vb.FlatClipSGD(model, lr=1.0, max_norm=5.0)
    static = model.state_init(B)

    def step():
        sgd.zero_grad()
        cur = [(h.detach(), c.detach()) for h, c in static]
        scores, new = model(x, cur)
        loss = vb.nll_loss(scores, y)
        loss.backward()
        sgd.step()
        with torch.no_grad():
            for (h, c), (h1, c1) in zip(static, new):
                h.copy_(h1)
                c.copy_(c1)
        return loss.detach()

    out = {"config": "cfg4 LM Model(10000,650,2,0.5,0.05,300,[300],'vmlmf') bptt 35", "batch": B, "seq_len": 35}
    tr = timeit(step, n)
    out.update({"train_ms": tr, "train_seq_per_s": B / tr * 1e3, "train_tokens_per_s": 35 * B / tr * 1e3})
    if graph:                                  # the same step as one CUDA graph (state lives in static buffers)
        from vmlmf_b200.graphs import GraphedCallable
        gstep = GraphedCallable(step)
        trg = timeit(gstep, n)
        out.update({"train_graph_ms": trg, "train_graph_seq_per_s": B / trg * 1e3, "train_graph_tokens_per_s": 35 * B / trg * 1e3})
    return out


out = {"gpu": torch.cuda.get_device_name(0), "results": []}
R = out["results"]
R.append(har("cfg1 UCI Net(9,[128],8,[6]) B=64 (reference batch)", 9, 128, 8, [6], vb.MyVMLMFCell, 64, 128, 6, graph=True))
R.append(har("cfg1 UCI shape, B=8192", 9, 128, 8, [6], vb.MyVMLMFCell, 8192, 128, 6))
R.append(har("cfg2 OPP Net(77,[256],8,[6]) B=81 (reference batch)", 77, 256, 8, [6], vb.MyVMLMFCell, 81, 24, 18, graph=True))
R.append(har("cfg2 OPP Net(77,[256],8,[6]) B=8192", 77, 256, 8, [6], vb.MyVMLMFCell, 8192, 24, 18))
R.append(har("cfg2 OPP Net(77,[256],32,[32]) B=8192 (generic regime)", 77, 256, 32, [32], vb.MyVMLMFCell, 8192, 24, 18, n=5))
R.append(har("cfg3 group Net(9,[128],8,[2,4],MyVMLMFCellg2) B=8192", 9, 128, 8, [2, 4], vb.MyVMLMFCellg2, 8192, 128, 6, n=10))
R.append(lm(20, graph=True))
R.append(lm(512))
R.append(har("cfg5 Net(9,[1024],64,[64]) B=2048 T=128 (generic regime)", 9, 1024, 64, [64], vb.MyVMLMFCell, 2048, 128, 6, n=3))
s = json.dumps(out, indent=1)
print(s)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(s)
