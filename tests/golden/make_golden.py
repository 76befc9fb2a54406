#!/usr/bin/env python
"""Generate golden vectors by executing the LIVE reference modules (build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (/root/reference, read-only) ships no golden vectors or known-answer tests for
the hot path (unittest/unit_test.py asserts shapes only), so parity is pinned by running the
reference's own nn.Modules on seeded weights and inputs, in fp32 (the reference dtype) and in
fp64 (error budget), and committing inputs, outputs and every gradient.  /root/reference does
not exist on the GPU box; tests only ever read the committed .npz files.

Key layout inside each .npz:  param/<state_dict key>, in/<name>, out/<name>, grad/<name>
(fp32 run) and out64/<name>, grad64/<name> (fp64 run of the same module on the same data).
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/rnn_compression_factorization_vmlmf/src"
sys.path.insert(0, REF)
from models.vmlmf import MyLSTM, MyLSTMCell, MyVMLMFCell, Net  # noqa: E402
from models.vmlmf_group import MyVMLMFCellg2, MyVMLMFgCellg2  # noqa: E402
from models.vmlmf_lm import Model, MyVMLSTM, MyVMLSTMGroup  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(build, make_inputs, run, seed):
    """build() -> module; make_inputs() -> dict of fp32 tensors (float ones get requires_grad);
    run(module, inputs) -> (dict outputs, scalar loss).  Executed in fp32 then fp64."""
    rec = {}
    for tag, dt in (("", torch.float32), ("64", torch.float64)):
        torch.set_default_dtype(torch.float32)
        torch.manual_seed(seed)
        mod = build().to(dt)             # same fp32-drawn weights in both runs
        torch.manual_seed(seed + 1)
        ins = make_inputs()
        torch.set_default_dtype(dt)      # so the reference's internal torch.zeros follow (SURVEY B-7)
        ins = {k: (v.to(dt).requires_grad_(True) if v.is_floating_point() else v) for k, v in ins.items()}
        outs, loss = run(mod, ins)
        loss.backward()
        if tag == "":
            for k, v in mod.state_dict().items():
                rec[f"param/{k}"] = v.detach().numpy().copy()
            for k, v in ins.items():
                rec[f"in/{k}"] = v.detach().numpy().copy()
        for k, v in outs.items():
            rec[f"out{tag}/{k}"] = v.detach().numpy().copy()
        for k, v in mod.named_parameters():
            if v.grad is not None:
                rec[f"grad{tag}/{k}"] = v.grad.numpy().copy()
        for k, v in ins.items():
            if v.is_floating_point() and v.grad is not None:
                rec[f"grad{tag}/in.{k}"] = v.grad.numpy().copy()
        torch.set_default_dtype(torch.float32)
    return rec


def _weighted(outs, ins):
    """scalar = sum_k <out_k, w_k> with the random upstream gradients stored in ins['w.<k>']."""
    return sum((outs[k] * ins[f"w.{k}"].detach()).sum() for k in outs)


def case_plain_cell():
    def run(m, i):
        h, c = m(i["x"], (i["h"], i["c"]))
        o = {"h": h, "c": c}
        return o, _weighted(o, i)
    return _run(lambda: MyVMLMFCell(9, 16, w_rank=3, u_ranks=2),
                lambda: dict(x=torch.randn(3, 9), h=torch.randn(3, 16) * .5, c=torch.randn(3, 16) * .5,
                             **{"w.h": torch.randn(3, 16), "w.c": torch.randn(3, 16)}), run, 11)


def _net_case(build, shape, seed, classes=18):
    def run(m, i):
        logits = m(i["x"])
        return {"logits": logits}, torch.nn.functional.cross_entropy(logits, i["label"])
    return _run(build, lambda: dict(x=torch.randn(*shape), label=torch.randint(0, classes, (shape[0],))), run, seed)


def case_net_plain():
    return _net_case(lambda: Net(9, [32], w_rank=8, u_rank=[6], cell=MyVMLMFCell), (4, 12, 9), 21, 6)


def case_net_opp_h180():
    # the reference unit-test fixture shape (unit_test.py:49-58) at a smaller batch
    return _net_case(lambda: Net(77, [180], w_rank=8, u_rank=[6], cell=MyVMLMFCell), (8, 24, 77), 31)


def case_net_group():
    return _net_case(lambda: Net(9, [16], w_rank=4, u_rank=[2, 3], cell=MyVMLMFCellg2), (5, 6, 9), 41)


def case_net_group_h180():
    return _net_case(lambda: Net(77, [180], w_rank=8, u_rank=[2, 4], cell=MyVMLMFCellg2), (4, 8, 77), 43)


def _stack_case(build, shape, hsum, seed):
    def run(m, i):
        seq, hcat = m(i["x"])
        o = {"seq": seq, "hcat": hcat}
        return o, _weighted(o, i)
    b, t, _ = shape
    return _run(build, lambda: dict(x=torch.randn(*shape), **{"w.seq": torch.randn(b, t, hsum[-1]),
                                                             "w.hcat": torch.randn(b, sum(hsum))}), run, seed)


def case_mylstm_2layer():
    return _stack_case(lambda: MyLSTM(9, [16, 24], w_rank=4, u_ranks=[3], cell=MyVMLMFCell), (3, 7, 9), [16, 24], 51)


def case_lstm_lowrank():
    # the reference's plain low-rank baseline cell (V/models/vmlmf.py:127-238), two layers
    return _stack_case(lambda: MyLSTM(9, [16, 24], w_rank=4, u_ranks=[3], cell=MyLSTMCell), (3, 7, 9), [16, 24], 53)


def case_lstm_dense():
    # ... and the uncompressed one (w_rank = u_ranks = None)
    return _stack_case(lambda: MyLSTM(9, [16], cell=MyLSTMCell), (3, 6, 9), [16], 55)


def case_group_ablation():
    return _stack_case(lambda: MyLSTM(9, [16], w_rank=4, u_ranks=[2, 3], cell=MyVMLMFgCellg2), (3, 5, 9), [16], 61)


def case_group_g4():
    # g=4 exercises multi-step rotation of the group index (vmlmf_group.py:123-124)
    def build():
        return MyLSTM(6, [16], w_rank=3, u_ranks=[2, 1, 3, 2], cell=MyVMLMFCellg2, g=4)
    return _stack_case(build, (3, 4, 6), [16], 63)


def _init_uniform(m, a):
    for p in m.parameters():
        torch.nn.init.uniform_(p, -a, a)
    return m


def case_lm_layer():
    def run(m, i):
        out, (h, c) = m(i["x"], (i["h0"], i["c0"]))
        o = {"out": out, "hT": h, "cT": c}
        return o, _weighted(o, i)
    return _run(lambda: _init_uniform(MyVMLSTM(24, 24, w_rank=5, u_ranks=7), 0.3),
                lambda: dict(x=torch.randn(6, 3, 24), h0=torch.randn(3, 24) * .5, c0=torch.randn(3, 24) * .5,
                             **{"w.out": torch.randn(6, 3, 24), "w.hT": torch.randn(3, 24),
                                "w.cT": torch.randn(3, 24)}), run, 71)


def case_lm_model():
    def nll(scores, y):   # lm_test.py:140-153 semantics, re-expressed
        b = y.size(1)
        p = torch.softmax(scores, 1)[torch.arange(y.numel()), y.reshape(-1)]
        return torch.mean(-torch.log(p) * b)

    def run(m, i):
        states = [(i["h0a"], i["c0a"]), (i["h0b"], i["c0b"])]
        scores, st = m(i["tok"], states)
        o = {"scores": scores, "hTa": st[0][0], "cTa": st[0][1], "hTb": st[1][0], "cTb": st[1][1]}
        return o, nll(scores, i["y"])
    return _run(lambda: Model(50, 16, 2, 0.0, 0.25, w_rank=4, u_ranks=[5], lstm_type="vmlmf"),
                lambda: dict(tok=torch.randint(0, 50, (5, 3)), y=torch.randint(0, 50, (5, 3)),
                             h0a=torch.randn(3, 16) * .3, c0a=torch.randn(3, 16) * .3,
                             h0b=torch.randn(3, 16) * .3, c0b=torch.randn(3, 16) * .3), run, 81)


def case_lm_group_b40():
    def run(m, i):
        out, (h, c) = m(i["x"], (i["h0"], i["c0"]))
        o = {"out": out, "hT": h, "cT": c}
        return o, _weighted(o, i)
    return _run(lambda: _init_uniform(MyVMLSTMGroup(16, 16, w_rank=4, u_ranks=[2, 3]), 0.3),
                lambda: dict(x=torch.randn(3, 40, 16), h0=torch.randn(40, 16) * .5, c0=torch.randn(40, 16) * .5,
                             **{"w.out": torch.randn(3, 40, 16), "w.hT": torch.randn(40, 16),
                                "w.cT": torch.randn(40, 16)}), run, 91)


CASES = {k[5:]: v for k, v in list(globals().items()) if k.startswith("case_")}

if __name__ == "__main__":
    for name, fn in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:      # optional: regenerate only the named cases
            continue
        rec = fn()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name:18s} {len(rec):3d} arrays  {os.path.getsize(path) / 1024:7.1f} KiB")
