"""CPU: pin the oracle (oracle/) against golden vectors produced by the live reference.

fp32 tolerance 2e-6 (the reference's own fp32-vs-fp64 noise floor is ~1e-7..7e-7);
fp64 tolerance 1e-11.
"""
import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import canonical_numpy as cn
from oracle import vmlmf_oracle as vo

TOL = {torch.float32: 2e-6, torch.float64: 1e-11}
TAG = {torch.float32: "", torch.float64: "64"}


def _t(a, dt, grad=False):
    t = torch.from_numpy(np.array(a))
    if t.is_floating_point():
        t = t.to(dt)
        if grad:
            t.requires_grad_(True)
    return t


def _params(g, prefix, dt):
    return {k[len("param/") + len(prefix):]: _t(v, dt, True) for k, v in g.items()
            if k.startswith("param/" + prefix)}


def _check(g, dt, outs, grads):
    for k, v in outs.items():
        assert_close(v.detach().numpy(), g[f"out{TAG[dt]}/{k}"], TOL[dt], f"out {k}")
    n = 0
    for k, v in grads.items():
        key = f"grad{TAG[dt]}/{k}"
        assert key in g, key
        assert v.grad is not None, k
        assert_close(v.grad.numpy(), g[key], TOL[dt] * 5, f"grad {k}")
        n += 1
    assert n > 0


def _weighted(outs, g, dt):
    return sum((outs[k] * _t(g[f"in/w.{k}"], dt)).sum() for k in outs)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_plain_cell_step(dt):
    g = load_golden("plain_cell")
    p = _params(g, "", dt)
    ins = {k: _t(g[f"in/{k}"], dt, True) for k in ("x", "h", "c")}
    h, c = vo.plain_cell_step(p, ins["x"], ins["h"], ins["c"])
    outs = {"h": h, "c": c}
    _weighted(outs, g, dt).backward()
    _check(g, dt, outs, {**p, **{f"in.{k}": v for k, v in ins.items()}})


def test_plain_cell_rejects_hidden_smaller_than_input():
    p = {k: torch.zeros(s) for k, s in dict(u_x=(5, 2), u_h=(3, 2), v_x=(12, 2), v_h=(12, 2), b_x=(12,),
                                            b_h=(12,), dia_x=(1, 5), dia_h=(1, 3)).items()}
    with pytest.raises(TypeError):
        vo.plain_cell_step(p, torch.zeros(2, 5), torch.zeros(2, 3), torch.zeros(2, 3))


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("case,kind,prefix", [("net_plain", "plain", "rnn.rnncells.0."),
                                              ("net_opp_h180", "plain", "rnn.rnncells.0."),
                                              ("net_group", "group", "rnn.rnncells.0.layers."),
                                              ("net_group_h180", "group", "rnn.rnncells.0.layers.")])
def test_net(case, kind, prefix, dt):
    g = load_golden(case)
    p = _params(g, prefix, dt)
    lw, lb = _t(g["param/lin.weight"], dt, True), _t(g["param/lin.bias"], dt, True)
    x = _t(g["in/x"], dt, True)
    logits = vo.net_forward([p], lw, lb, x, kind=kind)
    torch.nn.functional.cross_entropy(logits, _t(g["in/label"], dt)).backward()
    grads = {prefix + k: v for k, v in p.items()}
    grads.update({"lin.weight": lw, "lin.bias": lb, "in.x": x})
    _check(g, dt, {"logits": logits}, grads)
    # the dead Net.cell parameters never receive a gradient in the reference (SURVEY B-2)
    assert not any(k.startswith("grad/cell.") for k in g)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("case,kind,nl,kw", [("mylstm_2layer", "plain", 2, {}),
                                             ("group_ablation", "group_novm", 1, {}),
                                             ("group_g4", "group", 1, {"g": 4}),
                                             ("lstm_lowrank", "lstm", 2, {}), ("lstm_dense", "lstm", 1, {})])
def test_layer_stack(case, kind, nl, kw, dt):
    g = load_golden(case)
    sub = "" if kind in ("plain", "lstm") else "layers."
    cells = [_params(g, f"rnncells.{l}.{sub}", dt) for l in range(nl)]
    x = _t(g["in/x"], dt, True)
    seq, hcat = vo.layer_stack(cells, x, kind=kind, **kw)
    outs = {"seq": seq, "hcat": hcat}
    _weighted(outs, g, dt).backward()
    grads = {f"rnncells.{l}.{sub}{k}": v for l, c in enumerate(cells) for k, v in c.items()}
    grads["in.x"] = x
    _check(g, dt, outs, grads)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("case", ["lm_layer", "lm_group_b40"])
def test_lm_layer(case, dt):
    g = load_golden(case)
    p = _params(g, "", dt)
    x, h0, c0 = (_t(g[f"in/{k}"], dt, True) for k in ("x", "h0", "c0"))
    if case == "lm_layer":
        out, (h, c) = vo.lm_layer(p, x, (h0, c0))
    else:
        h, c, seq = h0, c0, []
        for x_t in x.unbind(0):
            h, c = vo.lm_group_cell_step(p, x_t, h, c)
            seq.append(h)
        out = torch.stack(seq)
    outs = {"out": out, "hT": h, "cT": c}
    _weighted(outs, g, dt).backward()
    _check(g, dt, outs, {**p, "in.x": x, "in.h0": h0, "in.c0": c0})


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_lm_model(dt):
    g = load_golden("lm_model")
    layers = [_params(g, f"rnns.{l}.", dt) for l in range(2)]
    ew, fw, fb = (_t(g[f"param/{k}"], dt, True) for k in ("embed.w", "fc.w", "fc.b"))
    st = [tuple(_t(g[f"in/{n}{s}"], dt, True) for n in ("h0", "c0")) for s in "ab"]
    scores, new = vo.lm_model_forward(ew, layers, fw, fb, _t(g["in/tok"], dt), st)
    vo.lm_nll_loss(scores, _t(g["in/y"], dt)).backward()
    outs = {"scores": scores, "hTa": new[0][0], "cTa": new[0][1], "hTb": new[1][0], "cTb": new[1][1]}
    grads = {f"rnns.{l}.{k}": v for l, c in enumerate(layers) for k, v in c.items()}
    grads.update({"embed.w": ew, "fc.w": fw, "fc.b": fb})
    _check(g, dt, outs, grads)


# ---------------- canonical numpy restatement (the kernels' spec) ---------------- #

def _np_params(g, prefix):
    return {k[len("param/") + len(prefix):]: v for k, v in g.items() if k.startswith("param/" + prefix)}


def _canon_roundtrip(cp, x_tm, dy_tm, dhT, dcT, h0=None, c0=None):
    y, hT, cT, saved = cn.forward(cp, x_tm, h0, c0)
    grads = cn.backward(cp, x_tm, y, saved, dy_tm, dhT, dcT, h0, c0)
    return y, hT, cT, grads


def test_canonical_plain_matches_reference_fp64():
    g = load_golden("lm_layer")
    p = _np_params(g, "")
    cp = cn.pack_plain(p, v_names=("w_x", "w_h"))
    x, h0, c0 = (g[f"in/{k}"].astype(np.float64) for k in ("x", "h0", "c0"))
    y, hT, cT, gr = _canon_roundtrip(cp, x, g["in/w.out"].astype(np.float64), g["in/w.hT"].astype(np.float64),
                                     g["in/w.cT"].astype(np.float64), h0, c0)
    assert_close(y, g["out64/out"], 1e-11, "out")
    assert_close(hT, g["out64/hT"], 1e-11, "hT")
    assert_close(cT, g["out64/cT"], 1e-11, "cT")
    ref = cn.unpack_grads_plain(p, gr, v_names=("w_x", "w_h"))
    for k, v in ref.items():
        assert_close(v, g[f"grad64/{k}"], 1e-10, k)
    assert_close(gr["dx"], g["grad64/in.x"], 1e-10, "dx")
    assert_close(gr["dh0"], g["grad64/in.h0"], 1e-10, "dh0")
    assert_close(gr["dc0"], g["grad64/in.c0"], 1e-10, "dc0")


@pytest.mark.parametrize("case,gsz,vm", [("group_g4", 4, True), ("group_ablation", 2, False)])
def test_canonical_group_matches_reference_fp64(case, gsz, vm):
    g = load_golden(case)
    p = _np_params(g, "rnncells.0.layers.")
    cp = cn.pack_group(p, g=gsz, with_vm=vm)
    x = g["in/x"].astype(np.float64).transpose(1, 0, 2).copy()          # batch-first -> time-major
    dy = g["in/w.seq"].astype(np.float64).transpose(1, 0, 2).copy()
    y, hT, _, gr = _canon_roundtrip(cp, x, dy, g["in/w.hcat"].astype(np.float64), None)
    assert_close(y.transpose(1, 0, 2), g["out64/seq"], 1e-11, "seq")
    assert_close(hT, g["out64/hcat"], 1e-11, "hcat")
    ref = cn.unpack_grads_group(p, gr, g=gsz, with_vm=vm)
    for k, v in ref.items():
        assert_close(v, g[f"grad64/rnncells.0.layers.{k}"], 1e-10, k)
    assert_close(gr["dx"].transpose(1, 0, 2), g["grad64/in.x"], 1e-10, "dx")
