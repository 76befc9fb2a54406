"""CPU: host-side logic -- C ABI surface, plan, drop-in module layout, canonical packing."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, assert_close, load_golden
from oracle import canonical_numpy as cn

import vmlmf_b200 as vb
from vmlmf_b200 import _lib, packing


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "vmlmf_b200.h")).read()
    names = set(re.findall(r"\b(vmlmf_[a-z_0-9]+)\s*\(", hdr))
    assert {"vmlmf_seq_fwd", "vmlmf_seq_bwd", "vmlmf_xproj_fwd", "vmlmf_seq_plan"} <= names
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), n
    assert set(_lib.SIGNATURES) == names
    assert _lib.lib().vmlmf_abi_version() == _lib.ABI_VERSION


def test_plan_and_error_codes():
    p = _lib.plan(24, 81, 77, 256, 8, 6)
    assert p.path == _lib.PATH_R1 and p.zx_pitch == 8 and p.z_pitch == 8 and p.bwd_workspace_bytes > 0
    assert p.gates_bytes == 24 * 81 * 4 * 256 * 4 and p.cs_bytes == 24 * 81 * 256 * 4
    # large batches plan the warp-MMA path: saved state padded to whole 16-sequence x 16-unit fragments
    p = _lib.plan(24, 8200, 77, 180, 8, 6)
    assert p.path == _lib.PATH_R1M and p.zx_pitch == 8 and p.z_pitch == 8
    assert p.gates_bytes == 24 * 513 * 4 * 12 * 256 * 4 and p.cs_bytes * 4 == p.gates_bytes
    assert 0 < p.bwd_workspace_bytes < p.gates_bytes                       # fused backward: per-CTA partials only, no dPre
    p = _lib.plan(24, 8200, 77, 180, 8, 16)                                # more than 16 contraction rows: split backward
    assert p.path == _lib.PATH_R1M and p.bwd_workspace_bytes >= p.gates_bytes
    assert _lib.plan(24, 8200, 77, 181, 8, 6).path == _lib.PATH_R1          # H % 4 != 0 stays on the SIMT kernels
    assert _lib.plan(35, 4096, 650, 650, 300, 300).path in _lib.LARGE_PATHS   # R2 on a GPU box, G without a driver
    with pytest.raises(TypeError):                       # H < I: the reference raises TypeError too
        _lib.plan(4, 2, 20, 10, 4, 4)
    with pytest.raises(RuntimeError):
        _lib.plan(0, 2, 4, 8, 4, 4)
    assert b"hidden_size" in _lib.lib().vmlmf_strerror(-2)


def test_no_cpu_fallback():
    torch.manual_seed(0)
    net = vb.Net(9, [16], w_rank=4, u_rank=[3], cell=vb.MyVMLMFCell)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.randn(2, 3, 9))


def _same_seed_module(case, build, seed):
    """the drop-in, built under the reference fixture's seed, must draw the very same weights"""
    g = load_golden(case)
    torch.manual_seed(seed)
    m = build()
    sd = m.state_dict()
    ref = {k[len("param/"):]: v for k, v in g.items() if k.startswith("param/")}
    assert list(sd.keys()) == list(ref.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == ref[k].shape, k
        np.testing.assert_array_equal(v.numpy(), ref[k], err_msg=k)


def test_state_dict_layout_and_init_match_reference():
    _same_seed_module("net_plain", lambda: vb.Net(9, [32], w_rank=8, u_rank=[6], cell=vb.MyVMLMFCell), 21)
    _same_seed_module("net_group", lambda: vb.Net(9, [16], w_rank=4, u_rank=[2, 3], cell=vb.MyVMLMFCellg2), 41)
    _same_seed_module("mylstm_2layer", lambda: vb.MyLSTM(9, [16, 24], w_rank=4, u_ranks=[3], cell=vb.MyVMLMFCell), 51)
    _same_seed_module("lstm_lowrank", lambda: vb.MyLSTM(9, [16, 24], w_rank=4, u_ranks=[3], cell=vb.MyLSTMCell), 53)
    _same_seed_module("lstm_dense", lambda: vb.MyLSTM(9, [16], cell=vb.MyLSTMCell), 55)
    _same_seed_module("group_ablation", lambda: vb.MyLSTM(9, [16], w_rank=4, u_ranks=[2, 3], cell=vb.MyVMLMFgCellg2), 61)
    _same_seed_module("group_g4", lambda: vb.MyLSTM(6, [16], w_rank=3, u_ranks=[2, 1, 3, 2], cell=vb.MyVMLMFCellg2, g=4), 63)
    _same_seed_module("lm_model", lambda: vb.Model(50, 16, 2, 0.0, 0.25, w_rank=4, u_ranks=[5], lstm_type="vmlmf"), 81)


def test_reference_unit_test_shapes():
    """the reference's six shape tests (V/unittest/unit_test.py:63-93), minus the forward calls"""
    c = vb.Net(77, layer_sizes=[180], w_rank=8, u_rank=[6], model=vb.MyLSTM, cell=vb.MyVMLMFCell)
    g = vb.Net(77, layer_sizes=[180], w_rank=8, u_rank=[2, 4], model=vb.MyLSTM, cell=vb.MyVMLMFCellg2)
    assert c.cell.dia_x.shape == (1, 77) and c.cell.dia_h.shape == (1, 180)
    assert c.cell.u_x.shape == (77, 8) and c.cell.u_h.shape == (180, 6)
    assert c.cell.v_x.shape == (720, 8) and c.cell.v_h.shape == (720, 6)
    L = g.cell.layers
    assert L["dia_x"].shape == (1, 77) and L["dia_h"].shape == (1, 180)
    assert L["u_h_0"].shape == (2, 90, 2) and L["u_h_1"].shape == (2, 90, 4)
    assert L["v_h_0"].shape == (2, 2, 360) and L["v_h_1"].shape == (2, 4, 360)


def _tparams(g, prefix):
    return {k[len("param/") + len(prefix):]: torch.from_numpy(v).double().requires_grad_(True)
            for k, v in g.items() if k.startswith("param/" + prefix)}


NAMES = ("Ux", "Vx", "Dx", "A", "Bm", "Dh", "bias")


def test_pack_plain_matches_numpy_spec_and_chain_rule():
    g = load_golden("lm_layer")
    p = _tparams(g, "")
    canon = packing.pack_plain(p["u_x"], p["u_h"], p["w_x"], p["w_h"], p["b_x"], p["b_h"], p["dia_x"], p["dia_h"])
    ref = cn.pack_plain({k: v.detach() for k, v in p.items()}, v_names=("w_x", "w_h"))
    rng = np.random.default_rng(0)
    gc = {}
    for n, t in zip(NAMES, canon):
        assert_close(t.detach().numpy(), ref[n], 1e-12, n)
        gc[n] = rng.standard_normal(ref[n].shape)
    sum((t * torch.from_numpy(gc[n])).sum() for n, t in zip(NAMES, canon)).backward()
    back = cn.unpack_grads_plain({k: v.detach() for k, v in p.items()}, gc, v_names=("w_x", "w_h"))
    for k, v in back.items():
        assert_close(p[k].grad.numpy().reshape(v.shape), v, 1e-12, k)


@pytest.mark.parametrize("case,gsz,vm", [("group_g4", 4, True), ("group_ablation", 2, False), ("net_group_h180", 2, True)])
def test_pack_group_matches_numpy_spec_and_chain_rule(case, gsz, vm):
    g = load_golden(case)
    prefix = "rnn.rnncells.0.layers." if case.startswith("net") else "rnncells.0.layers."
    p = _tparams(g, prefix)
    canon = packing.pack_group(p, gsz, with_vm=vm)
    pd = {k: v.detach() for k, v in p.items()}
    ref = cn.pack_group(pd, g=gsz, with_vm=vm)
    rng = np.random.default_rng(1)
    gc = {}
    for n, t in zip(NAMES, canon):
        assert_close(t.detach().numpy(), ref[n], 1e-12, n)
        gc[n] = rng.standard_normal(ref[n].shape)
    # structural zeros of A/Bm carry no gradient back to the parameters; mask the random upstream
    gc["A"] = gc["A"] * (ref["A"] != 0)
    gc["Bm"] = gc["Bm"] * (ref["Bm"] != 0)
    if not vm:
        gc["Dx"] *= 0
        gc["Dh"] *= 0
    sum((t * torch.from_numpy(gc[n])).sum() for n, t in zip(NAMES, canon)).backward()
    back = cn.unpack_grads_group(pd, gc, g=gsz, with_vm=vm)
    for k, v in back.items():
        assert_close(p[k].grad.numpy().reshape(v.shape), v, 1e-12, k)


def test_pack_lm_group_reproduces_reference_layer_fp64():
    """MyVMLSTMGroup (V/models/vmlmf_lm.py:97-160, arithmetic as shipped) == the canonical recurrence with
    packing.pack_lm_group's parameters: outputs against the live reference's fp64 run, and the parameter gradients
    through the packing's autograd chain rule + the numpy BPTT against the reference's fp64 autograd."""
    g = load_golden("lm_group_b40")
    p = {k[len("param/"):]: torch.from_numpy(v.astype(np.float64)).requires_grad_(True)
         for k, v in g.items() if k.startswith("param/")}
    canon = packing.pack_lm_group(p["u_x"], p["w_x"], [p["u_h.0"], p["u_h.1"]], [p["v_h.0"], p["v_h.1"]], p["b_x"],
                                  p["b_h"], p["dia_x"], p["dia_h"], 2)
    cp = {n: t.detach().numpy() for n, t in zip(NAMES, canon)}
    x, h0, c0 = (g[f"in/{k}"].astype(np.float64) for k in ("x", "h0", "c0"))
    y, hT, cT, saved = cn.forward(cp, x, h0, c0)
    assert_close(y, g["out64/out"], 1e-11, "out")
    assert_close(hT, g["out64/hT"], 1e-11, "hT")
    assert_close(cT, g["out64/cT"], 1e-11, "cT")
    gr = cn.backward(cp, x, y, saved, g["in/w.out"].astype(np.float64), g["in/w.hT"].astype(np.float64),
                     g["in/w.cT"].astype(np.float64), h0, c0)
    sum((t * torch.from_numpy(gr[n])).sum() for n, t in zip(NAMES, canon)).backward()
    for k, t in p.items():
        assert_close(t.grad.numpy(), g[f"grad64/{k}"], 1e-10, k)
    assert_close(gr["dx"], g["grad64/in.x"], 1e-10, "dx")


def test_lm_group_layer_constructor_and_dispatch():
    torch.manual_seed(91)
    m = vb.MyVMLSTMGroup(16, 16, w_rank=4, u_ranks=[2, 3])
    g = load_golden("lm_group_b40")
    ref = {k[len("param/"):]: v.shape for k, v in g.items() if k.startswith("param/")}
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == ref
    assert isinstance(vb.Model(50, 16, 1, 0.0, 0.1, w_rank=4, u_ranks=[2, 3], lstm_type="vmgroup").rnns[0], vb.MyVMLSTMGroup)
    assert isinstance(vb.Model(50, 16, 1, 0.0, 0.1, w_rank=4, u_ranks=[2, 3], lstm_type="vm_group").rnns[0], torch.nn.LSTM)


def test_grad_bucket_pack_zero_semantics_single_process():
    """GradBucket: live parameters found from the first backward, gradients gathered with one multi-tensor copy, `.grad`
    re-pointed at the bucket slices, `zero()` drops them so the next backward needs no accumulation kernels."""
    from vmlmf_b200.parallel import GradBucket
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    dead = torch.nn.Parameter(torch.randn(4))
    mod = torch.nn.Module()
    mod.lin, mod.dead = lin, dead
    bucket = GradBucket(mod)
    x = torch.randn(7, 5)
    for it in range(3):
        bucket.zero()
        assert all(p.grad is None for p in mod.parameters())
        lin(x).pow(2).sum().backward()
        ref = [p.grad.clone() for p in lin.parameters()]
        flat = bucket.all_reduce()                       # no process group: packs only
        assert flat.numel() == 3 * 5 + 3 and dead.grad is None
        off = 0
        for p, r in zip(lin.parameters(), ref):
            assert p.grad.data_ptr() == flat.data_ptr() + 4 * off          # .grad is the bucket slice
            assert torch.equal(p.grad, r)
            off += p.numel()
    # accumulation across two backwards before a step still works (the second one adds into the slices)
    lin(x).pow(2).sum().backward()
    assert torch.allclose(bucket.pack()[:15].view(3, 5), 2 * ref[0])


def test_grad_bucket_reports_empty_and_late_parameters():
    """GradBucket fixes the live parameter set at the first backward: both misuse cases raise instead of silently skipping
    (advisor finding, round 1)."""
    import torch
    from vmlmf_b200.parallel import GradBucket
    lin = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 2))
    b = GradBucket(lin)
    with pytest.raises(RuntimeError, match="no parameter has a gradient"):
        b.pack()
    for p in lin[1].parameters():
        p.requires_grad_(False)
    lin(torch.randn(5, 3)).sum().backward()
    b.pack()                                              # live set = first layer only
    assert len(b.params) == 2
    b.zero()
    for p in lin[1].parameters():
        p.requires_grad_(True)
    lin(torch.randn(5, 3)).sum().backward()
    with pytest.raises(RuntimeError, match="was not live at the first backward"):
        b.pack()


def test_header_and_binding_agree_on_regime_codes():
    """include/vmlmf_b200.h (the boundary a maintainer binds) and the ctypes binding name the same regimes and ABI version."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "vmlmf_b200.h")).read()
    codes = {m.group(1): int(m.group(2)) for m in re.finditer(r"VMLMF_PATH_(\w+)\s*=\s*(\d+)", hdr)}
    assert codes == {"R1": _lib.PATH_R1, "G": _lib.PATH_G, "R1M": _lib.PATH_R1M, "R2": _lib.PATH_R2, "R3": _lib.PATH_R3}
    assert int(re.search(r"#define VMLMF_ABI_VERSION (\d+)", hdr).group(1)) == _lib.ABI_VERSION
    assert set(_lib.LARGE_PATHS) == {_lib.PATH_G, _lib.PATH_R2, _lib.PATH_R3}
