"""GPU: the callers either side of the recurrence (SURVEY 8 rows f1, f2 and the Net head) against their PyTorch
definitions -- F.cross_entropy (V/train_test/train.py:63), the LM nll_loss (lm_test.py:140-153, via the oracle),
nn.Linear, torch.optim.Adam (train.py:47) and clip_grad_norm_ + SGD (lm_test.py:203-209).  fp32, tolerance 1e-5
relative unless a test says otherwise; everything goes through the C ABI."""
import pytest
import torch

from conftest import rel_err
from oracle import vmlmf_oracle as vo

import vmlmf_b200 as vb
from vmlmf_b200.graphs import GraphedTrainStep

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def _close(a, b, tol=TOL, what=""):
    e = max(rel_err(a.detach().cpu().numpy(), b.detach().cpu().numpy()))
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol}"


@pytest.mark.parametrize("rows,C,pitch", [(37, 18, 18), (9472, 18, 18), (5, 3, 7), (1000, 2049, 2052), (70, 10000, 10000),
                                          (3, 10001, 10001)])
def test_softmax_nll_matches_cross_entropy(rows, C, pitch):
    torch.manual_seed(rows + C)
    buf = (torch.randn(rows, pitch, device=DEV) * 3.0)
    y = torch.randint(0, C, (rows,), device=DEV)
    up = torch.tensor(0.37, device=DEV)
    a = buf[:, :C].detach().requires_grad_(True)
    b = buf[:, :C].detach().clone().requires_grad_(True)
    la = vb.cross_entropy(a, y)
    lb = torch.nn.functional.cross_entropy(b.double(), y)
    (la * up).backward()
    (lb * up.double()).backward()
    _close(la, lb.float(), 2e-6, "loss")
    _close(a.grad, b.grad, TOL, "dlogits")
    again = vb.cross_entropy(a.detach(), y)
    assert torch.equal(again, la.detach()), "fixed-order sums must reproduce bit for bit"


def test_lm_nll_loss_matches_reference_definition():
    torch.manual_seed(9)
    T, B, V = 7, 5, 50
    s = torch.randn(T * B, V, device=DEV, requires_grad=True)
    y = torch.randint(0, V, (T, B), device=DEV)
    s_ref = s.detach().cpu().double().requires_grad_(True)
    la, lb = vb.nll_loss(s, y), vo.lm_nll_loss(s_ref, y.cpu())
    la.backward()
    lb.backward()
    _close(la, lb.float(), 2e-6, "loss")
    _close(s.grad, s_ref.grad.float(), TOL, "dscores")
    big = torch.full((4, 8), 200.0, device=DEV)          # exp(200) overflows fp32: the reference returns nan/inf here
    assert torch.isfinite(vb.nll_loss(big, torch.zeros(2, 2, dtype=torch.long, device=DEV)))


@pytest.mark.parametrize("B,K,N", [(37, 32, 18), (9472, 256, 18), (5, 180, 6), (300, 650, 32), (2, 16, 1), (1025, 1024, 9)])
def test_head_linear_matches_nn_linear(B, K, N):
    torch.manual_seed(B + K)
    wide = torch.randn(B, K + 8, device=DEV)
    h1 = wide[:, 8:].detach().requires_grad_(True)                  # strided rows, like h_last[:, -top:]
    h2 = wide[:, 8:].detach().clone().requires_grad_(True)
    lin = torch.nn.Linear(K, N).to(DEV)
    w2, b2 = lin.weight.detach().clone().requires_grad_(True), lin.bias.detach().clone().requires_grad_(True)
    up = torch.randn(B, N, device=DEV)
    o1 = vb.head_linear(h1, lin.weight, lin.bias)
    o2 = torch.nn.functional.linear(h2.double(), w2.double(), b2.double())
    (o1 * up).sum().backward()
    (o2 * up.double()).sum().backward()
    _close(o1, o2.float(), TOL, "out")
    _close(h1.grad, h2.grad, TOL, "dh")
    _close(lin.weight.grad, w2.grad, TOL, "dW")
    _close(lin.bias.grad, b2.grad, TOL, "db")


def _har(seed=3):
    torch.manual_seed(seed)
    net = vb.Net(9, [32], w_rank=4, u_rank=[3], cell=vb.MyVMLMFCell).to(DEV)
    x = torch.randn(48, 6, 9, device=DEV)
    y = torch.randint(0, 6, (48,), device=DEV)
    return net, x, y


def test_flat_adam_matches_torch_adam():
    net1, x, y = _har()
    net2, _, _ = _har()
    o1 = vb.FlatAdam(net1, lr=0.002)
    o2 = torch.optim.Adam(net2.parameters(), lr=0.002)
    for _ in range(6):
        o1.zero_grad()
        vb.cross_entropy(net1(x), y).backward()
        o1.step()
        o2.zero_grad()
        torch.nn.functional.cross_entropy(net2(x), y).backward()
        o2.step()
    for (k, p1), (_, p2) in zip(net1.named_parameters(), net2.named_parameters()):
        _close(p1, p2, 2e-5, k)            # six Adam steps amplify 1e-7 gradient differences where v is tiny
    assert not any(p.grad is not None for k, p in net1.named_parameters() if k.startswith("cell."))
    # parameters are views of one flat buffer; state_dict still has the reference's keys and shapes
    assert list(net1.state_dict().keys()) == list(net2.state_dict().keys())


def test_flat_adam_state_dict_resume():
    """Checkpoint / resume of the flat optimizer (state exported per parameter name): a run interrupted after three steps and
    resumed in a fresh optimizer continues bit-identically to the uninterrupted run."""
    def steps(net, opt, n):
        for _ in range(n):
            opt.zero_grad()
            vb.cross_entropy(net(x), y).backward()
            opt.step()
    net_a, x, y = _har()
    opt_a = vb.FlatAdam(net_a, lr=0.002)
    steps(net_a, opt_a, 3)
    weights = {k: v.clone() for k, v in net_a.state_dict().items()}
    state = opt_a.state_dict()
    assert state["step"] == 3 and set(state["state"]) == {k for k, p in net_a.named_parameters() if not k.startswith("cell.")}
    steps(net_a, opt_a, 2)                                   # uninterrupted: five steps
    net_b, _, _ = _har()
    opt_b = vb.FlatAdam(net_b, lr=0.5)                       # wrong hyper-parameters on purpose: the checkpoint restores them
    with pytest.raises(RuntimeError, match="before load_state_dict"):
        opt_b.load_state_dict(state)
    steps(net_b, opt_b, 1)                                   # builds the flat buffers (live set from the first backward)
    net_b.load_state_dict(weights)                           # copies into the flat-buffer views
    opt_b.load_state_dict(state)
    assert opt_b.lr == 0.002
    steps(net_b, opt_b, 2)
    for (k, pa), (_, pb) in zip(net_a.named_parameters(), net_b.named_parameters()):
        assert torch.equal(pa, pb), k


def test_flat_adam_inside_cuda_graph():
    net1, x, y = _har()
    net2, _, _ = _har()
    o1 = vb.FlatAdam(net1, lr=0.002)
    o2 = vb.FlatAdam(net2, lr=0.002)
    step = GraphedTrainStep(net1, o1, vb.cross_entropy, x, y, warmup=2, zero_fn=o1.zero_grad)      # 2 eager + capture
    for _ in range(3):
        step(x, y)
    for _ in range(2 + 3):                  # capture itself does not execute the step
        o2.zero_grad()
        vb.cross_entropy(net2(x), y).backward()
        o2.step()
    assert float(o1.t) == float(o2.t) == 5.0
    for (k, p1), (_, p2) in zip(net1.named_parameters(), net2.named_parameters()):
        _close(p1, p2, 2e-5, k)


def test_flat_clip_sgd_matches_clip_grad_norm_and_sgd():
    def build():
        torch.manual_seed(5)
        return vb.Model(50, 16, 2, 0.0, 0.25, w_rank=4, u_ranks=[5], lstm_type="vmlmf").to(DEV)
    m1, m2 = build(), build()
    tok = torch.randint(0, 50, (6, 4), device=DEV)
    y = torch.randint(0, 50, (6, 4), device=DEV)
    opt = vb.FlatClipSGD(m1, lr=1.0, max_norm=0.25)
    for it in range(3):
        opt.zero_grad()
        s1, _ = m1(tok, m1.state_init(4))
        vb.nll_loss(s1, y).backward()
        norm1 = opt.step()
        m2.zero_grad()
        s2, _ = m2(tok, m2.state_init(4))
        vb.nll_loss(s2, y).backward()
        with torch.no_grad():
            norm2 = torch.nn.utils.clip_grad_norm_(m2.parameters(), 0.25)
            for p in m2.parameters():
                p -= 1.0 * p.grad
        _close(norm1, norm2, TOL, f"norm step {it}")
        assert float(norm2) > 0.25, "the clip must be active for this test to mean anything"
    for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        _close(p1, p2, TOL, k)
        _close(p1.grad, p2.grad, TOL, k + ".grad")          # gradients rescaled in place like clip_grad_norm_


def test_graphed_step_with_static_inputs_reads_the_callers_buffers():
    """static_inputs=True: the tensors passed at construction are the graph's input buffers, so new data written into
    them (e.g. by the host-to-device copy of a double-buffered pipeline) is what the next replay trains on."""
    net1, x, y = _har()
    net2, _, _ = _har()
    xa, ya = x.clone(), y.clone()
    o1 = vb.FlatAdam(net1, lr=0.002)
    step = GraphedTrainStep(net1, o1, vb.cross_entropy, xa, ya, warmup=1, zero_fn=o1.zero_grad, static_inputs=True)
    assert step.static_x.data_ptr() == xa.data_ptr()
    x2 = torch.randn_like(x)
    xa.copy_(x2)                                   # new batch lands in the graph's own input buffer
    l1 = float(step(xa, ya))
    o2 = vb.FlatAdam(net2, lr=0.002)
    for data in (x, x2):                           # 1 eager warm-up step on x, then the replayed step on x2
        o2.zero_grad()
        l2 = vb.cross_entropy(net2(data), y)
        l2.backward()
        o2.step()
    assert abs(l1 - float(l2)) <= 1e-5 * max(1.0, abs(float(l2)))
    for (k, p1), (_, p2) in zip(net1.named_parameters(), net2.named_parameters()):
        _close(p1, p2, 2e-5, k)


@pytest.mark.parametrize("p", [0.0, 0.5])
def test_fused_embed_dropout_matches_indexing(p):
    """vmlmf_embed_dropout_fwd (Embed + dropout of the LM input, V/models/vmlmf_lm.py:48,:436): every element is either 0
    or w[token] / (1 - p), the view is backed by a pitch-padded buffer with zero pad columns, and the weight gradient is
    the dense embedding backward of the masked upstream gradient."""
    from vmlmf_b200.functional import embed_dropout
    torch.manual_seed(5)
    V, E, T, B = 37, 650, 6, 5
    w = (torch.randn(V, E, device=DEV) + 3.0).requires_grad_(True)          # no exact zeros: the mask is recoverable from the output
    tok = torch.randint(0, V, (T, B), device=DEV)
    out = embed_dropout(tok, w, p, training=True)
    assert out.shape == (T, B, E) and out.stride(-1) == 1 and out.stride(1) % 4 == 0
    dense = w.detach()[tok]
    keep = out.detach() != 0
    scale = 1.0 / (1.0 - p)
    _close(out.detach()[keep], dense[keep] * scale, 1e-6, "kept values")
    if p == 0.0:
        assert bool(keep.all())
    else:
        frac = keep.float().mean().item()
        assert abs(frac - (1 - p)) < 0.03, f"keep fraction {frac}"
    pad = out.detach().as_strided((T * B, out.stride(1)), (out.stride(1), 1))[:, E:]
    assert pad.numel() == 0 or bool((pad == 0).all())
    up = torch.randn(T, B, E, device=DEV)
    (out * up).sum().backward()
    ref = torch.zeros_like(w)
    ref.index_add_(0, tok.reshape(-1), (up * keep * scale).reshape(-1, E))
    _close(w.grad, ref, 1e-5, "dW")
    ev = embed_dropout(tok, w.detach(), p, training=False)
    assert torch.equal(ev, dense)                                            # eval: plain gather


def test_fast_tf32_mode():
    """Optional single-pass TF32 mode of the tcgen05 GEMMs (vb.set_fast_tf32): stated bound 2e-3 relative against an fp64
    product (1e-5 class in the default mode), and the decisions taken from the result (row-wise argmax of an LM-head sized
    product) do not change on margins above that bound."""
    from vmlmf_b200.functional import gemm_nt, gemm_tn
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(700, 650, device=DEV, generator=g)
    w = torch.randn(1000, 650, device=DEV, generator=g) * 0.05
    ref = x.double() @ w.double().t()
    try:
        vb.set_fast_tf32(False)
        acc = gemm_nt(x, w)
        vb.set_fast_tf32(True)
        fast = gemm_nt(x, w)
        fast_tn = gemm_tn(x, x[:, :96].contiguous())
    finally:
        vb.set_fast_tf32(False)
    e_acc = max(rel_err(acc.cpu().numpy(), ref.cpu().numpy()))
    e_fast = max(rel_err(fast.cpu().numpy(), ref.cpu().numpy()))
    e_tn = max(rel_err(fast_tn.cpu().numpy(), (x.double().t() @ x[:, :96].double()).cpu().numpy()))
    assert e_acc <= 3e-6, e_acc
    assert 1e-5 < e_fast <= 2e-3, f"single-pass TF32 error {e_fast:.2e} outside the stated window"
    assert e_tn <= 2e-3, e_tn
    top2 = ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 4e-3 * ref.abs().max()           # rows whose decision margin exceeds the error bound
    assert bool(clear.any())
    assert torch.equal(fast.argmax(1)[clear], ref.argmax(1)[clear])
    assert torch.equal(acc.argmax(1), ref.argmax(1))
