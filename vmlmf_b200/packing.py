"""Reference parameters -> canonical parameters (differentiable torch ops, once per forward).

Canonical set (gate order i,f,o,n; see include/vmlmf_b200.h):
    Ux[I,RX] Vx[4H,RX] Dx[4,I]   A[H,RH] Bm[4H,RH] Dh[4,H]   bias[4H]
The reference re-derives the diagonal corrections inside every timestep
(V/models/vmlmf.py:102-106, vmlmf_group.py:104-110, vmlmf_lm.py:251-255); they are loop invariant,
so they are hoisted here and autograd supplies the matching chain rule on the way back.
"""
from __future__ import annotations

import torch

# chunk position used by the group cells for canonical gate k=(i,f,o,n): they chunk as (f,i,n,o)
# (V/models/vmlmf_group.py:134,142,211,232)
GROUP_Q_OF_K = (1, 0, 3, 2)


def _diag_corr(u, v, n):
    """sum_r u[j,r] v[kH+j,r] for k<4, j<n  -> [4,n]"""
    hidden = v.shape[0] // 4
    return (u.unsqueeze(0) * v.view(4, hidden, -1)[:, :n]).sum(-1)


def pack_plain(u_x, u_h, v_x, v_h, b_x, b_h, dia_x, dia_h):
    """MyVMLMFCell (V/models/vmlmf.py:78-125) and MyVMLSTM (vmlmf_lm.py:222-269, v=w_x/w_h)."""
    n_in, hidden = u_x.shape[0], u_h.shape[0]
    if u_x.is_cuda:                      # fused: one kernel per side each way (host-side tests use the torch form below)
        from .functional import diag_correction
        dx, dh = diag_correction(u_x, v_x, dia_x), diag_correction(u_h, v_h, dia_h)
    else:
        dx = dia_x.reshape(1, n_in) - _diag_corr(u_x, v_x, n_in)
        dh = dia_h.reshape(1, hidden) - _diag_corr(u_h, v_h, hidden)
    # b_h passes through a multiply so that autograd hands b_x and b_h two DIFFERENT gradient tensors: a bare
    # `b_x + b_h` gives both AccumulateGrad nodes the same tensor, they alias .grad, and any in-place accumulation
    # into persistent .grad buffers (zero_grad(set_to_none=False), CUDA-graph training) then doubles the bias gradient
    return u_x, v_x, dx, u_h, v_h, dh, b_x.reshape(-1) + b_h.reshape(-1) * 1.0


def _reorder_gates(t, dim):
    """chunks of `dim` (size 4) taken in GROUP_Q_OF_K order, without an index tensor (an index list would be built on the
    host and copied to the device on every call, which also breaks CUDA-graph capture)"""
    return torch.stack([t.select(dim, q) for q in GROUP_Q_OF_K], dim)


def _gate_perm(hidden, device):
    return torch.cat([torch.arange(q * hidden, (q + 1) * hidden, device=device) for q in GROUP_Q_OF_K])


def pack_group(layers, g, with_vm=True):
    """MyVMLMFCellg2 / MyVMLMFgCellg2 (V/models/vmlmf_group.py:85-155, :203-251).

    The hidden-side map is g x g blocks per gate; block (src (j+off)%g -> dst j) has rank r_off.
    It is packed densely into A[H,R], Bm[4H,R], R = g * sum(r_off), zero outside the blocks."""
    u_x, v_x = layers["u_x"], layers["v_x"]
    n_in, hidden = u_x.shape[0], v_x.shape[0] // 4
    hg = hidden // g
    a_cols, b_cols = [], []
    for off in range(g):
        u, v = layers[f"u_h_{off}"], layers[f"v_h_{off}"]          # [g,Hg,r], [g,r,4Hg]
        r = u.shape[2]
        # A: rows of source group s=(j+off)%g, one column block per destination group j
        blk = u.new_zeros(g, hg, g, r)                             # [s, m, j, r]
        for j in range(g):
            blk[(j + off) % g, :, j, :] = u[j]
        a_cols.append(blk.reshape(hidden, g * r))
        # Bm: rows (k, j, m), columns (j', r) non-zero for j'==j
        vq = _reorder_gates(v.view(g, r, 4, hg), 2)                 # [j, r, k, m]
        bb = u.new_zeros(4, g, hg, g, r)                           # [k, j, m, j', r]
        for j in range(g):
            bb[:, j, :, j, :] = vq[j].permute(1, 2, 0)             # [k, m, r]
        b_cols.append(bb.reshape(4 * hidden, g * r))
    a = torch.cat(a_cols, 1)
    bm = torch.cat(b_cols, 1)
    perm = _gate_perm(hidden, u_x.device)
    bx, bh = layers["bias_x"].reshape(-1), layers["bias_h"].reshape(-1)
    if with_vm:
        dx = layers["dia_x"].reshape(1, n_in) - _diag_corr(u_x, v_x, n_in)
        u0, v0 = layers["u_h_0"], layers["v_h_0"]
        v0q = _reorder_gates(v0.view(g, v0.shape[1], 4, hg), 2)                # [j, r, k, m]
        corr = torch.einsum("jmr,jrkm->kjm", u0, v0q).reshape(4, hidden)   # diag of offset-0 blocks (:101-110)
        dh = layers["dia_h"].reshape(1, hidden) - corr
        return u_x, v_x, dx, a, bm, dh, bx + bh[perm]
    zx = u_x.new_zeros(4, n_in)
    zh = u_x.new_zeros(4, hidden)
    return u_x, v_x[perm], zx, a, bm, zh, bx[perm] + bh[perm]


def pack_lm_group(u_x, w_x, u_h, v_h, b_x, b_h, dia_x, dia_h, g):
    """MyVMLSTMGroup.lstm_step (V/models/vmlmf_lm.py:97-160), arithmetic kept as shipped.

    u_h[off]: [g, Hg, r_off], v_h[off]: [g, r_off, 4Hg].  The bmm result [B, g, 4Hg] is flattened GROUP-major to
    [B, 4H] (:135) before chunk(4) (:155), so canonical row kH+p of Bm is flat column c = kH+p = j*4Hg + m: group j's
    output m -- with g = 2, gates i,f read group 0 and o,n group 1.  The "diagonal" that is subtracted (:141-148)
    pairs re_uh[p] = u_h[0] viewed [H, r0] with re_vh[kH+p] = v_h[0] transposed and viewed [4H, r0]; it multiplies
    h[p] element-wise, which is exactly a Dh coefficient, whether or not it is the true diagonal (SURVEY B-5)."""
    n_in, hidden = u_x.shape[0], w_x.shape[0] // 4
    hg = hidden // g
    a_cols, b_cols = [], []
    for off in range(g):
        u, v = u_h[off], v_h[off]
        r = u.shape[2]
        blk = u.new_zeros(g, hg, g, r)                             # [source group s, m', position j, r]
        for j in range(g):
            blk[(j + off) % g, :, j, :] = u[j]
        a_cols.append(blk.reshape(hidden, g * r))
        bb = u.new_zeros(g, 4 * hg, g, r)                          # [j, m, j', r], flat row j*4Hg + m
        for j in range(g):
            bb[j, :, j, :] = v[j].t()
        b_cols.append(bb.reshape(4 * hidden, g * r))
    a, bm = torch.cat(a_cols, 1), torch.cat(b_cols, 1)
    r0 = u_h[0].shape[2]
    re_u = u_h[0].reshape(hidden, r0)
    re_v = v_h[0].transpose(1, 2).reshape(4 * hidden, r0)
    dh = dia_h.reshape(1, hidden) - (re_u.unsqueeze(0) * re_v.view(4, hidden, r0)).sum(-1)
    dx = dia_x.reshape(1, n_in) - _diag_corr(u_x, w_x, n_in)
    return u_x, w_x, dx, a, bm, dh, b_x.reshape(-1) + b_h.reshape(-1) * 1.0
