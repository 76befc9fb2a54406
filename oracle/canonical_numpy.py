"""numpy restatement of the canonical VMLMF recurrence and its hand-derived BPTT.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): never imported by the product.

All four reference cells reduce to one recurrence over *canonical* parameters

    pre_t[b,k,j] = sum_r (x_t Ux)[b,r] Vx[kH+j,r] + [j<I] x_t[b,j] Dx[k,j]
                 + sum_r (h_{t-1} A)[b,r] Bm[kH+j,r] +   h_{t-1}[b,j] Dh[k,j] + bias[kH+j]
    i,f,o = sigmoid(pre_0,1,2)   n = tanh(pre_3)   c_t = f c_{t-1} + i n   h_t = o tanh(c_t)

with canonical gate order k = (i,f,o,n).  ``pack_*`` build the canonical set from the
reference's parameter names; ``forward`` / ``backward`` are the spec that the CUDA
kernels (vmlmf_b200/csrc) implement: same inputs, same outputs, same saved tensors.

Reference anchors: V/models/vmlmf.py:78-125 (plain), V/models/vmlmf_lm.py:222-269 (LM),
V/models/vmlmf_group.py:85-155 (group), :203-251 (group, no VM).
"""
from __future__ import annotations

import numpy as np

# h-side chunk position q used by the group cells for canonical gate k=(i,f,o,n):
# the group cells chunk their h-side (and, for the ablation, x-side) blocks as (f,i,n,o)
# (V/models/vmlmf_group.py:134,142,211,232).
GROUP_Q_OF_K = (1, 0, 3, 2)


def _np(t, dtype):
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t, dtype=dtype)


# --------------------------------------------------------------------------- #
# packing: reference parameters -> canonical parameters
# --------------------------------------------------------------------------- #


def pack_plain(p, dtype=np.float64, v_names=("v_x", "v_h")):
    """MyVMLMFCell (v_x/v_h) or MyVMLSTM (pass v_names=("w_x","w_h"))."""
    ux, uh = _np(p["u_x"], dtype), _np(p["u_h"], dtype)
    vx, vh = _np(p[v_names[0]], dtype), _np(p[v_names[1]], dtype)
    n_in, hidden = ux.shape[0], uh.shape[0]
    dia_x = _np(p["dia_x"], dtype).reshape(-1)
    dia_h = _np(p["dia_h"], dtype).reshape(-1)
    dx = dia_x[None] - np.einsum("jr,kjr->kj", ux, vx.reshape(4, hidden, -1)[:, :n_in])
    dh = dia_h[None] - np.einsum("jr,kjr->kj", uh, vh.reshape(4, hidden, -1))
    bias = _np(p["b_x"], dtype).reshape(-1) + _np(p["b_h"], dtype).reshape(-1)
    return dict(Ux=ux, Vx=vx, Dx=dx, A=uh, Bm=vh, Dh=dh, bias=bias)


def pack_group(p, g=2, dtype=np.float64, with_vm=True):
    """MyVMLMFCellg2 (with_vm) / MyVMLMFgCellg2 (not with_vm) -> dense canonical A,Bm with zero blocks."""
    ux, vx = _np(p["u_x"], dtype), _np(p["v_x"], dtype)
    n_in = ux.shape[0]
    hidden = vx.shape[0] // 4
    hg = hidden // g
    ranks = [p[f"u_h_{o}"].shape[2] for o in range(g)]
    rtot = g * sum(ranks)
    a = np.zeros((hidden, rtot), dtype)
    bm = np.zeros((4 * hidden, rtot), dtype)
    base = 0
    for off in range(g):
        u = _np(p[f"u_h_{off}"], dtype)      # [g,Hg,r]
        v = _np(p[f"v_h_{off}"], dtype)      # [g,r,4Hg]
        r = ranks[off]
        for j in range(g):                   # destination group
            s = (j + off) % g                # source group after `off` left-rotations (:123-124)
            cols = slice(base + j * r, base + (j + 1) * r)
            a[s * hg:(s + 1) * hg, cols] = u[j]
            for k in range(4):
                q = GROUP_Q_OF_K[k]
                bm[k * hidden + j * hg:k * hidden + (j + 1) * hg, cols] = v[j][:, q * hg:(q + 1) * hg].T
        base += g * r
    bx = _np(p["bias_x"], dtype).reshape(-1)
    bh = _np(p["bias_h"], dtype).reshape(-1)
    if with_vm:
        dia_x = _np(p["dia_x"], dtype).reshape(-1)
        dia_h = _np(p["dia_h"], dtype).reshape(-1)
        dx = dia_x[None] - np.einsum("jr,kjr->kj", ux, vx.reshape(4, hidden, -1)[:, :n_in])
        u0 = _np(p["u_h_0"], dtype)
        v0 = _np(p["v_h_0"], dtype)
        dh = np.empty((4, hidden), dtype)
        for k in range(4):
            q = GROUP_Q_OF_K[k]
            # diag of the offset-0 block only (:101-110)
            corr = np.einsum("jmr,jrm->jm", u0, v0[:, :, q * hg:(q + 1) * hg]).reshape(-1)
            dh[k] = dia_h - corr
        vx_c = vx                                            # x side is already (i,f,o,n) (:113)
        bias = bx + np.concatenate([bh[GROUP_Q_OF_K[k] * hidden:(GROUP_Q_OF_K[k] + 1) * hidden] for k in range(4)])
    else:
        dx = np.zeros((4, n_in), dtype)
        dh = np.zeros((4, hidden), dtype)
        perm = np.concatenate([np.arange(GROUP_Q_OF_K[k] * hidden, (GROUP_Q_OF_K[k] + 1) * hidden) for k in range(4)])
        vx_c = vx[perm]                                      # x side is (f,i,n,o) too (:211)
        bias = bx[perm] + bh[perm]
    return dict(Ux=ux, Vx=vx_c, Dx=dx, A=a, Bm=bm, Dh=dh, bias=bias)


# --------------------------------------------------------------------------- #
# canonical forward / backward   (time-major x: [T,B,I])
# --------------------------------------------------------------------------- #


def _sig(v):
    return 1.0 / (1.0 + np.exp(-v))


def forward(cp, x, h0=None, c0=None):
    """Returns (y[T,B,H], hT, cT, saved).  saved holds what the CUDA forward saves:
    gates[T,B,4,H] (i,f,o,n), c[T,B,H], z[T,B,R], zx[T,B,RX]."""
    t_len, batch, n_in = x.shape
    hidden = cp["A"].shape[0]
    dt = x.dtype
    h = np.zeros((batch, hidden), dt) if h0 is None else h0.astype(dt)
    c = np.zeros((batch, hidden), dt) if c0 is None else c0.astype(dt)
    y = np.empty((t_len, batch, hidden), dt)
    gates = np.empty((t_len, batch, 4, hidden), dt)
    cs = np.empty((t_len, batch, hidden), dt)
    zs = np.empty((t_len, batch, cp["A"].shape[1]), dt)
    zxs = np.empty((t_len, batch, cp["Ux"].shape[1]), dt)
    for t in range(t_len):
        zx = x[t] @ cp["Ux"]
        z = h @ cp["A"]
        pre = (zx @ cp["Vx"].T + z @ cp["Bm"].T + cp["bias"]).reshape(batch, 4, hidden)
        pre = pre + h[:, None, :] * cp["Dh"][None]
        pre[:, :, :n_in] += x[t][:, None, :] * cp["Dx"][None]
        i, f, o = _sig(pre[:, 0]), _sig(pre[:, 1]), _sig(pre[:, 2])
        n = np.tanh(pre[:, 3])
        c = f * c + i * n
        h = o * np.tanh(c)
        y[t], cs[t], zs[t], zxs[t] = h, c, z, zx
        gates[t, :, 0], gates[t, :, 1], gates[t, :, 2], gates[t, :, 3] = i, f, o, n
    return y, h, c, dict(gates=gates, c=cs, z=zs, zx=zxs)


def backward(cp, x, y, saved, dy=None, dhT=None, dcT=None, h0=None, c0=None):
    """BPTT in canonical form.  Returns dict of gradients for
    Ux,Vx,Dx,A,Bm,Dh,bias plus dx[T,B,I], dh0, dc0.  No [H,4H] matrix is ever formed."""
    t_len, batch, n_in = x.shape
    hidden = cp["A"].shape[0]
    dt = x.dtype
    zero = np.zeros((batch, hidden), dt)
    h0 = zero if h0 is None else h0
    c0 = zero if c0 is None else c0
    g = {k: np.zeros_like(v) for k, v in cp.items()}
    dx = np.zeros_like(x)
    dh_next = zero.copy() if dhT is None else dhT.astype(dt).copy()
    dc_next = zero.copy() if dcT is None else dcT.astype(dt).copy()
    for t in range(t_len - 1, -1, -1):
        i, f, o, n = (saved["gates"][t, :, k] for k in range(4))
        c_t = saved["c"][t]
        c_prev = saved["c"][t - 1] if t > 0 else c0
        h_prev = y[t - 1] if t > 0 else h0
        dh = dh_next + (dy[t] if dy is not None else 0)
        tc = np.tanh(c_t)
        dc = dh * o * (1 - tc * tc) + dc_next
        dpre = np.stack([dc * n * i * (1 - i), dc * c_prev * f * (1 - f),
                         dh * tc * o * (1 - o), dc * i * (1 - n * n)], 1)      # [B,4,H]
        dc_next = dc * f
        flat = dpre.reshape(batch, 4 * hidden)
        z, zx = saved["z"][t], saved["zx"][t]
        dz = flat @ cp["Bm"]
        dzx = flat @ cp["Vx"]
        g["Bm"] += flat.T @ z
        g["Vx"] += flat.T @ zx
        g["A"] += h_prev.T @ dz
        g["Ux"] += x[t].T @ dzx
        g["Dh"] += np.einsum("bkj,bj->kj", dpre, h_prev)
        g["Dx"] += np.einsum("bkj,bj->kj", dpre[:, :, :n_in], x[t])
        g["bias"] += flat.sum(0)
        dh_next = dz @ cp["A"].T + (dpre * cp["Dh"][None]).sum(1)
        dx[t] = dzx @ cp["Ux"].T + (dpre[:, :, :n_in] * cp["Dx"][None]).sum(1)
    g["dx"], g["dh0"], g["dc0"] = dx, dh_next, dc_next
    return g


# --------------------------------------------------------------------------- #
# chain rule back to the reference parameters (the spec for the pack backward)
# --------------------------------------------------------------------------- #


def unpack_grads_plain(p, g, dtype=np.float64, v_names=("v_x", "v_h")):
    """Canonical grads -> grads of u_x,u_h,v_x,v_h,b_x,b_h,dia_x,dia_h (SURVEY A.3 last line)."""
    ux, uh = _np(p["u_x"], dtype), _np(p["u_h"], dtype)
    vx, vh = _np(p[v_names[0]], dtype), _np(p[v_names[1]], dtype)
    n_in, hidden = ux.shape[0], uh.shape[0]
    out = {"dia_x": g["Dx"].sum(0)[None], "dia_h": g["Dh"].sum(0)[None],
           "b_x": g["bias"].copy(), "b_h": g["bias"].copy()}
    dux, duh = g["Ux"].copy(), g["A"].copy()
    dvx, dvh = g["Vx"].copy(), g["Bm"].copy()
    for k in range(4):
        dux -= g["Dx"][k][:, None] * vx[k * hidden:k * hidden + n_in]
        dvx[k * hidden:k * hidden + n_in] -= g["Dx"][k][:, None] * ux
        duh -= g["Dh"][k][:, None] * vh[k * hidden:(k + 1) * hidden]
        dvh[k * hidden:(k + 1) * hidden] -= g["Dh"][k][:, None] * uh
    out.update({"u_x": dux, "u_h": duh, v_names[0]: dvx, v_names[1]: dvh})
    return out


def unpack_grads_group(p, g_c, g=2, dtype=np.float64, with_vm=True):
    """Canonical grads -> grads of the MyVMLMFCellg2 / MyVMLMFgCellg2 ParameterDict entries."""
    ux, vx = _np(p["u_x"], dtype), _np(p["v_x"], dtype)
    n_in = ux.shape[0]
    hidden = vx.shape[0] // 4
    hg = hidden // g
    ranks = [p[f"u_h_{o}"].shape[2] for o in range(g)]
    out = {}
    base = 0
    for off in range(g):
        r = ranks[off]
        du = np.zeros((g, hg, r), dtype)
        dv = np.zeros((g, r, 4 * hg), dtype)
        for j in range(g):
            s = (j + off) % g
            cols = slice(base + j * r, base + (j + 1) * r)
            du[j] = g_c["A"][s * hg:(s + 1) * hg, cols]
            for k in range(4):
                q = GROUP_Q_OF_K[k]
                dv[j][:, q * hg:(q + 1) * hg] = g_c["Bm"][k * hidden + j * hg:k * hidden + (j + 1) * hg, cols].T
        out[f"u_h_{off}"], out[f"v_h_{off}"] = du, dv
        base += g * r
    perm = np.concatenate([np.arange(GROUP_Q_OF_K[k] * hidden, (GROUP_Q_OF_K[k] + 1) * hidden) for k in range(4)])
    dbh = np.zeros(4 * hidden, dtype)
    dbh[perm] = g_c["bias"]
    out["bias_h"] = dbh[None]
    if with_vm:
        out["bias_x"] = g_c["bias"][None].copy()
        out["dia_x"], out["dia_h"] = g_c["Dx"].sum(0)[None], g_c["Dh"].sum(0)[None]
        dux, dvx = g_c["Ux"].copy(), g_c["Vx"].copy()
        for k in range(4):
            dux -= g_c["Dx"][k][:, None] * vx[k * hidden:k * hidden + n_in]
            dvx[k * hidden:k * hidden + n_in] -= g_c["Dx"][k][:, None] * ux
        out["u_x"], out["v_x"] = dux, dvx
        u0, v0 = _np(p["u_h_0"], dtype), _np(p["v_h_0"], dtype)
        for k in range(4):
            q = GROUP_Q_OF_K[k]
            dd = g_c["Dh"][k].reshape(g, hg)                         # [g,Hg]
            out["u_h_0"] -= dd[:, :, None] * v0[:, :, q * hg:(q + 1) * hg].transpose(0, 2, 1)
            out["v_h_0"][:, :, q * hg:(q + 1) * hg] -= dd[:, None, :] * u0.transpose(0, 2, 1)
    else:
        dbx = np.zeros(4 * hidden, dtype)
        dbx[perm] = g_c["bias"]
        out["bias_x"] = dbx[None]
        dvx = np.zeros_like(vx)
        dvx[perm] = g_c["Vx"]
        out["u_x"], out["v_x"] = g_c["Ux"].copy(), dvx
    return out
