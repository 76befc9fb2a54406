"""Eager-torch CPU restatement of the reference VMLMF cells / layers / nets.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): never imported by the product.

Every function takes its weights as a plain ``dict`` keyed by the reference's
parameter names, so a reference ``state_dict`` (minus prefixes) can be fed
straight in.  Op order inside one timestep follows the reference so that fp32
outputs agree to rounding; per-step work (including the loop-invariant diagonal
corrections) is deliberately *not* hoisted, because this file is also the
"port" timed as the CPU baseline and must cost what the reference costs.

Reference root: /root/reference/rnn_compression_factorization_vmlmf/src  ("V/")
"""
from __future__ import annotations

import torch

# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #


def _gate_diag(u, v, n_rows, hidden):
    """sum_r u[j,r]*v[k*hidden+j,r] for the four gate blocks -> [4, n_rows].

    The reference recomputes this inside every step, gate by gate
    (V/models/vmlmf.py:102-106, V/models/vmlmf_lm.py:250-255)."""
    return torch.stack([(u * v[k * hidden:k * hidden + n_rows]).sum(1) for k in range(4)])


def _finish(i_pre, f_pre, o_pre, n_pre, c):
    """Gate non-linearities and state update (V/models/vmlmf.py:117-125)."""
    i = torch.sigmoid(i_pre)
    f = torch.sigmoid(f_pre)
    o = torch.sigmoid(o_pre)
    n = torch.tanh(n_pre)
    c_new = f * c + i * n
    return o * torch.tanh(c_new), c_new


# --------------------------------------------------------------------------- #
# a2: plain VMLMF cell  (V/models/vmlmf.py:78-125)
# --------------------------------------------------------------------------- #


def plain_cell_step(p, x, h, c):
    """One step of MyVMLMFCell.  p: u_x,u_h,v_x,v_h,b_x,b_h,dia_x,dia_h."""
    batch, n_in = x.shape
    hidden = h.shape[1]
    if hidden < n_in:
        # the reference builds vm_x=None in this case and dies on the add (:92-94,:117)
        raise TypeError("MyVMLMFCell needs hidden_size >= input_size")
    # vector-multiplication terms, shared by the four gates (:92-95)
    vm_x = torch.cat([p["dia_x"] * x, x.new_zeros(batch, hidden - n_in)], 1)
    vm_h = p["dia_h"] * h
    # low-rank products (:98-99)
    low_x = (x @ p["u_x"]) @ p["v_x"].t()
    low_h = (h @ p["u_h"]) @ p["v_h"].t()
    # remove what the low-rank product put on the diagonal (:102-106)
    dgx = _gate_diag(p["u_x"], p["v_x"], n_in, hidden)
    dgh = _gate_diag(p["u_h"], p["v_h"], hidden, hidden)
    fix_x = x.new_zeros(batch, 4 * hidden)
    fix_h = x.new_zeros(batch, 4 * hidden)
    for k in range(4):
        fix_x[:, k * hidden:k * hidden + n_in] = x * dgx[k]
        fix_h[:, k * hidden:(k + 1) * hidden] = h * dgh[k]
    gx = low_x - fix_x + p["b_x"]          # (:109)
    gh = low_h - fix_h + p["b_h"]          # (:110)
    xi, xf, xo, xn = gx.chunk(4, 1)        # gate order i,f,o,n (:113-114)
    hi, hf, ho, hn = gh.chunk(4, 1)
    return _finish(xi + hi + vm_x + vm_h, xf + hf + vm_x + vm_h,
                   xo + ho + vm_x + vm_h, xn + hn + vm_x + vm_h, c)


# --------------------------------------------------------------------------- #
# f4: uncompressed / plain low-rank baseline cell  (V/models/vmlmf.py:127-238)
# --------------------------------------------------------------------------- #


def lstm_cell_step(p, x, h, c):
    """One step of MyLSTMCell.  p: [w,] w1..w4, [u,] u1..u4, bias_f, bias_i, bias_c, bias_o.
    `w` / `u` present = low-rank factorisation of that side (:193-220); gate k uses (w_k, u_k) in the order
    i, f, o, c~ (:222-231)."""
    xs = x @ p["w"] if "w" in p else x
    hs = h @ p["u"] if "u" in p else h
    pre = [xs @ p[f"w{k}"] + hs @ p[f"u{k}"] for k in (1, 2, 3, 4)]
    return _finish(pre[0] + p["bias_i"], pre[1] + p["bias_f"], pre[2] + p["bias_o"], pre[3] + p["bias_c"], c)


# --------------------------------------------------------------------------- #
# a6 / a7: group cells  (V/models/vmlmf_group.py:85-155, :203-251)
# --------------------------------------------------------------------------- #


def _group_lowrank_h(p, h, g):
    """Sum over rotation offsets of the per-group two-stage bmm (:118-132).

    Returns [B, g, 4*Hg]; the last axis is chunked (f,i,n,o) by the caller."""
    batch, hidden = h.shape
    hg = hidden // g
    order = list(range(g))
    acc = None
    for off in range(g):
        if off > 0:
            order = order[1:] + order[:1]
        hv = h.view(batch, g, hg)
        if off > 0:
            hv = hv[:, order, :]
        hv = hv.transpose(0, 1)                               # [g,B,Hg]
        t = torch.bmm(torch.bmm(hv, p[f"u_h_{off}"]), p[f"v_h_{off}"])  # [g,B,4Hg]
        t = t.transpose(0, 1)
        acc = t if acc is None else acc + t
    return acc


def group_cell_step(p, x, h, c, g=2):
    """One step of MyVMLMFCellg2.  p: dia_x,dia_h,u_x,v_x,u_h_i,v_h_i,bias_x,bias_h."""
    batch, n_in = x.shape
    hidden = h.shape[1]
    hg = hidden // g
    if hidden < n_in:
        raise TypeError("MyVMLMFCellg2 needs hidden_size >= input_size")
    vm_x = torch.cat([p["dia_x"] * x, x.new_zeros(batch, hidden - n_in)], 1)   # (:92-94)
    vm_h = p["dia_h"] * h                                                      # (:95)
    low_x = (x @ p["u_x"]) @ p["v_x"].t()                                      # (:98)
    r0 = p["u_h_0"].shape[2]
    u0 = p["u_h_0"].reshape(hidden, r0)                                        # (:101)
    v0t = p["v_h_0"].transpose(1, 2).contiguous()                              # [g,4Hg,r0] (:102)
    fix_x = x.new_zeros(batch, 4 * hidden)
    fix_h = x.new_zeros(batch, 4 * hidden)
    for q in range(4):                                                         # (:104-110)
        fix_x[:, q * hidden:q * hidden + n_in] = x * (p["u_x"] * p["v_x"][q * hidden:q * hidden + n_in]).sum(1)
        vq = v0t[:, q * hg:(q + 1) * hg, :].reshape(-1, r0)
        fix_h[:, q * hidden:(q + 1) * hidden] = h * (u0 * vq).sum(1)
    gx = low_x - fix_x + p["bias_x"]                                           # (:112)
    xi, xf, xo, xn = gx.chunk(4, 1)                                            # x side: i,f,o,n (:113)
    mix = _group_lowrank_h(p, h, g)
    f_h, i_h, n_h, o_h = [t.contiguous().view(batch, hidden) for t in mix.chunk(4, 2)]  # (:134-139)
    gh = p["bias_h"] - fix_h                                                   # (:141)
    hf, hi, hn, ho = gh.chunk(4, 1)                                            # h side: f,i,n,o (:142)
    return _finish(xi + (hi + i_h) + vm_x + vm_h, xf + (hf + f_h) + vm_x + vm_h,
                   xo + (ho + o_h) + vm_x + vm_h, xn + (hn + n_h) + vm_x + vm_h, c)


def group_ablation_step(p, x, h, c, g=2):
    """One step of MyVMLMFgCellg2 (no vector-multiplication terms), :203-251.

    Gate order is (f,i,n,o) on BOTH sides here (:211, :232)."""
    batch = x.shape[0]
    hidden = h.shape[1]
    gx = (x @ p["u_x"]) @ p["v_x"].t() + p["bias_x"]
    xf, xi, xn, xo = gx.chunk(4, 1)
    mix = _group_lowrank_h(p, h, g)
    f_h, i_h, n_h, o_h = [t.contiguous().view(batch, hidden) for t in mix.chunk(4, 2)]
    hf, hi, hn, ho = p["bias_h"].chunk(4, 1)
    return _finish(xi + (hi + i_h), xf + (hf + f_h), xo + (ho + o_h), xn + (hn + n_h), c)


# --------------------------------------------------------------------------- #
# a9 / a10: LM layer  (V/models/vmlmf_lm.py:222-280)
# --------------------------------------------------------------------------- #


def lm_cell_step(p, x, h, c):
    """MyVMLSTM.lstm_step.  p: u_x,u_h,w_x,w_h,b_x,b_h,dia_x,dia_h; needs I == H."""
    batch, n_in = x.shape
    hidden = h.shape[1]
    vm_x = torch.cat([p["dia_x"] * x] * 4, 1)        # (:241-244) -> [B,4I]; added to [B,4H] => I==H
    vm_h = torch.cat([p["dia_h"] * h] * 4, 1)
    low_x = (x @ p["u_x"]) @ p["w_x"].t()            # (:246-247)
    low_h = (h @ p["u_h"]) @ p["w_h"].t()
    dgx = _gate_diag(p["u_x"], p["w_x"], n_in, hidden)
    dgh = _gate_diag(p["u_h"], p["w_h"], hidden, hidden)
    fix_x = x.new_zeros(batch, 4 * hidden)
    fix_h = x.new_zeros(batch, 4 * hidden)
    for k in range(4):                               # (:251-255)
        fix_x[:, k * hidden:k * hidden + n_in] = x * dgx[k]
        fix_h[:, k * hidden:(k + 1) * hidden] = h * dgh[k]
    gx = vm_x + low_x - fix_x + p["b_x"]             # (:256)
    gh = vm_h + low_h - fix_h + p["b_h"]             # (:257)
    xi, xf, xo, xn = gx.chunk(4, 1)
    hi, hf, ho, hn = gh.chunk(4, 1)
    return _finish(xi + hi, xf + hf, xo + ho, xn + hn, c)


def lm_layer(p, x, state):
    """MyVMLSTM.forward: time-major [T,B,X], carried (h,c) (:272-280)."""
    h, c = state
    outs = []
    for x_t in x.unbind(0):
        h, c = lm_cell_step(p, x_t, h, c)
        outs.append(h)
    return torch.stack(outs), (h, c)


def lm_group_cell_step(p, x, h, c, g=2):
    """MyVMLSTMGroup.lstm_step (V/models/vmlmf_lm.py:97-160), bug-for-bug.

    p: u_x,w_x,u_h.{i},v_h.{i},b_x,b_h,dia_x,dia_h.  Mirrors: the hard-coded 40-row scratch
    (:112-113), the group-major flatten of the bmm result before chunk(4) (:135,:155) and the
    re_uh/re_vh "diagonal" (:141-148)."""
    hidden = h.shape[1]
    n_in = x.shape[1]
    fix_x = x.new_zeros(40, 4 * hidden)
    fix_h = x.new_zeros(40, 4 * hidden)
    vm_x = torch.cat([p["dia_x"] * x] * 4, 1)
    vm_h = torch.cat([p["dia_h"] * h] * 4, 1)
    low_x = (x @ p["u_x"]) @ p["w_x"].t()
    order = list(range(g))
    low_h = None
    for off in range(g):
        hv = h.view(-1, g, hidden // g)
        if off > 0:
            order = order[1:] + order[:1]
            hv = hv[:, order, :]
        hv = hv.transpose(0, 1)
        t = torch.bmm(torch.bmm(hv, p[f"u_h.{off}"]), p[f"v_h.{off}"]).transpose(0, 1)
        t = t.contiguous().view(-1, 4 * hidden)
        low_h = t if low_h is None else t + low_h
    r0 = p["u_h.0"].shape[2]
    re_u = p["u_h.0"].reshape(hidden, r0)
    re_v = p["v_h.0"].transpose(1, 2).contiguous().view(4 * hidden, r0)
    for k in range(4):
        fix_x[:, k * hidden:k * hidden + n_in] = x * (p["u_x"] * p["w_x"][k * hidden:k * hidden + n_in]).sum(1)
        fix_h[:, k * hidden:(k + 1) * hidden] = h * (re_u * re_v[k * hidden:(k + 1) * hidden]).sum(1)
    gx = vm_x + low_x - fix_x + p["b_x"]
    gh = vm_h + low_h - fix_h + p["b_h"]
    xi, xf, xo, xn = gx.chunk(4, 1)
    hi, hf, ho, hn = gh.chunk(4, 1)
    return _finish(xi + hi, xf + hf, xo + ho, xn + hn, c)


# --------------------------------------------------------------------------- #
# a3 / a4: layer stack and classifier net  (V/models/vmlmf.py:294-316, :352-355)
# --------------------------------------------------------------------------- #

_STEP = {"plain": plain_cell_step, "group": group_cell_step, "group_novm": group_ablation_step, "lstm": lstm_cell_step}


def layer_stack(cells, x, kind="plain", batch_first=True, **kw):
    """MyLSTM.forward: zero initial state per layer, python time loop, returns
    (sequence of the last layer, cat of every layer's last h)."""
    t_dim = 1 if batch_first else 0
    b_dim = 0 if batch_first else 1
    step = _STEP[kind]
    last = []
    for p in cells:
        hidden = (p["dia_h"].shape[1] if "dia_h" in p else p["bias_f"].shape[1] if "bias_f" in p
                  else p["bias_h"].shape[1] // 4)
        h = x.new_zeros(x.size(b_dim), hidden)
        c = x.new_zeros(x.size(b_dim), hidden)
        outs = []
        for x_t in torch.unbind(x, t_dim):
            h, c = step(p, x_t, h, c, **kw)
            outs.append(h)
        x = torch.stack(outs, t_dim)
        last.append(h)
    return x, torch.cat(last, -1)


def net_forward(cells, lin_w, lin_b, x, kind="plain", **kw):
    """Net.forward: only the last timestep feeds the 18-way head (:352-355)."""
    y, _ = layer_stack(cells, x, kind=kind, **kw)
    return torch.nn.functional.linear(y[:, -1], lin_w, lin_b).squeeze(1)


# --------------------------------------------------------------------------- #
# a11: LM model, loss  (V/models/vmlmf_lm.py:433-441, V/train_test/lm_test.py:140-153)
# --------------------------------------------------------------------------- #


def lm_model_forward(embed_w, layers, fc_w, fc_b, tokens, states, drop_p=0.0, training=False):
    """Model.forward with lstm_type == "vmlmf": embed -> dropout -> layers -> dropout -> fc."""
    drop = (lambda t: torch.nn.functional.dropout(t, drop_p, training)) if drop_p > 0 else (lambda t: t)
    x = drop(embed_w[tokens])
    new_states = []
    for p, st in zip(layers, states):
        x, st = lm_layer(p, x, st)
        new_states.append(st)
        x = drop(x)
    scores = torch.addmm(fc_b, x.view(-1, x.size(2)), fc_w.t())
    return scores, new_states


def lm_nll_loss(scores, y):
    """lm_test.py:140-153 -- hand-rolled softmax NLL, token mean times batch size."""
    batch = y.size(1)
    e = scores.exp()
    prob = e / e.sum(1, keepdim=True)
    pick = prob[torch.arange(y.numel()), y.reshape(-1)]
    return torch.mean(-torch.log(pick) * batch)


# --------------------------------------------------------------------------- #
# convenience: split a reference state_dict into per-cell dicts
# --------------------------------------------------------------------------- #


def split_state_dict(sd, prefix):
    """{'<prefix>name': t} -> {'name': t} for the keys under prefix."""
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
