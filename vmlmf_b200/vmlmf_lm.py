"""Drop-in replacements for V/models/vmlmf_lm.py: the language-model VMLMF layer and network.

MyVMLSTM.forward runs its whole [T,B,X] window in one fused call with the carried (h,c) state
(reference: python loop over lstm_step, V/models/vmlmf_lm.py:272-280).  Embedding gather, dropout and
the vocabulary projection are ordinary dense ops around the path and stay in PyTorch."""
from __future__ import annotations

import torch
from torch import nn

from . import packing
from .functional import vmlmf_plain_sequence, vmlmf_sequence  # noqa: F401


class Embed(nn.Module):
    """index -> row of w (V/models/vmlmf_lm.py:33-51)"""

    def __init__(self, vocab_size, embed_size):
        super().__init__()
        self.vocab_size, self.embed_size = vocab_size, embed_size
        self.w = nn.Parameter(torch.Tensor(vocab_size, embed_size))

    def forward(self, x):
        # same values as the reference's `self.w[x]` (:48); F.embedding's backward needs no host synchronisation, so the
        # whole LM step can be captured in a CUDA graph
        return torch.nn.functional.embedding(x, self.w)

    def __repr__(self):
        return f"Embedding(vocab: {self.vocab_size}, embedding: {self.embed_size})"


class MyVMLSTM(nn.Module):
    """VMLMF LSTM layer for the LM (V/models/vmlmf_lm.py:178-280); needs input_size == hidden_size
    because the reference adds a [B,4I] tensor to a [B,4H] one (:243,:256).  Parameters are created
    uninitialised like the reference; Model.reset_parameters fills them."""

    def __init__(self, input_size, hidden_size, dropout=0, w_rank=None, u_ranks=None):
        super().__init__()
        self.input_size, self.hidden_size, self.dropout = input_size, hidden_size, dropout
        self.w_rank, self.u_ranks = w_rank, u_ranks
        self.u_x = nn.Parameter(torch.Tensor(input_size, w_rank))
        self.u_h = nn.Parameter(torch.Tensor(hidden_size, u_ranks))
        self.w_x = nn.Parameter(torch.Tensor(4 * hidden_size, w_rank))
        self.w_h = nn.Parameter(torch.Tensor(4 * hidden_size, u_ranks))
        self.b_x = nn.Parameter(torch.Tensor(4 * hidden_size))
        self.b_h = nn.Parameter(torch.Tensor(4 * hidden_size))
        self.dia_x = nn.Parameter(torch.Tensor(1, input_size))
        self.dia_h = nn.Parameter(torch.Tensor(1, hidden_size))
        self.cnt = 0

    def __repr__(self):
        return f"LSTM(input: {self.input_size}, hidden: {self.hidden_size})"

    def canonical(self):
        if self.input_size != self.hidden_size:
            raise RuntimeError("MyVMLSTM requires input_size == hidden_size (as the reference does)")
        return packing.pack_plain(self.u_x, self.u_h, self.w_x, self.w_h, self.b_x, self.b_h, self.dia_x, self.dia_h)

    def plain_params(self):
        if self.input_size != self.hidden_size:
            raise RuntimeError("MyVMLSTM requires input_size == hidden_size (as the reference does)")
        return (self.u_x, self.u_h, self.w_x, self.w_h, self.b_x, self.b_h, self.dia_x, self.dia_h)

    def lstm_step(self, x, h, c):
        _, h1, c1 = vmlmf_plain_sequence(x.unsqueeze(0), h, c, self.plain_params(), batch_first=False)
        return h1, c1

    def forward(self, x, states):
        """x[T,B,X], (h,c) -> (out[T,B,H], (h_T, c_T))"""
        h, c = states
        out, h1, c1 = vmlmf_plain_sequence(x, h, c, self.plain_params(), batch_first=False)
        return out, (h1, c1)


class MyVMLSTMGroup(nn.Module):
    """Group VMLMF layer for the LM (V/models/vmlmf_lm.py:53-174), arithmetic as shipped (SURVEY B-5).

    The reference hard-codes 40-row scratch tensors (:112-113): it runs at batch 40, and a batch of 1 broadcasts
    to 40 identical rows; both behaviours are kept, any other batch raises like the reference's shape error.
    Needs input_size == hidden_size (:116-119 add a [B,4I] tensor to a [B,4H] one)."""

    BATCH = 40

    def __init__(self, input_size, hidden_size, dropout=0, w_rank=None, u_ranks=None, g=2):
        super().__init__()
        self.input_size, self.hidden_size, self.dropout, self.g = input_size, hidden_size, dropout, g
        self.w_rank, self.u_ranks = w_rank, u_ranks
        hg = int(hidden_size / g)
        self.u_x = nn.Parameter(torch.Tensor(input_size, w_rank))
        self.w_x = nn.Parameter(torch.Tensor(4 * hidden_size, w_rank))
        self.u_h = nn.ParameterList([nn.Parameter(torch.Tensor(g, hg, u_ranks[i])) for i in range(g)])
        self.v_h = nn.ParameterList([nn.Parameter(torch.Tensor(g, u_ranks[i], 4 * hg)) for i in range(g)])
        self.b_x = nn.Parameter(torch.Tensor(4 * hidden_size))
        self.b_h = nn.Parameter(torch.Tensor(4 * hidden_size))
        self.dia_x = nn.Parameter(torch.Tensor(1, input_size))
        self.dia_h = nn.Parameter(torch.Tensor(1, hidden_size))
        self.cnt = 0

    def __repr__(self):
        return f"LSTM(input: {self.input_size}, hidden: {self.hidden_size})"

    def canonical(self):
        if self.input_size != self.hidden_size:
            raise RuntimeError("MyVMLSTMGroup requires input_size == hidden_size (as the reference does)")
        return packing.pack_lm_group(self.u_x, self.w_x, list(self.u_h), list(self.v_h), self.b_x, self.b_h,
                                     self.dia_x, self.dia_h, self.g)

    def _batch40(self, x, h, c, bdim):
        nb = x.size(bdim)
        if nb == 1:                                   # the 40-row scratch broadcasts a single sequence to 40 rows
            x = x.expand(*x.shape[:bdim], self.BATCH, *x.shape[bdim + 1:])
            h, c = h.expand(self.BATCH, -1), c.expand(self.BATCH, -1)
        elif nb != self.BATCH:
            raise RuntimeError(f"The expanded size of the tensor ({self.BATCH}) must match the existing size ({nb}) at "
                               "non-singleton dimension 0 (MyVMLSTMGroup is hard-wired to batch 40, vmlmf_lm.py:112)")
        return x, h, c

    def lstm_step(self, x, h, c):
        x, h, c = self._batch40(x, h, c, 0)
        _, h1, c1 = vmlmf_sequence(x.unsqueeze(0), h, c, self.canonical(), batch_first=False)
        return h1, c1

    def forward(self, x, states):
        """x[T,40,X], (h,c) -> (out[T,40,H], (h_T, c_T))"""
        h, c = states
        x, h, c = self._batch40(x, h, c, 1)
        out, h1, c1 = vmlmf_sequence(x, h, c, self.canonical(), batch_first=False)
        return out, (h1, c1)


class LSTM(nn.Module):
    """Plain dense LSTM layer, the reference's "custom" baseline (V/models/vmlmf_lm.py:283-339).  On CUDA it runs through the
    same fused kernels as the compressed layers, as a canonical recurrence with identity first factors and no
    vector-multiplication terms (pre = x W_x^T + h W_h^T + b_x + b_h; +25 % GEMM work for the identity products, like for
    like otherwise), so compressed-vs-dense speed comparisons use one code path.  Host tensors keep the eager loop."""

    def __init__(self, input_size, hidden_size, dropout=0):
        super().__init__()
        self.input_size, self.hidden_size, self.dropout = input_size, hidden_size, dropout
        self.w_x = nn.Parameter(torch.Tensor(4 * hidden_size, input_size))
        self.w_h = nn.Parameter(torch.Tensor(4 * hidden_size, hidden_size))
        self.b_x = nn.Parameter(torch.Tensor(4 * hidden_size))
        self.b_h = nn.Parameter(torch.Tensor(4 * hidden_size))

    def __repr__(self):
        return f"LSTM(input: {self.input_size}, hidden: {self.hidden_size})"

    def canonical(self):
        eye_x = torch.eye(self.input_size, device=self.w_x.device, dtype=self.w_x.dtype)
        eye_h = eye_x if self.input_size == self.hidden_size else torch.eye(self.hidden_size, device=self.w_x.device, dtype=self.w_x.dtype)
        zx = self.w_x.new_zeros(4, self.input_size)
        zh = self.w_x.new_zeros(4, self.hidden_size)
        return eye_x, self.w_x, zx, eye_h, self.w_h, zh, self.b_x + self.b_h * 1.0

    def forward(self, x, states):
        h, c = states
        if x.is_cuda and x.dtype == torch.float32 and self.input_size <= self.hidden_size:
            out, h1, c1 = vmlmf_sequence(x, h, c, self.canonical(), batch_first=False)
            return out, (h1, c1)
        gx_all = torch.addmm(self.b_x, x.reshape(-1, x.size(2)), self.w_x.t()).view(x.size(0), x.size(1), -1)
        outs = []
        for gx in gx_all.unbind(0):
            i, f, o, n = (gx + torch.addmm(self.b_h, h, self.w_h.t())).chunk(4, 1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(n)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        return torch.stack(outs), (h, c)


class Linear(nn.Module):
    """[T,B,H] -> [T*B, V] projection (V/models/vmlmf_lm.py:341-361)"""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.w = nn.Parameter(torch.Tensor(hidden_size, input_size))
        self.b = nn.Parameter(torch.Tensor(hidden_size))

    def forward(self, x):
        x2 = x.reshape(-1, x.size(2))
        if x2.is_cuda and x2.dtype == torch.float32:
            # vocabulary projection on the fp32-accurate tcgen05 GEMMs (forward, dX: vmlmf_gemm_nt; dW: vmlmf_gemm_tn)
            from .functional import linear_tc
            return linear_tc(x2, self.w, self.b)
        return torch.addmm(self.b, x2, self.w.t())      # host tensors (state_dict / shape tests): no kernel of ours involved

    def __repr__(self):
        return f"FC(input: {self.input_size}, output: {self.hidden_size})"


class Model(nn.Module):
    """Embed -> dropout -> L x LSTM layers -> dropout -> FC (V/models/vmlmf_lm.py:363-441).

    lstm_type "vmlmf" selects the fused VMLMF layers, "vmgroup" the group layers (batch 40 only, like the
    reference's layer); "custom" / "pytorch" (and the reference's misspelt "vm_group") build the dense baselines."""

    def __init__(self, vocab_size, hidden_size, layer_num, dropout, winit, w_rank=None, u_ranks=None,
                 lstm_type="pytorch"):
        super().__init__()
        self.vocab_size, self.hidden_size, self.layer_num = vocab_size, hidden_size, layer_num
        self.winit, self.lstm_type = winit, lstm_type
        self.embed = Embed(vocab_size, hidden_size)
        # The reference reduces u_ranks to its last entry unless lstm_type is "vm_group" (:387-388) and builds the group
        # layers when it is "vmgroup" (:390): as shipped "vmgroup" dies on `u_ranks[g_idx]` of an int and "vm_group"
        # falls through to nn.LSTM (SURVEY B-4).  Here "vmgroup" keeps the rank list and builds what was meant;
        # "vm_group" still gets the reference's nn.LSTM.
        if u_ranks is not None and lstm_type != "vmgroup":
            u_ranks = u_ranks[-1] if lstm_type != "vm_group" else u_ranks
        if lstm_type == "vmgroup":
            rnns = [MyVMLSTMGroup(hidden_size, hidden_size, w_rank=w_rank, u_ranks=u_ranks) for _ in range(layer_num)]
        elif lstm_type == "vmlmf":
            rnns = [MyVMLSTM(hidden_size, hidden_size, w_rank=w_rank, u_ranks=u_ranks) for _ in range(layer_num)]
        elif lstm_type == "custom":
            rnns = [LSTM(hidden_size, hidden_size) for _ in range(layer_num)]
        else:
            rnns = [nn.LSTM(hidden_size, hidden_size) for _ in range(layer_num)]
        self.rnns = nn.ModuleList(rnns)
        self.fc = Linear(hidden_size, vocab_size)
        self.dropout = nn.Dropout(p=dropout)
        self.reset_parameters()

    def reset_parameters(self):
        for param in self.parameters():
            nn.init.uniform_(param, -self.winit, self.winit)

    def state_init(self, batch_size):
        dev = next(self.parameters()).device
        flat = self.lstm_type in ("custom", "vmlmf", "vmgroup", "hmd")
        shape = (lambda l: (batch_size, l.hidden_size)) if flat else (lambda l: (1, batch_size, l.hidden_size))
        return [(torch.zeros(*shape(l), device=dev), torch.zeros(*shape(l), device=dev)) for l in self.rnns]

    def detach(self, states):
        return [(h.detach(), c.detach()) for (h, c) in states]

    def forward(self, x, states):
        if self.embed.w.is_cuda and self.embed.w.dtype == torch.float32:
            # Embed + dropout fused, output pitch-padded for the TMA-fed x projection of the first layer
            from .functional import embed_dropout
            x = embed_dropout(x, self.embed.w, self.dropout.p, self.training)
        else:
            x = self.dropout(self.embed(x))
        for i, rnn in enumerate(self.rnns):
            x, states[i] = rnn(x, states[i])
            x = self.dropout(x)
        return self.fc(x), states
