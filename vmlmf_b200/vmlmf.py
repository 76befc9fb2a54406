"""Drop-in replacements for the reference's V/models/vmlmf.py classes (same constructor
signatures, parameter names/shapes, state_dict keys, forward signatures and return values).

MyVMLMFCell / MyLSTM / Net run on the fused sm_100a kernels: MyLSTM.forward issues ONE fused call
per layer instead of the reference's `for t in range(seqlen): h, c = cell(...)` loop
(V/models/vmlmf.py:308-310).  MyLSTMCell (the uncompressed / plain low-rank baseline,
V/models/vmlmf.py:127-238) runs through the same kernels as a canonical recurrence without vector-multiplication
terms (identity first factor on a dense side), so compressed-vs-dense comparisons share one code path.
"""
from __future__ import annotations

import torch
from torch import nn

from . import packing
from .functional import head_linear, vmlmf_plain_sequence, vmlmf_sequence

TIME_STEPS = 128
RECURRENT_MAX = pow(2, 1 / TIME_STEPS)
RECURRENT_MIN = pow(1 / 2, 1 / TIME_STEPS)


def _scalar_rank(u_ranks):
    return u_ranks[-1] if isinstance(u_ranks, (list, tuple)) and len(u_ranks) < 2 else u_ranks


class MyVMLMFCell(nn.Module):
    """VMLMF LSTM cell, W_k = offdiag(U V_k^T) + diag(dia) on both sides (V/models/vmlmf.py:38-125).

    Parameters are created in the reference's order with the reference's initialiser
    (0.1 * randn, :56-69), so the same torch seed yields the same weights."""

    def __init__(self, input_size, hidden_size, w_rank=None, u_ranks=None):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.w_rank = w_rank
        self.u_ranks = _scalar_rank(u_ranks)
        ur = self.u_ranks
        self.u_x = nn.Parameter(0.1 * torch.randn([input_size, w_rank]))
        self.u_h = nn.Parameter(0.1 * torch.randn([hidden_size, ur]))
        self.v_x = nn.Parameter(0.1 * torch.randn([4 * hidden_size, w_rank]))
        self.v_h = nn.Parameter(0.1 * torch.randn([4 * hidden_size, ur]))
        self.b_x = nn.Parameter(0.1 * torch.randn([4 * hidden_size]))
        self.b_h = nn.Parameter(0.1 * torch.randn([4 * hidden_size]))
        self.dia_x = nn.Parameter(0.1 * torch.randn([1, input_size]))
        self.dia_h = nn.Parameter(0.1 * torch.randn([1, hidden_size]))
        self.cnt = 0

    def __repr__(self):
        return (f"LSTM_FINAL(input: {self.input_size}, hidden: {self.hidden_size}, "
                f"w_rank: {self.w_rank}, u_ranks: {self.u_ranks})")

    def canonical(self):
        """(Ux,Vx,Dx,A,Bm,Dh,bias) for the fused kernels."""
        return packing.pack_plain(self.u_x, self.u_h, self.v_x, self.v_h, self.b_x, self.b_h, self.dia_x, self.dia_h)

    def plain_params(self):
        """the reference's eight parameters in PlainCellSeqFunction's order"""
        return (self.u_x, self.u_h, self.v_x, self.v_h, self.b_x, self.b_h, self.dia_x, self.dia_h)

    def forward(self, x, hidden_states):
        """One step: x[B,I], (h[B,H], c[B,H]) -> (h', c').  Runs the fused kernel with T=1."""
        h, c = hidden_states
        _, h1, c1 = vmlmf_plain_sequence(x.unsqueeze(1), h, c, self.plain_params(), batch_first=True)
        return h1, c1


class MyLSTMCell(nn.Module):
    """Vanilla (w_rank/u_ranks None) or plain low-rank LSTM cell -- the reference's uncompressed
    baseline (V/models/vmlmf.py:127-238).  It is the canonical recurrence with no diagonal terms
    (Dx = Dh = 0): low-rank sides map to (Ux, Vx) / (A, Bm) directly, dense sides use an identity
    first factor, so the baseline runs through the same kernels as the compressed cells and the
    compressed-vs-vanilla comparison is like for like.  `hidden_size < input_size`,
    which the canonical form cannot express, is stepped in eager PyTorch."""

    def __init__(self, input_size, hidden_size, w_rank=None, u_ranks=None, recurrent_init=None, hidden_init=None):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.recurrent_init, self.hidden_init = recurrent_init, hidden_init
        self.w_rank = w_rank
        self.u_ranks = u_ranks[0] if isinstance(u_ranks, list) else u_ranks

        def mat(*shape):
            return nn.Parameter(0.1 * torch.randn(list(shape)))

        if w_rank is not None:
            self.w = mat(input_size, w_rank)
        for n in ("w1", "w2", "w3", "w4"):
            setattr(self, n, mat(input_size if w_rank is None else w_rank, hidden_size))
        if u_ranks is not None:
            self.u = mat(hidden_size, self.u_ranks)
        for n in ("u1", "u2", "u3", "u4"):
            setattr(self, n, mat(hidden_size if u_ranks is None else self.u_ranks, hidden_size))
        for n in ("bias_f", "bias_i", "bias_c", "bias_o"):
            setattr(self, n, nn.Parameter(torch.ones([1, hidden_size])))

    def canonical(self):
        """(Ux, Vx, Dx, A, Bm, Dh, bias) in gate order (i, f, o, c~) = (w1..w4, u1..u4) (:222-231); None if H < I."""
        n_in, hidden = self.input_size, self.hidden_size
        if hidden < n_in:
            return None
        ref = self.w1
        ux = self.w if self.w_rank is not None else torch.eye(n_in, dtype=ref.dtype, device=ref.device)
        a = self.u if self.u_ranks is not None else torch.eye(hidden, dtype=ref.dtype, device=ref.device)
        vx = torch.cat([self.w1, self.w2, self.w3, self.w4], 1).t()          # [4H, RX]
        bm = torch.cat([self.u1, self.u2, self.u3, self.u4], 1).t()          # [4H, RH]
        bias = torch.cat([self.bias_i, self.bias_f, self.bias_o, self.bias_c], 1).reshape(-1)
        return ux, vx, ref.new_zeros(4, n_in), a, bm, ref.new_zeros(4, hidden), bias

    def forward(self, x, hidden_states):
        h, c = hidden_states
        canon = self.canonical() if x.is_cuda else None
        if canon is not None:
            _, h1, c1 = vmlmf_sequence(x.unsqueeze(1), h, c, canon, batch_first=True)
            return h1, c1
        xs = x if self.w_rank is None else x @ self.w
        hs = h if self.u_ranks is None else h @ self.u
        pre = [xs @ getattr(self, f"w{k}") + hs @ getattr(self, f"u{k}") for k in (1, 2, 3, 4)]   # i, f, o, c~
        i = torch.sigmoid(pre[0] + self.bias_i)
        f = torch.sigmoid(pre[1] + self.bias_f)
        o = torch.sigmoid(pre[2] + self.bias_o)
        n = torch.tanh(pre[3] + self.bias_c)
        c_next = f * c + i * n
        return o * torch.tanh(c_next), c_next


class MyLSTM(nn.Module):
    """Stack of cells over a sequence (V/models/vmlmf.py:241-316).

    forward(x) -> (sequence of the last layer, cat over layers of the last hidden state).
    Cells exposing `canonical()` run fused (one kernel sequence per layer); any other cell class
    (e.g. MyLSTMCell) is stepped in python exactly like the reference."""

    def __init__(self, input_size, hidden_layer_sizes=None, batch_first=True, recurrent_inits=None,
                 hidden_inits=None, w_rank=None, u_ranks=None, cell=MyLSTMCell, **kwargs):
        super().__init__()
        if hidden_layer_sizes is None:
            hidden_layer_sizes = [32, 32]
        self.input_size = input_size
        self.hidden_layer_sizes = hidden_layer_sizes
        self.batch_first = batch_first
        self.w_rank = w_rank
        self.drop = nn.Dropout(p=0.5)          # present (and unused) in the reference too (:268)
        self.cell = cell
        self.u_ranks = u_ranks[0] if isinstance(u_ranks, list) and len(u_ranks) < 2 else u_ranks
        self.time_index, self.batch_index = (1, 0) if batch_first else (0, 1)
        cells, in_size = [], input_size
        for i, hidden_size in enumerate(hidden_layer_sizes):
            if recurrent_inits is not None:
                kwargs["recurrent_init"] = recurrent_inits[i]
            if hidden_inits is not None:
                kwargs["hidden_init"] = hidden_inits[i]
            cells.append(cell(in_size, hidden_size, w_rank=self.w_rank, u_ranks=self.u_ranks, **kwargs))
            in_size = hidden_size
        self.rnncells = nn.ModuleList(cells)

    def forward(self, x, need_sequence=True):
        """need_sequence=False (used by Net, which reads the last step only, V/models/vmlmf.py:354-355) lets the last
        layer skip writing its [B,T,H] output; the first return value may then be None."""
        last = []
        for i, cell in enumerate(self.rnncells):
            need_y = need_sequence or i + 1 < len(self.rnncells)
            canon = None
            if hasattr(cell, "plain_params"):            # plain VMLMF cell: parameter map fused into the call
                x, h, _ = vmlmf_plain_sequence(x, None, None, cell.plain_params(), batch_first=self.batch_first,
                                               need_y=need_y)
                last.append(h)
                continue
            if hasattr(cell, "canonical"):
                canon = cell.canonical()
            if canon is not None:
                x, h, _ = vmlmf_sequence(x, None, None, canon, batch_first=self.batch_first, need_y=need_y)
            else:
                nb = x.size(self.batch_index)
                h = x.new_zeros(nb, self.hidden_layer_sizes[i])
                c = x.new_zeros(nb, self.hidden_layer_sizes[i])
                outs = []
                for x_t in torch.unbind(x, self.time_index):
                    h, c = cell(x_t, (h, c))
                    outs.append(h)
                x = torch.stack(outs, self.time_index)
            last.append(h)
        return x, (last[0] if len(last) == 1 else torch.cat(last, -1))


class Net(nn.Module):
    """LSTM stack + 18-way linear head on the last timestep (V/models/vmlmf.py:319-355).

    Quirks kept on purpose: the head is 18-wide whatever the dataset (:345) and `self.cell` is an
    extra, never-called cell whose parameters appear in state_dict and never get a gradient
    (:348-350, read by the reference's unit tests)."""

    def __init__(self, input_size, layer_sizes=None, w_rank=None, u_rank=None, model=MyLSTM, cell=MyLSTMCell):
        super().__init__()
        if layer_sizes is None:
            layer_sizes = [32, 32]
        self.rnn = model(input_size, hidden_layer_sizes=layer_sizes, batch_first=True, w_rank=w_rank,
                         u_ranks=u_rank, cell=cell)
        self.lin = nn.Linear(layer_sizes[-1], 18)
        self.lin.bias.data.fill_(.1)
        self.lin.weight.data.normal_(0, .01)
        u = u_rank[-1] if cell == MyVMLMFCell else u_rank
        self.cell = cell(input_size, layer_sizes[-1], w_rank=w_rank, u_ranks=u)

    def forward(self, x):
        """x[B,T,I] -> logits[B,18].  The reference feeds y[:, -1] to the head (:354-355); the last
        layer's final hidden state is that same tensor, and using it lets backward skip reading a
        [B,T,H] upstream gradient that is zero everywhere but the last step."""
        _, h_last = self.rnn(x, need_sequence=False) if isinstance(self.rnn, MyLSTM) else self.rnn(x)
        top = self.rnn.hidden_layer_sizes[-1]
        if h_last.size(-1) != top:                      # several layers: the head reads the last layer's slice of the cat
            h_last = h_last[:, -top:]
        return head_linear(h_last, self.lin.weight, self.lin.bias).squeeze(1)
