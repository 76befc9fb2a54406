"""Drop-in replacements for V/models/vmlmf_group.py: the group-structured VMLMF cell and its
no-vector-multiplication ablation.  Both reduce to the canonical recurrence with block-structured
hidden-side factors (packing.pack_group) and run on the same fused kernels as the plain cell."""
from __future__ import annotations

import torch
from torch import nn

from . import packing
from .functional import vmlmf_sequence

TIME_STEPS = 128
RECURRENT_MAX = pow(2, 1 / TIME_STEPS)
RECURRENT_MIN = pow(1 / 2, 1 / TIME_STEPS)


class _GroupCellBase(nn.Module):
    with_vm = True

    def __init__(self, input_size, hidden_size, w_rank=None, u_ranks=None, g=2, recurrent_init=None,
                 hidden_init=None):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.recurrent_init, self.hidden_init = recurrent_init, hidden_init
        self.w_rank, self.u_ranks, self.g = w_rank, u_ranks, g
        hg = int(hidden_size / g)
        L = nn.ParameterDict()
        if self.with_vm:                               # vm vectors first, as in the reference (:64-65)
            L["dia_x"] = nn.Parameter(0.1 * torch.randn([1, input_size]))
            L["dia_h"] = nn.Parameter(0.1 * torch.randn([1, hidden_size]))
        L["u_x"] = nn.Parameter(0.1 * torch.randn([input_size, w_rank]))
        L["v_x"] = nn.Parameter(0.1 * torch.randn([4 * hidden_size, w_rank]))
        for i in range(g):
            L[f"u_h_{i}"] = nn.Parameter(0.1 * torch.randn([g, hg, u_ranks[i]]))
            L[f"v_h_{i}"] = nn.Parameter(0.1 * torch.randn([g, u_ranks[i], 4 * hg]))
        for vec in ("x", "h"):
            L[f"bias_{vec}"] = nn.Parameter(torch.ones([1, 4 * hidden_size]))
        self.layers = L

    def __repr__(self):
        return (f"LSTM VM Group (input:{self.input_size}, hidden:{self.hidden_size}, "
                f"w_rank:{self.w_rank}, u_ranks:{self.u_ranks})")

    def canonical(self):
        return packing.pack_group(self.layers, self.g, with_vm=self.with_vm)

    def forward(self, x, hidden_states):
        h, c = hidden_states
        _, h1, c1 = vmlmf_sequence(x.unsqueeze(1), h, c, self.canonical(), batch_first=True)
        return h1, c1


class MyVMLMFCellg2(_GroupCellBase):
    """Group VMLMF cell (V/models/vmlmf_group.py:37-155): per gate the hidden matrix is g x g blocks,
    block (src (j+i)%g -> dst j) of rank u_ranks[i]; the true diagonal is replaced by dia_h."""
    with_vm = True


class MyVMLMFgCellg2(_GroupCellBase):
    """Ablation without the vector-multiplication terms (V/models/vmlmf_group.py:158-251)."""
    with_vm = False
