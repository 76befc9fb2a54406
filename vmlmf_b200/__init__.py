"""vmlmf_b200 -- B200-native (sm_100a) implementation of the VMLMF compressed-LSTM hot path behind
the reference's own nn.Module API (snudm-starlab/VMLMF, V/models/{vmlmf,vmlmf_group,vmlmf_lm}.py)."""
from .functional import VmlmfSeqFunction, cross_entropy, head_linear, nll_loss, vmlmf_sequence  # noqa: F401
from . import graphs, parallel  # noqa: F401
from .optim import FlatAdam, FlatClipSGD  # noqa: F401
from .vmlmf import MyLSTM, MyLSTMCell, MyVMLMFCell, Net  # noqa: F401
from .vmlmf_group import MyVMLMFCellg2, MyVMLMFgCellg2  # noqa: F401
from .vmlmf_lm import LSTM, Embed, Linear, Model, MyVMLSTM, MyVMLSTMGroup  # noqa: F401

__version__ = "0.1.0"
