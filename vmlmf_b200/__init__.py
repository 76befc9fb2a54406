"""vmlmf_b200 -- B200-native (sm_100a) implementation of the VMLMF compressed-LSTM hot path behind
the reference's own nn.Module API (snudm-starlab/VMLMF, V/models/{vmlmf,vmlmf_group,vmlmf_lm}.py)."""
from .functional import VmlmfSeqFunction, cross_entropy, head_linear, nll_loss, vmlmf_sequence  # noqa: F401
from . import data, graphs, parallel  # noqa: F401
from .optim import FlatAdam, FlatClipSGD  # noqa: F401
from .vmlmf import MyLSTM, MyLSTMCell, MyVMLMFCell, Net  # noqa: F401
from .vmlmf_group import MyVMLMFCellg2, MyVMLMFgCellg2  # noqa: F401
from .vmlmf_lm import LSTM, Embed, Linear, Model, MyVMLSTM, MyVMLSTMGroup  # noqa: F401

__version__ = "0.2.0"


def set_fast_tf32(enabled: bool) -> None:
    """OPTIONAL looser-precision mode of the tcgen05 GEMMs around the recurrence (LM vocabulary projection and its two
    backward GEMMs, the LM x projection, the time-parallel weight-gradient GEMMs): single-pass TF32 instead of the
    fp32-accurate 3xTF32 scheme.  Results then agree with the fp32 reference to about 1e-3 relative instead of 1e-5
    (bound and unchanged-argmax check: tests/test_gpu_tail.py::test_fast_tf32_mode).  The recurrence kernels themselves
    always run fp32-accurate.  Off by default; the library reads the switch on every call."""
    import os
    os.environ["VMLMF_FAST_TF32"] = "1" if enabled else "0"
