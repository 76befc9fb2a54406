"""ctypes binding of libvmlmf_b200.so (the C ABI declared in include/vmlmf_b200.h).

No CPU fallback exists by contract: if the shared library is missing, or a compute entry point is
called without a CUDA device, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvmlmf_b200.so")
ABI_VERSION = 5

PATH_R1, PATH_G, PATH_R1M, PATH_R2, PATH_R3 = 1, 2, 3, 4, 5
LARGE_PATHS = (PATH_G, PATH_R2, PATH_R3)      # regimes for shapes beyond the register-resident kernels


class Plan(C.Structure):
    """mirror of struct vmlmf_plan"""
    _fields_ = [("path", C.c_int), ("zx_pitch", C.c_int), ("z_pitch", C.c_int), ("xp_cols", C.c_int),
                ("fwd_workspace_bytes", C.c_longlong), ("bwd_workspace_bytes", C.c_longlong),
                ("gates_bytes", C.c_longlong), ("cs_bytes", C.c_longlong), ("reserved", C.c_int * 8)]


_P, _LL, _I, _F = C.c_void_p, C.c_longlong, C.c_int, C.c_float

# name -> argtypes; every function returns int except where noted
SIGNATURES = {
    "vmlmf_abi_version": [],
    "vmlmf_strerror": [_I],
    "vmlmf_seq_plan": [_I] * 6 + [C.POINTER(Plan)],
    "vmlmf_diag_fwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "vmlmf_diag_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "vmlmf_pack_plain_fwd": [_P] * 11 + [_I] * 4 + [_P],
    "vmlmf_pack_plain_bwd": [_P] * 14 + [_I] * 4 + [_P],
    "vmlmf_gemm_nt": [_P, _LL, _P, _LL, _P, _LL, _P, _I, _I, _I, _I, _P, _LL, _P],
    "vmlmf_gemm_tn": [_P, _LL, _P, _LL, _P, _LL, _I, _I, _LL, _I, _P, _LL, _P],
    "vmlmf_xproj_fwd": [_P, _LL, _LL, _P, _P, _I, _I, _I, _I, _I, _P],
    "vmlmf_seq_fwd": [C.POINTER(Plan), _P, _LL, _LL] + [_P] * 10 + [_P, _LL, _LL] + [_P] * 6 + [_I] * 6 + [_P],
    "vmlmf_seq_bwd": [C.POINTER(Plan), _P, _LL, _LL] + [_P] * 9 + [_P, _LL, _LL] + [_P] * 3 + [_P, _LL, _LL]
                     + [_P] * 2 + [_P, _LL, _LL] + [_P] * 10 + [_I] * 6 + [_P],
    "vmlmf_softmax_nll_workspace_bytes": [_LL, _I],                          # returns long long
    "vmlmf_softmax_nll_fwd": [_P, _LL, _P, _P, _P, _F, _P, _LL, _I, _P],
    "vmlmf_softmax_nll_bwd": [_P, _LL, _P, _P, _P, _F, _P, _LL, _LL, _I, _P],
    "vmlmf_head_fwd": [_P, _LL, _P, _P, _P, _I, _I, _I, _P],
    "vmlmf_head_bwd_workspace_bytes": [_I, _I, _I],                          # returns long long
    "vmlmf_head_bwd": [_P, _LL, _P, _P, _P, _LL, _P, _P, _P, _I, _I, _I, _P],
    "vmlmf_adam_step": [_P, _P, _P, _P, _LL, _F, _F, _F, _F, _P, _I, _P],
    "vmlmf_embed_dropout_fwd": [_P, _P, _P, _F, _P, _LL, _LL, _I, _I, _P],
    "vmlmf_p2p_adam_step": [_P, _P, _P, _P, _I, _LL, _F, _F, _F, _F, _F, _P, _I, _P],
    "vmlmf_sgd_clip_workspace_bytes": [_LL],                                 # returns long long
    "vmlmf_sgd_clip_step": [_P, _P, _LL, _F, _F, _I, _P, _P, _P],
}

_lib = None


def lib():
    """Load (once) and return the CDLL; raises RuntimeError when the extension is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"vmlmf_b200: CUDA extension not built ({LIB_PATH} missing). "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                "There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = args
            fn.restype = (C.c_char_p if name == "vmlmf_strerror" else
                          C.c_longlong if name.endswith("_workspace_bytes") else C.c_int)
        if handle.vmlmf_abi_version() != ABI_VERSION:
            raise RuntimeError("vmlmf_b200: libvmlmf_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError(lib().vmlmf_strerror(rc).decode())


def plan(T, B, I, H, RX, RH) -> Plan:
    p = Plan()
    rc = lib().vmlmf_seq_plan(T, B, I, H, RX, RH, C.byref(p))
    if rc == -2:
        # same exception type the reference ends up raising for H < I (V/models/vmlmf.py:92-94,117)
        raise TypeError(lib().vmlmf_strerror(rc).decode())
    check(rc)
    return p
