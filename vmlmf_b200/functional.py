"""torch.autograd.Function over the C ABI: one fused call per layer per direction.

    y, hT, cT = vmlmf_sequence(x, h0, c0, Ux, Vx, Dx, A, Bm, Dh, bias, batch_first)

replaces the reference's python time loop and everything inside it
(V/models/vmlmf.py:308-310, V/models/vmlmf_lm.py:277-279).  PyTorch owns every buffer (inputs,
outputs, saved activations, scratch); the library only launches kernels on the current stream.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


# bench.py sets this to a list; each C-ABI call then appends (name, start_event, end_event) recorded on the
# launching stream.  None (the default) adds no work.
EVENT_LOG = None


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if EVENT_LOG is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if EVENT_LOG is not None:
            self.b.record()
            EVENT_LOG.append((self.name, self.a, self.b))
        return False


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _row_contig(t):
    return t if t.stride(-1) == 1 else t.contiguous()


def _tb_strides(t, batch_first):
    """(time stride, batch stride) in elements of a [B,T,F] / [T,B,F] tensor"""
    return (t.stride(1), t.stride(0)) if batch_first else (t.stride(0), t.stride(1))


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vmlmf_b200: tensors must live on a CUDA device (no CPU fallback exists)")
        if t is not None and t.dtype != torch.float32:
            raise RuntimeError("vmlmf_b200: fp32 only (the reference computes in fp32)")


class VmlmfSeqFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h0, c0, Ux, Vx, Dx, A, Bm, Dh, bias, batch_first, save=True, need_y=True):
        _require_cuda(x, h0, c0, Ux, Vx, Dx, A, Bm, Dh, bias)
        x = _row_contig(x)
        params = [p.contiguous() for p in (Ux, Vx, Dx, A, Bm, Dh, bias)]
        Ux, Vx, Dx, A, Bm, Dh, bias = params
        if batch_first:
            B, T, I = x.shape
        else:
            T, B, I = x.shape
        H, RH = A.shape
        RX = Ux.shape[1]
        h0 = None if h0 is None else h0.contiguous()
        c0 = None if c0 is None else c0.contiguous()
        plan = _lib.plan(T, B, I, H, RX, RH)
        lib = _lib.lib()
        new = x.new_empty
        # grad mode is always off inside Function.forward and needs_input_grad ignores torch.no_grad():
        # the caller (vmlmf_sequence) decides whether anything has to be kept for backward
        need_grad = save and any(ctx.needs_input_grad)
        # last-step-only callers (Net.forward) skip the [T,B,H] output when nothing reads it back: backward and the
        # generic regime take h_{t-1} from y
        if need_y or need_grad or plan.path == _lib.PATH_G:
            y = new((B, T, H)) if batch_first else new((T, B, H))
        else:
            y = None
        hT, cT = new((B, H)), new((B, H))
        zx = new((T * B, plan.zx_pitch))
        if need_grad:
            gates, cs, z = new((plan.gates_bytes // 4,)), new((plan.cs_bytes // 4,)), new((T * B, plan.z_pitch))
        else:
            gates = cs = z = None
        ws = new((plan.fwd_workspace_bytes + 3) // 4) if plan.fwd_workspace_bytes else None
        xs_t, xs_b = _tb_strides(x, batch_first)
        ys_t, ys_b = _tb_strides(y, batch_first) if y is not None else (0, 0)
        with torch.cuda.device_of(x):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            with _timed("xproj_fwd"):
                _lib.check(lib.vmlmf_xproj_fwd(_ptr(x), xs_t, xs_b, _ptr(Ux), _ptr(zx), T, B, I, RX, plan.zx_pitch, st))
            with _timed("seq_fwd"):
              _lib.check(lib.vmlmf_seq_fwd(C.byref(plan), _ptr(x), xs_t, xs_b, _ptr(zx), _ptr(Ux), _ptr(Vx), _ptr(Dx),
                                         _ptr(A), _ptr(Bm), _ptr(Dh), _ptr(bias), _ptr(h0), _ptr(c0), _ptr(y), ys_t,
                                         ys_b, _ptr(hT), _ptr(cT), _ptr(gates), _ptr(cs), _ptr(z), _ptr(ws),
                                         T, B, I, H, RX, RH, st))
        if need_grad:
            ctx.save_for_backward(x, zx, Ux, Vx, Dx, A, Bm, Dh, h0, c0, y, gates, cs, z)
            ctx.plan = plan
            ctx.batch_first = batch_first
            ctx.dims = (T, B, I, H, RX, RH)
            ctx.set_materialize_grads(False)
        return y, hT, cT

    @staticmethod
    def backward(ctx, dy, dhT, dcT):
        x, zx, Ux, Vx, Dx, A, Bm, Dh, h0, c0, y, gates, cs, z = ctx.saved_tensors
        T, B, I, H, RX, RH = ctx.dims
        plan, bf = ctx.plan, ctx.batch_first
        lib = _lib.lib()
        new = x.new_empty
        dy = None if dy is None else _row_contig(dy)
        dhT = None if dhT is None else dhT.contiguous()
        dcT = None if dcT is None else dcT.contiguous()
        need = ctx.needs_input_grad
        dx = torch.empty_like(x, memory_format=torch.contiguous_format) if need[0] else None
        dh0 = new((B, H)) if (h0 is not None and need[1]) else None
        dc0 = new((B, H)) if (c0 is not None and need[2]) else None
        dUx, dVx, dDx, dA, dBm, dDh = (torch.empty_like(p) for p in (Ux, Vx, Dx, A, Bm, Dh))
        dbias = new((4 * H,))
        ws = new((plan.bwd_workspace_bytes + 3) // 4) if plan.bwd_workspace_bytes else None
        xs = _tb_strides(x, bf)
        ys = _tb_strides(y, bf)
        dys = _tb_strides(dy, bf) if dy is not None else (0, 0)
        dxs = _tb_strides(dx, bf) if dx is not None else (0, 0)
        with torch.cuda.device_of(x):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            with _timed("seq_bwd"):
              _lib.check(lib.vmlmf_seq_bwd(C.byref(plan), _ptr(x), xs[0], xs[1], _ptr(zx), _ptr(Ux), _ptr(Vx), _ptr(Dx),
                                         _ptr(A), _ptr(Bm), _ptr(Dh), _ptr(h0), _ptr(c0), _ptr(y), ys[0], ys[1],
                                         _ptr(gates), _ptr(cs), _ptr(z), _ptr(dy), dys[0], dys[1], _ptr(dhT),
                                         _ptr(dcT), _ptr(dx), dxs[0], dxs[1], _ptr(dh0), _ptr(dc0), _ptr(dUx),
                                         _ptr(dVx), _ptr(dDx), _ptr(dA), _ptr(dBm), _ptr(dDh), _ptr(dbias), _ptr(ws),
                                         T, B, I, H, RX, RH, st))
        return dx, dh0, dc0, dUx, dVx, dDx, dA, dBm, dDh, dbias, None, None, None


class DiagCorrFunction(torch.autograd.Function):
    """D[k,j] = dia[j] - sum_r u[j,r] v[kH+j,r]  and its chain rule, one kernel each way (K0 / K5).
    The reference recomputes this inside every timestep (V/models/vmlmf.py:102-106)."""

    @staticmethod
    def forward(ctx, u, v, dia):
        _require_cuda(u, v, dia)
        u, v, d1 = u.contiguous(), v.contiguous(), dia.reshape(-1).contiguous()
        n, r = u.shape
        hidden = v.shape[0] // 4
        out = u.new_empty((4, n))
        with torch.cuda.device_of(u):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().vmlmf_diag_fwd(_ptr(u), _ptr(v), _ptr(d1), _ptr(out), n, hidden, r, st))
        ctx.save_for_backward(u, v)
        ctx.dia_shape = dia.shape
        return out

    @staticmethod
    def backward(ctx, d_out):
        u, v = ctx.saved_tensors
        n, r = u.shape
        hidden = v.shape[0] // 4
        d_out = d_out.contiguous()
        du, dv, ddia = torch.empty_like(u), torch.empty_like(v), u.new_empty((n,))
        with torch.cuda.device_of(u):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().vmlmf_diag_bwd(_ptr(u), _ptr(v), _ptr(d_out), _ptr(du), _ptr(dv), _ptr(ddia), n, hidden, r, st))
        return du, dv, ddia.view(ctx.dia_shape)


def diag_correction(u, v, dia):
    """[4,n] vector-multiplication coefficients of one side (n = rows of u)."""
    return DiagCorrFunction.apply(u, v, dia)


def _pad4(t):
    """copy of a 2-D tensor whose row pitch is a multiple of 4 floats and whose base is 16-byte aligned (TMA operand)"""
    rows, cols = t.shape
    if t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0:
        return t, t.stride(0)
    ld = (cols + 3) // 4 * 4
    buf = t.new_zeros((rows, ld))
    buf[:, :cols].copy_(t)
    return buf, ld


def gemm_nt(a, b, bias=None, out=None, accumulate=False):
    """out[M,N] (+)= a[M,K] @ b[N,K]^T (+ bias) on the tcgen05 3xTF32 GEMM (vmlmf_gemm_nt); fp32-accurate."""
    _require_cuda(a, b, bias)
    (m, k), n = a.shape, b.shape[0]
    a2, lda = _pad4(a)
    b2, ldb = _pad4(b)
    if out is None:
        out = a.new_empty((m, n))
    ws = a.new_empty((min(16, max(1, -(-k // 128))) * m * n,)) if m * n <= (1 << 24) else None
    with torch.cuda.device_of(a):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().vmlmf_gemm_nt(_ptr(a2), lda, _ptr(b2), ldb, _ptr(out), out.stride(0), _ptr(bias), m, n, k,
                                            1 if accumulate else 0, _ptr(ws), 0 if ws is None else ws.numel() * 4, st))
    return out


class LinearTCFunction(torch.autograd.Function):
    """y = x w^T + b with all three GEMMs (forward, dX, dW) on the tensor-core kernel.
    Replaces `torch.addmm(self.b, x, self.w.t())` of the LM head (V/models/vmlmf_lm.py:357) and its autograd."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return gemm_nt(x, w, b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = gemm_nt(dy, w.t().contiguous()) if ctx.needs_input_grad[0] else None          # [M,N] x [K,N]^T
        dw = gemm_nt(dy.t().contiguous(), x.t().contiguous()) if ctx.needs_input_grad[1] else None   # [N,M] x [K,M]^T
        db = dy.sum(0) if ctx.needs_input_grad[2] else None
        return dx, dw, db


def linear_tc(x, w, b):
    return LinearTCFunction.apply(x, w, b)


def vmlmf_sequence(x, h0, c0, canon, batch_first=True, need_y=True):
    """Run one VMLMF layer over a whole sequence.  canon = (Ux,Vx,Dx,A,Bm,Dh,bias).  Returns (y, hT, cT); with
    need_y=False the caller declares it reads only (hT, cT) and y may come back as None (inference)."""
    save = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, h0, c0, *canon))
    return VmlmfSeqFunction.apply(x, h0, c0, *canon, batch_first, save, need_y)
