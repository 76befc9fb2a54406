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


def _state(t):
    """contiguous [B,H] state / state-gradient tensor whose base is 16-byte aligned (the kernels move float2 / float4);
    a contiguous view that starts at an odd float offset is copied."""
    if t is None:
        return None
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone()


def _tb_strides(t, batch_first):
    """(time stride, batch stride) in elements of a [B,T,F] / [T,B,F] tensor"""
    return (t.stride(1), t.stride(0)) if batch_first else (t.stride(0), t.stride(1))


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vmlmf_b200: tensors must live on a CUDA device (no CPU fallback exists)")
        if t is not None and t.dtype != torch.float32:
            raise RuntimeError("vmlmf_b200: fp32 only (the reference computes in fp32)")


def _xproj_on_tensor_cores(x, batch_first, T, B, I, RX, plan):
    """the x projection as one tcgen05 GEMM: time-major contiguous input (row t*B+b = the order zx is consumed in), a
    product big enough for 128 x 128 tiles, and an unpadded zx (RX % 4 == 0: the GEMM writes exactly RX columns)"""
    rows_uniform = x.stride(1) * B == x.stride(0) or T == 1          # row t*B+b lives at (t*B+b) * pitch
    return (plan.path in (_lib.PATH_R2, _lib.PATH_R3) and not batch_first and rows_uniform and I >= 128 and RX >= 32 and
            T * B >= 256 and plan.zx_pitch == RX)


def _seq_forward(x, h0, c0, canon, batch_first, need_grad, need_y):
    """xproj + the whole time loop of one layer through the C ABI.  Returns (y, hT, cT, saved) where `saved` is what
    _seq_backward needs (None when nothing is kept)."""
    _require_cuda(x, h0, c0, *canon)
    x = _row_contig(x)
    Ux, Vx, Dx, A, Bm, Dh, bias = [p.contiguous() for p in canon]
    if batch_first:
        B, T, I = x.shape
    else:
        T, B, I = x.shape
    H, RH = A.shape
    RX = Ux.shape[1]
    h0, c0 = _state(h0), _state(c0)
    plan = _lib.plan(T, B, I, H, RX, RH)
    lib = _lib.lib()
    new = x.new_empty
    # last-step-only callers (Net.forward) skip the [T,B,H] output when nothing reads it back: backward and the
    # generic regime take h_{t-1} from y
    if need_y or need_grad or plan.path in _lib.LARGE_PATHS:
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
            if _xproj_on_tensor_cores(x, batch_first, T, B, I, RX, plan):
                # wide time-major inputs (the LM layers: x[T,B,650] Ux[650,300]): ZX = X Ux on the tcgen05 GEMM.  Rows of 650
                # floats are not 16-byte multiples (TMA), so x is copied once into a pitch-padded buffer (46 MB at B=512
                # against a 7 GFLOP product that the SIMT fallback takes 0.55 ms for)
                ip = (I + 3) // 4 * 4
                pitch = x.stride(1)
                if pitch % 4 == 0 and pitch >= ip and x.data_ptr() % 16 == 0:
                    # already pitch-padded with zero pad columns (functional.embed_dropout): read in place
                    xa = x.as_strided((T * B, ip), (pitch, 1)) if pitch > I else x.reshape(T * B, I)
                else:
                    xa = x.new_zeros((T * B, ip))
                    xa[:, :I].copy_(x.reshape(T * B, I))
                uxt = x.new_zeros((RX, ip))
                uxt[:, :I].copy_(Ux.t())
                gemm_nt(xa, uxt, out=zx)
            else:
                _lib.check(lib.vmlmf_xproj_fwd(_ptr(x), xs_t, xs_b, _ptr(Ux), _ptr(zx), T, B, I, RX, plan.zx_pitch, st))
        with _timed("seq_fwd"):
            _lib.check(lib.vmlmf_seq_fwd(C.byref(plan), _ptr(x), xs_t, xs_b, _ptr(zx), _ptr(Ux), _ptr(Vx), _ptr(Dx),
                                         _ptr(A), _ptr(Bm), _ptr(Dh), _ptr(bias), _ptr(h0), _ptr(c0), _ptr(y), ys_t,
                                         ys_b, _ptr(hT), _ptr(cT), _ptr(gates), _ptr(cs), _ptr(z), _ptr(ws),
                                         T, B, I, H, RX, RH, st))
    saved = None
    if need_grad:
        saved = dict(tensors=(x, zx, Ux, Vx, Dx, A, Bm, Dh, h0, c0, y, gates, cs, z), plan=plan, batch_first=batch_first,
                     dims=(T, B, I, H, RX, RH))
    return y, hT, cT, saved


def _seq_backward(tensors, plan, batch_first, dims, dy, dhT, dcT, need_dx, need_dh0, need_dc0):
    """BPTT of one layer through the C ABI -> (dx, dh0, dc0, dUx, dVx, dDx, dA, dBm, dDh, dbias)."""
    x, zx, Ux, Vx, Dx, A, Bm, Dh, h0, c0, y, gates, cs, z = tensors
    T, B, I, H, RX, RH = dims
    bf = batch_first
    lib = _lib.lib()
    new = x.new_empty
    dy = None if dy is None else _row_contig(dy)
    dhT, dcT = _state(dhT), _state(dcT)
    dx = torch.empty_like(x, memory_format=torch.contiguous_format) if need_dx else None
    dh0 = new((B, H)) if (h0 is not None and need_dh0) else None
    dc0 = new((B, H)) if (c0 is not None and need_dc0) else None
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
    return dx, dh0, dc0, dUx, dVx, dDx, dA, dBm, dDh, dbias


def _stash(ctx, saved):
    ctx.save_for_backward(*saved["tensors"])
    ctx.plan, ctx.batch_first, ctx.dims = saved["plan"], saved["batch_first"], saved["dims"]
    ctx.set_materialize_grads(False)


class VmlmfSeqFunction(torch.autograd.Function):
    """One layer over a whole sequence in canonical parameters (Ux, Vx, Dx, A, Bm, Dh, bias)."""

    @staticmethod
    def forward(ctx, x, h0, c0, Ux, Vx, Dx, A, Bm, Dh, bias, batch_first, save=True, need_y=True):
        # grad mode is always off inside Function.forward and needs_input_grad ignores torch.no_grad():
        # the caller (vmlmf_sequence) decides whether anything has to be kept for backward
        need_grad = save and any(ctx.needs_input_grad)
        y, hT, cT, saved = _seq_forward(x, h0, c0, (Ux, Vx, Dx, A, Bm, Dh, bias), batch_first, need_grad, need_y)
        if saved is not None:
            _stash(ctx, saved)
        return y, hT, cT

    @staticmethod
    def backward(ctx, dy, dhT, dcT):
        need = ctx.needs_input_grad
        out = _seq_backward(ctx.saved_tensors, ctx.plan, ctx.batch_first, ctx.dims, dy, dhT, dcT, need[0], need[1], need[2])
        return (*out, None, None, None)


class PlainCellSeqFunction(torch.autograd.Function):
    """One MyVMLMFCell / MyVMLSTM layer over a whole sequence in the REFERENCE's parameters (u_x, u_h, v_x, v_h, b_x,
    b_h, dia_x, dia_h; V/models/vmlmf.py:56-69, vmlmf_lm.py:200-213).  Same kernels as VmlmfSeqFunction, with the map
    to canonical parameters and its chain rule as one launch each way (vmlmf_pack_plain_fwd / _bwd), so a train
    step has no per-parameter elementwise launches besides autograd's own accumulation into .grad."""

    @staticmethod
    def forward(ctx, x, h0, c0, u_x, u_h, v_x, v_h, b_x, b_h, dia_x, dia_h, batch_first, save=True, need_y=True):
        _require_cuda(x, u_x, u_h, v_x, v_h, b_x, b_h, dia_x, dia_h)
        u_x, u_h, v_x, v_h = u_x.contiguous(), u_h.contiguous(), v_x.contiguous(), v_h.contiguous()
        n_in, rx = u_x.shape
        hidden, rh = u_h.shape
        if hidden < n_in:                                   # same exception type as the reference (V/models/vmlmf.py:92-94,117)
            raise TypeError(_lib.lib().vmlmf_strerror(-2).decode())
        new = u_x.new_empty
        Dx, Dh, bias = new((4, n_in)), new((4, hidden)), new((4 * hidden,))
        with torch.cuda.device_of(u_x):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().vmlmf_pack_plain_fwd(_ptr(u_x), _ptr(v_x), _ptr(dia_x.contiguous()), _ptr(u_h), _ptr(v_h),
                                                       _ptr(dia_h.contiguous()), _ptr(b_x.contiguous()),
                                                       _ptr(b_h.contiguous()), _ptr(Dx), _ptr(Dh), _ptr(bias), n_in,
                                                       hidden, rx, rh, st))
        need_grad = save and any(ctx.needs_input_grad)
        y, hT, cT, saved = _seq_forward(x, h0, c0, (u_x, v_x, Dx, u_h, v_h, Dh, bias), batch_first, need_grad, need_y)
        if saved is not None:
            _stash(ctx, saved)
            ctx.shapes = (b_x.shape, b_h.shape, dia_x.shape, dia_h.shape)
        return y, hT, cT

    @staticmethod
    def backward(ctx, dy, dhT, dcT):
        need = ctx.needs_input_grad
        tensors = ctx.saved_tensors
        dx, dh0, dc0, dUx, dVx, dDx, dA, dBm, dDh, dbias = _seq_backward(tensors, ctx.plan, ctx.batch_first, ctx.dims,
                                                                         dy, dhT, dcT, need[0], need[1], need[2])
        u_x, v_x, u_h, v_h = tensors[2], tensors[3], tensors[5], tensors[6]
        _, _, n_in, hidden, rx, rh = ctx.dims
        new = u_x.new_empty
        ddia_x, ddia_h, db_h = new((n_in,)), new((hidden,)), new((4 * hidden,))
        with torch.cuda.device_of(u_x):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().vmlmf_pack_plain_bwd(_ptr(u_x), _ptr(v_x), _ptr(u_h), _ptr(v_h), _ptr(dDx), _ptr(dDh),
                                                       _ptr(dUx), _ptr(dVx), _ptr(dA), _ptr(dBm), _ptr(ddia_x),
                                                       _ptr(ddia_h), _ptr(dbias), _ptr(db_h), n_in, hidden, rx, rh, st))
        sb_x, sb_h, sd_x, sd_h = ctx.shapes
        # b_x and b_h get two different tensors: sharing one would alias their .grad (see packing.pack_plain)
        return (dx, dh0, dc0, dUx, dA, dVx, dBm, dbias.view(sb_x), db_h.view(sb_h), ddia_x.view(sd_x), ddia_h.view(sd_h),
                None, None, None)


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
    # split-K partials: up to 16 splits to fill the SMs, more for long contractions (<= 1536 of K per split keeps the
    # tensor core's truncating accumulate below 2e-6), within 64 MB
    nsp = max(min(16, max(1, -(-k // 128))), min(128, -(-k // 1536)))
    nsp = max(1, min(nsp, (1 << 24) // max(1, m * n)))
    ws = a.new_empty((nsp * m * n,)) if m * n <= (1 << 24) else None
    with torch.cuda.device_of(a):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().vmlmf_gemm_nt(_ptr(a2), lda, _ptr(b2), ldb, _ptr(out), out.stride(0), _ptr(bias), m, n, k,
                                            1 if accumulate else 0, _ptr(ws), 0 if ws is None else ws.numel() * 4, st))
    return out


def gemm_tn(at, bt, out=None, accumulate=False):
    """out[M,N] (+)= at[K,M]^T @ bt[K,N] on the tcgen05 3xTF32 GEMM with MN-major operands (vmlmf_gemm_tn): the
    contraction runs over the ROWS of both inputs, which are read in place -- no transposed copies.  fp32-accurate."""
    _require_cuda(at, bt)
    (k, m), n = at.shape, bt.shape[1]
    a2, lda = _pad4(at)
    b2, ldb = _pad4(bt)
    if out is None:
        out = at.new_empty((m, n))
    nsp = max(min(32, max(1, -(-k // 128))), min(128, -(-k // 1536)))
    nsp = max(1, min(nsp, (1 << 24) // max(1, m * n)))
    ws = at.new_empty((nsp * m * n,)) if m * n <= (1 << 24) else None
    with torch.cuda.device_of(at):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().vmlmf_gemm_tn(_ptr(a2), lda, _ptr(b2), ldb, _ptr(out), out.stride(0), m, n, k,
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
        dw = gemm_tn(dy, x) if ctx.needs_input_grad[1] else None                            # dy^T x, operands read in place
        db = dy.sum(0) if ctx.needs_input_grad[2] else None
        return dx, dw, db


def linear_tc(x, w, b):
    return LinearTCFunction.apply(x, w, b)


class SoftmaxNLLFunction(torch.autograd.Function):
    """loss = scale * sum_r (logsumexp(scores[r]) - scores[r, labels[r]]): one read of the scores forward (online
    max/sum), one read + one write backward, sums in a fixed order (vmlmf_softmax_nll_fwd/_bwd)."""

    @staticmethod
    def forward(ctx, scores, labels, scale):
        _require_cuda(scores)
        if not labels.is_cuda:
            raise RuntimeError("vmlmf_b200: tensors must live on a CUDA device (no CPU fallback exists)")
        if scores.stride(1) != 1:
            scores = scores.contiguous()
        labels = labels.reshape(-1).to(torch.int64).contiguous()
        rows, ncls = scores.shape
        lib = _lib.lib()
        lse = scores.new_empty((rows,))
        loss = scores.new_empty(())
        ws = scores.new_empty(((lib.vmlmf_softmax_nll_workspace_bytes(rows, ncls) + 3) // 4,))
        with torch.cuda.device_of(scores):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.vmlmf_softmax_nll_fwd(_ptr(scores), scores.stride(0), _ptr(labels), _ptr(lse), _ptr(loss),
                                                 float(scale), _ptr(ws), rows, ncls, st))
        ctx.save_for_backward(scores, labels, lse)
        ctx.scale = float(scale)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        scores, labels, lse = ctx.saved_tensors
        rows, ncls = scores.shape
        dscores = torch.empty_like(scores, memory_format=torch.contiguous_format)
        dloss = dloss.contiguous()
        with torch.cuda.device_of(scores):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().vmlmf_softmax_nll_bwd(_ptr(scores), scores.stride(0), _ptr(labels), _ptr(lse),
                                                        _ptr(dloss), ctx.scale, _ptr(dscores), dscores.stride(0),
                                                        rows, ncls, st))
        return dscores, None, None


class EmbedDropoutFunction(torch.autograd.Function):
    """x = dropout(w[tokens]) (V/models/vmlmf_lm.py:48, :436) in one kernel, written into a buffer whose row pitch is a
    multiple of four floats: the returned [.., E] view is what the first layer's x-projection GEMM reads in place as its TMA
    operand.  Backward is the deterministic dense embedding backward of the masked upstream gradient."""

    @staticmethod
    def forward(ctx, tokens, w, p, training):
        _require_cuda(w)
        if not tokens.is_cuda:
            tokens = tokens.to(w.device)                  # the reference indexes a CUDA table with CPU indices (lm_test.py:200)
        tok = tokens.reshape(-1).to(torch.int64).contiguous()
        rows, (vocab, emb) = tok.numel(), w.shape
        pitch = (emb + 3) // 4 * 4
        w = w.contiguous()
        mask = None
        scale = 1.0
        if training and p > 0.0:
            mask = (torch.rand((rows, emb), device=w.device) >= p).to(torch.uint8)
            scale = 1.0 / (1.0 - p)
        buf = w.new_empty((rows, pitch))
        with torch.cuda.device_of(w):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().vmlmf_embed_dropout_fwd(_ptr(tok), _ptr(w), _ptr(mask), scale, _ptr(buf), pitch, rows, emb, vocab, st))
        ctx.save_for_backward(tok, mask)
        ctx.scale, ctx.vocab = scale, vocab
        return buf[:, :emb].view(*tokens.shape, emb) if pitch == emb else buf.view(*tokens.shape, pitch)[..., :emb]

    @staticmethod
    def backward(ctx, dout):
        tok, mask = ctx.saved_tensors
        g = dout.reshape(tok.numel(), -1)
        if mask is not None:
            g = g * mask * ctx.scale
        dw = torch.ops.aten.embedding_dense_backward(g.contiguous(), tok, ctx.vocab, -1, False)
        return None, dw, None, None


def embed_dropout(tokens, w, p=0.0, training=False):
    return EmbedDropoutFunction.apply(tokens, w, float(p), bool(training))


def cross_entropy(logits, target):
    """F.cross_entropy(logits[N,C], target[N]) with mean reduction -- the HAR training loss (V/train_test/train.py:63)."""
    return SoftmaxNLLFunction.apply(logits, target, 1.0 / logits.shape[0])


def nll_loss(scores, y):
    """The LM loss of V/train_test/lm_test.py:140-153: mean over tokens of -log softmax(scores)[y], times the batch
    size y.size(1); computed through log-sum-exp, so large scores do not overflow the way the reference's exp() does."""
    return SoftmaxNLLFunction.apply(scores, y, float(y.size(1)) / scores.shape[0])


class HeadLinearFunction(torch.autograd.Function):
    """out = h W^T + b for the small classifier head of Net (nn.Linear(H_last, 18), V/models/vmlmf.py:345-347):
    one kernel forward, two backward (per-block partial dW / db, then a fixed-order reduce)."""

    @staticmethod
    def forward(ctx, h, w, b):
        _require_cuda(h, w, b)
        if h.stride(1) != 1:
            h = h.contiguous()
        w = w.contiguous()
        B, K = h.shape
        N = w.shape[0]
        out = h.new_empty((B, N))
        with torch.cuda.device_of(h):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().vmlmf_head_fwd(_ptr(h), h.stride(0), _ptr(w), _ptr(b), _ptr(out), B, K, N, st))
        ctx.save_for_backward(h, w)
        ctx.has_bias = b is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        h, w = ctx.saved_tensors
        B, K = h.shape
        N = w.shape[0]
        lib = _lib.lib()
        dout = dout.contiguous()
        dh = h.new_empty((B, K)) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w)
        db = h.new_empty((N,)) if ctx.has_bias else None
        ws = h.new_empty(((lib.vmlmf_head_bwd_workspace_bytes(B, K, N) + 3) // 4,))
        with torch.cuda.device_of(h):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.vmlmf_head_bwd(_ptr(h), h.stride(0), _ptr(w), _ptr(dout), _ptr(dh), K, _ptr(dw), _ptr(db),
                                          _ptr(ws), B, K, N, st))
        return dh, dw, db


def head_linear(h, w, b):
    """nn.Linear on CUDA fp32: the narrow-head kernels (out_features <= 32, in_features <= 1024: the reference's 18-way
    head at every hidden size it ships) or, for wider shapes, the tcgen05 GEMMs of linear_tc.  Host tensors (module
    construction / state_dict tests) go to F.linear; no library GEMM runs on the device path."""
    if h.is_cuda and h.dim() == 2 and h.dtype == torch.float32:
        if w.shape[0] <= 32 and w.shape[1] <= 1024:
            return HeadLinearFunction.apply(h, w, b)
        return linear_tc(h, w, b)
    return torch.nn.functional.linear(h, w, b)


def vmlmf_plain_sequence(x, h0, c0, params, batch_first=True, need_y=True):
    """Run one plain VMLMF layer given the reference's eight parameters (u_x, u_h, v_x, v_h, b_x, b_h, dia_x, dia_h)."""
    save = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, h0, c0, *params))
    return PlainCellSeqFunction.apply(x, h0, c0, *params, batch_first, save, need_y)


def vmlmf_sequence(x, h0, c0, canon, batch_first=True, need_y=True):
    """Run one VMLMF layer over a whole sequence.  canon = (Ux,Vx,Dx,A,Bm,Dh,bias).  Returns (y, hT, cT); with
    need_y=False the caller declares it reads only (hT, cT) and y may come back as None (inference)."""
    save = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, h0, c0, *canon))
    return VmlmfSeqFunction.apply(x, h0, c0, *canon, batch_first, save, need_y)
