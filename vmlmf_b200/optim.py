"""Flat-bucket optimizer steps: the reference's two update rules as single kernels over one
contiguous parameter / gradient buffer.

  FlatAdam     torch.optim.Adam(model.parameters(), lr)              V/train_test/train.py:47,65
  FlatClipSGD  clip_grad_norm_(params, max_norm); p -= lr * p.grad   V/train_test/lm_test.py:203-209

Both re-point every live parameter's `.data` and `.grad` at slices of two flat buffers (the gradient one is
`parallel.GradBucket`'s, i.e. the buffer the data-parallel all-reduce already sums), so a step is one launch
(two for the clipped SGD: the norm needs a grid-wide reduction) whatever the number of parameter tensors, and
zeroing the gradients is one fill.  Parameters that never receive a gradient (the reference's dead `Net.cell`,
V/models/vmlmf.py:348-350) are left alone, exactly as torch's optimizers skip `grad is None`.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .parallel import GradBucket


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _FlatOptimizer:
    def __init__(self, module_or_bucket, average=True):
        self.bucket = module_or_bucket if isinstance(module_or_bucket, GradBucket) else GradBucket(module_or_bucket, average)
        self.pflat = None

    def _build(self):
        """first step: the bucket learns which parameters are live from their .grad; parameters move into one buffer"""
        b = self.bucket
        b.pack()
        ref = b.params[0]
        if not ref.is_cuda:
            raise RuntimeError("vmlmf_b200.optim: parameters must live on a CUDA device (no CPU fallback)")
        self.pflat = torch.empty_like(b.flat)
        off = 0
        for p in b.params:
            view = self.pflat[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            off += p.numel()

    def zero_grad(self):
        self.bucket.zero()

    def all_reduce(self):
        return self.bucket.all_reduce()

    # ---- checkpointing: the flat buffers are private, so the state is exported per parameter NAME (like a module's
    # state_dict) and can be loaded into a freshly built optimizer of the same model after its first backward / step.
    # Note: parameters are re-pointed at slices of one flat buffer on the first step; module.to() / .cuda() afterwards
    # detaches them from it -- move the module first, then build the optimizer.
    _STATE_TENSORS = ()

    def _named_live(self):
        names = {id(p): n for n, p in self.bucket.module.named_parameters()}
        off = 0
        for p in self.bucket.params:
            yield names.get(id(p), str(off)), off, p.numel(), p.shape
            off += p.numel()

    def state_dict(self):
        if self.pflat is None:
            return {"state": {}, "hyper": self._hyper()}
        state = {}
        for name, off, n, shape in self._named_live():
            state[name] = {k: getattr(self, k)[off:off + n].view(shape).clone() for k in self._STATE_TENSORS}
        out = {"state": state, "hyper": self._hyper()}
        if getattr(self, "t", None) is not None:
            out["step"] = float(self.t.item())
        return out

    def load_state_dict(self, sd):
        if self.pflat is None:
            raise RuntimeError("vmlmf_b200.optim: run one backward pass and build the optimizer state (a step, or _build()) "
                               "before load_state_dict(): the live parameter set is taken from the first backward")
        for name, off, n, shape in self._named_live():
            if name in sd.get("state", {}):
                for k in self._STATE_TENSORS:
                    getattr(self, k)[off:off + n].copy_(sd["state"][name][k].reshape(-1))
        if "step" in sd and getattr(self, "t", None) is not None:
            self.t.fill_(sd["step"])
        for k, v in sd.get("hyper", {}).items():
            setattr(self, k, tuple(v) if isinstance(v, list) else v)

    def _hyper(self):
        return {}


class FlatAdam(_FlatOptimizer):
    """Adam with torch.optim.Adam's defaults and arithmetic (no amsgrad, no weight decay).  The step counter lives on
    the device so that a captured CUDA graph advances it on replay."""

    def __init__(self, module_or_bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, average=True):
        super().__init__(module_or_bucket, average)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.m = self.v = self.t = None

    P2P_MAX_BYTES = 2 << 20      # one-shot all-reduce: every rank reads world x bucket over NVLink
    _STATE_TENSORS = ("m", "v")

    def _hyper(self):
        return {"lr": self.lr, "betas": self.betas, "eps": self.eps}

    def _build(self):
        super()._build()
        self.m = torch.zeros_like(self.pflat)
        self.v = torch.zeros_like(self.pflat)
        self.t = torch.zeros((), dtype=torch.float32, device=self.pflat.device)
        b = self.bucket
        # data parallel with a peer-mapped bucket: fuse the all-reduce into this step (vmlmf_p2p_adam_step)
        if b.symm is not None and b.nbytes <= self.P2P_MAX_BYTES and b.symm.world_size <= 16:
            b.fused_reduce = True
            off = int(getattr(b.symm, "offset", 0))     # the bucket's byte offset inside the symmetric allocation
            self.peer_ptrs = torch.tensor([int(q) + off for q in b.symm.buffer_ptrs], dtype=torch.int64, device=self.pflat.device)

    @torch.no_grad()
    def step(self):
        if self.pflat is None:
            self._build()
        g = self.bucket.pack()
        self.t += 1
        b = self.bucket
        if b.fused_reduce:
            h = b.symm
            scale = 1.0 / h.world_size if b.average else 1.0
            with torch.cuda.device_of(g):
                st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                h.barrier(channel=0)                    # every rank's bucket is complete
                _lib.check(_lib.lib().vmlmf_p2p_adam_step(_ptr(self.pflat), _ptr(self.m), _ptr(self.v), _ptr(self.peer_ptrs),
                                                          h.world_size, g.numel(), scale, self.lr, self.betas[0], self.betas[1],
                                                          self.eps, _ptr(self.t), 0, st))
                h.barrier(channel=1)                    # every rank has read every bucket: the next backward may overwrite it
            return
        with torch.cuda.device_of(g):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().vmlmf_adam_step(_ptr(self.pflat), _ptr(g), _ptr(self.m), _ptr(self.v), g.numel(),
                                                  self.lr, self.betas[0], self.betas[1], self.eps, _ptr(self.t), 0, st))


class FlatClipSGD(_FlatOptimizer):
    """Global-norm clipping followed by plain SGD.  `step(lr)` returns the pre-clip gradient norm as a device scalar
    (what clip_grad_norm_ returns); gradients are rescaled in place like clip_grad_norm_ does."""

    def __init__(self, module_or_bucket, lr=1.0, max_norm=5.0, average=True):
        super().__init__(module_or_bucket, average)
        self.lr, self.max_norm = lr, max_norm
        self.ws = self.norm = None

    def _hyper(self):
        return {"lr": self.lr, "max_norm": self.max_norm}

    def _build(self):
        super()._build()
        n = self.pflat.numel()
        self.ws = self.pflat.new_empty(((_lib.lib().vmlmf_sgd_clip_workspace_bytes(n) + 3) // 4,))
        self.norm = self.pflat.new_zeros(())

    @torch.no_grad()
    def step(self, lr=None):
        if self.pflat is None:
            self._build()
        g = self.bucket.pack()
        with torch.cuda.device_of(g):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().vmlmf_sgd_clip_step(_ptr(self.pflat), _ptr(g), g.numel(),
                                                      self.lr if lr is None else lr, self.max_norm, 1, _ptr(self.norm),
                                                      _ptr(self.ws), st))
        return self.norm
