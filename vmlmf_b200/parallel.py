"""Batch-sharded data parallelism for the VMLMF nets: one process per GPU, replicated parameters,
one flat-bucket all-reduce of the (small) factor gradients per step.

The reference has no multi-device code at all (single `cuda:{gpu_id}`, V/train_test/main.py:109-110).
Sequences are independent in forward and backward, so the only exchange is the gradient sum
(SURVEY 8e).  Loss semantics to preserve: HAR uses a batch-MEAN cross entropy (train.py:63) =>
gradients are averaged over ranks; the LM loss is token-mean x batch size (lm_test.py:147,153) =>
callers that shard B pass average=False and scale the local loss themselves.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradBucket:
    """One flat buffer holding every live parameter gradient, so the all-reduce (and the flat optimizer steps of
    vmlmf_b200.optim) touch a single tensor.  Parameters that never receive a gradient (the reference's dead
    `Net.cell`, V/models/vmlmf.py:348-350) are left out -- found from the first backward.

    `zero()` drops the `.grad` tensors, so the next backward hands each parameter its freshly computed gradient
    without an accumulation kernel; `pack()` (called by `all_reduce()` and by the optimizers) copies them into the
    bucket with one multi-tensor launch and re-points `.grad` at the bucket slices, which is what callers see
    afterwards (reduced / clipped values)."""

    def __init__(self, module, average=True, symmetric=False):
        self.module = module
        self.average = average
        self.flat = None
        self.params = []
        self.views = []
        # symmetric=True: allocate the bucket as NVLink peer-mapped symmetric memory so that an optimizer can read every
        # rank's gradients directly (optim.FlatAdam fuses the all-reduce into its step); falls back to a plain buffer
        # when symmetric memory is not available (single process, no NVLink peer access)
        self.symmetric = symmetric
        self.symm = None          # torch symmetric-memory handle once rendezvoused
        self.fused_reduce = False # set by an optimizer that performs the reduction itself
        self._ids = set()

    def _build(self):
        self.params = [p for p in self.module.parameters() if p.grad is not None]
        if not self.params:
            raise RuntimeError("vmlmf_b200.parallel.GradBucket: no parameter has a gradient yet -- run a backward pass before "
                               "pack() / all_reduce() / optimizer.step() (the live parameter set is taken from the first backward)")
        self._ids = {id(p) for p in self.params}
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = None
        if self.symmetric and ref.is_cuda and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                flat = symm_mem.empty(n, dtype=ref.dtype, device=ref.device)
                self.symm = symm_mem.rendezvous(flat, dist.group.WORLD.group_name)
                flat.zero_()
                self.flat = flat
            except Exception as e:                      # no peer access / backend missing: plain bucket + NCCL
                import sys
                sys.stderr.write(f"vmlmf_b200.parallel: symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL\n")
                self.symm = None
        if self.flat is None:
            self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        off = 0
        self.views = []
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    @torch.no_grad()
    def pack(self):
        """gather the current .grad tensors into the bucket (no-op for those that already live there)"""
        if self.flat is None:
            self._build()
        else:
            # the live set was fixed by the first backward: a parameter that starts receiving gradients later (an unfrozen
            # layer, a conditional branch) would silently be neither reduced nor updated
            for name, p in self.module.named_parameters():
                if p.grad is not None and id(p) not in self._ids:
                    raise RuntimeError(f"vmlmf_b200.parallel.GradBucket: parameter '{name}' received a gradient but was not live at the "
                                       "first backward; build a new GradBucket / optimizer after changing which parameters train")
        dst, src = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                dst.append(v)
                src.append(p.grad)
            p.grad = v
        if dst:
            torch._foreach_copy_(dst, src)
        return self.flat

    def zero(self):
        """call instead of module.zero_grad()"""
        if self.flat is not None:
            for p in self.params:
                p.grad = None
        else:
            self.module.zero_grad(set_to_none=True)

    def all_reduce(self):
        """sum (or mean) the bucket over all ranks; packs first; no collective for a single process"""
        self.pack()
        if self.fused_reduce:                           # the optimizer step reads every rank's bucket itself
            return self.flat
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if self.average:
                self.flat.div_(dist.get_world_size())
        return self.flat

    @property
    def nbytes(self):
        return 0 if self.flat is None else self.flat.numel() * self.flat.element_size()


def broadcast_parameters(module, src=0):
    """make every rank start from rank `src`'s weights"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src)


def shard_batch(x, dim=0):
    """this rank's contiguous slice of a global batch along `dim`"""
    if not (dist.is_available() and dist.is_initialized()):
        return x
    w, r = dist.get_world_size(), dist.get_rank()
    n = x.size(dim)
    assert n % w == 0, "global batch must divide evenly over ranks"
    return x.narrow(dim, r * (n // w), n // w)
