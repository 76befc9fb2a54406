"""CPU, world_size 2, gloo: the host-side data-parallel logic (batch sharding, flat gradient bucket, parameter
broadcast).  The compute on each rank is a plain torch stand-in (the fused kernels need a GPU); what is checked is
that the sharded, all-reduced gradients equal the single-process gradients of the global batch, the way
bench.py --gpus N relies on."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vmlmf_b200 as vb
from vmlmf_b200.parallel import GradBucket, broadcast_parameters, shard_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Toy(torch.nn.Module):
    """Parameters shaped like the live part of a VMLMF Net plus one parameter that never gets a gradient
    (the reference's dead Net.cell, V/models/vmlmf.py:348-350)."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(7)
        self.u = torch.nn.Parameter(0.1 * torch.randn(9, 4))
        self.v = torch.nn.Parameter(0.1 * torch.randn(18, 4))
        self.dead = torch.nn.Parameter(torch.randn(5))

    def forward(self, x):                      # [B,T,9] -> [B,18]
        return torch.tanh(x.mean(1) @ self.u) @ self.v.t()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _Toy()
        if rank == 1:                          # start from different weights: broadcast must fix it
            with torch.no_grad():
                net.u.add_(1.0)
        broadcast_parameters(net, src=0)
        g = torch.Generator().manual_seed(11)
        x = torch.randn(8, 6, 9, generator=g)
        y = torch.randint(0, 18, (8,), generator=g)
        bucket = GradBucket(net, average=True)
        for _ in range(2):                     # second pass exercises zero() and re-packing
            bucket.zero()
            torch.nn.functional.cross_entropy(net(shard_batch(x)), shard_batch(y)).backward()
            flat = bucket.all_reduce()
        assert net.dead.grad is None and all(p.grad is not None for p in (net.u, net.v))
        assert flat.numel() == net.u.numel() + net.v.numel()          # dead parameter left out of the bucket
        assert net.u.grad.data_ptr() == flat.data_ptr()               # after the all-reduce .grad is the bucket slice
        if rank == 0:
            torch.save({"u": net.u.grad.clone(), "v": net.v.grad.clone(), "w_u": net.u.detach().clone()}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_allreduce_equals_global_batch(tmp_path):
    out = str(tmp_path / "grads.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    ref = _Toy()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(8, 6, 9, generator=g)
    y = torch.randint(0, 18, (8,), generator=g)
    torch.nn.functional.cross_entropy(ref(x), y).backward()           # batch-mean loss over the GLOBAL batch
    assert torch.equal(got["w_u"], ref.u.detach())                    # rank 0's weights won the broadcast
    assert torch.allclose(got["u"], ref.u.grad, rtol=1e-5, atol=1e-7)
    assert torch.allclose(got["v"], ref.v.grad, rtol=1e-5, atol=1e-7)


def test_single_process_bucket_is_a_noop_collective():
    net = vb.Net(9, [16], w_rank=4, u_rank=[3], cell=vb.MyVMLMFCell)
    for p in net.rnn.parameters():
        p.grad = torch.ones_like(p)
    bucket = GradBucket(net)
    flat = bucket.all_reduce()
    assert flat.numel() == sum(p.numel() for p in net.rnn.parameters()) and bool((flat == 1).all())
    assert shard_batch(torch.arange(6)).tolist() == [0, 1, 2, 3, 4, 5]


# ---- LM data-parallel semantics (V/train_test/lm_test.py:140-153, :204): local loss = token-mean x LOCAL batch, gradients
# SUMMED over ranks (GradBucket(average=False)) = gradient of token-mean x GLOBAL batch; the clip uses the norm of the
# reduced gradient.  Stand-in model on CPU, the oracle's lm_nll_loss as the loss.

class _ToyLM(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(9)
        self.emb = torch.nn.Parameter(0.3 * torch.randn(11, 6))
        self.fc = torch.nn.Parameter(0.3 * torch.randn(11, 6))

    def forward(self, tok):                    # [T,B] -> scores [T*B, V]
        return torch.tanh(self.emb[tok]).reshape(-1, 6) @ self.fc.t()


def _lm_worker(rank, world, port, out):
    from oracle import vmlmf_oracle as vo
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _ToyLM()
        g = torch.Generator().manual_seed(3)
        tok = torch.randint(0, 11, (5, 8), generator=g)
        y = torch.randint(0, 11, (5, 8), generator=g)
        bucket = GradBucket(net, average=False)
        bucket.zero()
        vo.lm_nll_loss(net(shard_batch(tok, 1)), shard_batch(y, 1)).backward()      # token streams sharded along the batch axis
        flat = bucket.all_reduce()
        norm = flat.norm()
        coef = torch.clamp(5.0 / (norm + 1e-6), max=1.0)                            # clip_grad_norm_ on the REDUCED gradient
        if rank == 0:
            torch.save({"emb": net.emb.grad.clone(), "fc": net.fc.grad.clone(), "norm": norm, "coef": coef}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_lm_loss_sum_over_ranks_equals_global_batch(tmp_path):
    from oracle import vmlmf_oracle as vo
    out = str(tmp_path / "lm.pt")
    mp.spawn(_lm_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    ref = _ToyLM()
    g = torch.Generator().manual_seed(3)
    tok = torch.randint(0, 11, (5, 8), generator=g)
    y = torch.randint(0, 11, (5, 8), generator=g)
    vo.lm_nll_loss(ref(tok), y).backward()                                          # token-mean x GLOBAL batch on one process
    assert torch.allclose(got["emb"], ref.emb.grad, rtol=1e-5, atol=1e-7)
    assert torch.allclose(got["fc"], ref.fc.grad, rtol=1e-5, atol=1e-7)
    ref_norm = torch.cat([ref.emb.grad.reshape(-1), ref.fc.grad.reshape(-1)]).norm()
    assert torch.allclose(got["norm"], ref_norm, rtol=1e-5)
