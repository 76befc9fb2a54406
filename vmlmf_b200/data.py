"""Synthetic input pipeline for the benchmark / training harness.

The reference feeds its nets from `torch.utils.data.DataLoader`s over in-memory numpy windows
(V/utils/ucidataloader.py:107-125, batch 64 hard-coded; V/utils/oppdataloader.py:50-69): host tensors, one
`data.to(device)` per batch inside the loop (V/train_test/train.py:60).  Datasets are not available offline, so
this loader produces tensors of the same shapes ([B,128,9] / [B,24,77] fp32 windows + int64 labels, or [T,B] int64
token windows for the LM) from a seeded generator and moves them the way a production loader should:

  * `source="host"`   : a rotating pool of PINNED host batches; batch i+1 is copied host->device on a side stream
                        while step i computes (double buffered, two device staging buffers, events both ways).
                        This is the end-to-end path bench.py times (`e2e`).
  * `source="device"` : the pool lives in HBM (what a device-side augmentation / generation pipeline would hand
                        over); iteration returns the resident tensors, no copies.

Both yield `(x, y)` on the device; the staging buffers are stable across iterations, so a step captured as a CUDA
graph per staging buffer replays without an extra device-to-device copy.
"""
from __future__ import annotations

import torch


class SyntheticLoader:
    def __init__(self, make_batch, device, pool=4, source="host", seed=1234):
        """make_batch(generator) -> (x, y) CPU tensors of one per-GPU batch."""
        self.device = torch.device(device)
        self.source = source
        g = torch.Generator().manual_seed(seed)
        host = [make_batch(g) for _ in range(pool)]
        if source == "device":
            self.pool = [(x.to(self.device), y.to(self.device)) for x, y in host]
            self.host = None
        else:
            self.host = [(x.pin_memory(), y.pin_memory()) for x, y in host]
            self.pool = None
            self.copy_stream = torch.cuda.Stream(device=self.device)
            x0, y0 = self.host[0]
            self.stage = [(torch.empty_like(x0, device=self.device), torch.empty_like(y0, device=self.device)) for _ in range(2)]
            self.ready = [torch.cuda.Event() for _ in range(2)]
            self.freed = [torch.cuda.Event() for _ in range(2)]
        self.bytes_per_batch = sum(t.numel() * t.element_size() for t in host[0])

    def staging_buffers(self):
        """the device tensors iteration hands out (for graph capture per buffer)"""
        return self.pool if self.source == "device" else self.stage

    def _issue(self, i):
        s = i % 2
        hx, hy = self.host[i % len(self.host)]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.freed[s])
            self.stage[s][0].copy_(hx, non_blocking=True)
            self.stage[s][1].copy_(hy, non_blocking=True)
            self.ready[s].record(self.copy_stream)

    def batches(self, n):
        """yield n device batches; with source="host" the copy of batch i+1 overlaps the consumer's work on batch i.
        The consumer must finish enqueueing its use of a batch before asking for the next one (ordinary loop)."""
        if self.source == "device":
            for i in range(n):
                yield self.pool[i % len(self.pool)]
            return
        cur = torch.cuda.current_stream(self.device)
        for s in range(2):
            self.freed[s].record(cur)
        self._issue(0)
        for i in range(n):
            s = i % 2
            if i + 1 < n:
                self._issue(i + 1)
            cur.wait_event(self.ready[s])
            yield self.stage[s]
            self.freed[s].record(cur)
