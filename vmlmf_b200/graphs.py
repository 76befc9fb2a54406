"""CUDA-graph capture of a whole training step (zero grads, forward, loss, backward, optimizer).

At the reference's own batch sizes (64 / 81 sequences) a step is ~60 kernel launches of a few microseconds each,
so the step time is launch overhead, not GPU work: replaying one captured graph removes it (cfg1, B=64:
1.32 ms -> 0.36 ms per step on B200).  The fused kernels are launched through ctypes on the current stream and
the library keeps no state, so they capture like any other kernel; tensor maps and workspace pointers are baked
into the graph, which is why inputs are copied into static buffers before every replay.
"""
from __future__ import annotations

import torch


class GraphedCallable:
    """Capture `fn()` (a closure over static tensors that launches only on the current stream) into one CUDA graph.
    `fn` is run `warmup` times eagerly on a side stream first; calling the object replays the graph and returns
    what `fn` returned at capture time (static output tensors)."""

    def __init__(self, fn, warmup=3):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out


class GraphedTrainStep:
    """step = GraphedTrainStep(net, opt, loss_fn, x_example, y_example);  loss = step(x, y)

    `opt` must be capturable (e.g. torch.optim.Adam(..., capturable=True) or fused=True with capturable=True).
    `zero_fn` replaces `opt.zero_grad(set_to_none=False)` (e.g. GradBucket.zero / FlatAdam.zero_grad); `after_backward` is called between backward and the optimizer step INSIDE the capture
    (leave None when it would issue a collective).

    Side effect to know about: the `warmup` eager iterations (CUDA requires them before a capture) are REAL training steps on
    the example batch -- parameters and optimizer state advance `warmup` times before the first replay.  Pass a batch you
    would train on anyway, or checkpoint / restore around construction when that matters."""

    def __init__(self, module, opt, loss_fn, x_example, y_example, warmup=3, zero_fn=None, after_backward=None,
                 static_inputs=False):
        self.module, self.opt, self.loss_fn = module, opt, loss_fn
        # static_inputs=True: the caller's tensors ARE the graph's input buffers (e.g. the two halves of a double-buffered
        # host->device pipeline, one graph each): calling with them replays without the device-to-device input copy
        self.static_x = x_example if static_inputs else x_example.clone()
        self.static_y = y_example if static_inputs else y_example.clone()
        if zero_fn is None:
            # .grad buffers created earlier on another stream would make autograd synchronise with that stream
            # during capture (cudaErrorStreamCaptureImplicit): let the side-stream warm-up create them
            opt.zero_grad(set_to_none=True)
        self.zero_fn = zero_fn or (lambda: opt.zero_grad(set_to_none=False))
        self.after_backward = after_backward
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up off the default stream, as capture requires
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._eager()

    def _eager(self):
        self.zero_fn()
        loss = self.loss_fn(self.module(self.static_x), self.static_y)
        loss.backward()
        if self.after_backward is not None:
            self.after_backward()
        self.opt.step()
        return loss.detach()

    def __call__(self, x, y):
        if x.data_ptr() != self.static_x.data_ptr():
            self.static_x.copy_(x, non_blocking=True)
        if y.data_ptr() != self.static_y.data_ptr():
            self.static_y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.static_loss
