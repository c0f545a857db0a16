"""CUDA-graph capture of a whole criterion step (forward + backward).

The step is a chain of ~100 small kernels (EMA, heads, attention, fused InfoNCE, enqueue and
their backward); at BASELINE sizes each runs for microseconds, so an eager step is bound by
launch latency.  ``GraphedStep`` warms the step up, captures ONE CUDA graph of it and replays
that graph per step: no Python between kernels, no host syncs.

Requirements on the step function (all met by the moma_b200 modules):
  * static shapes, inputs read from static buffers;
  * the queue pointer lives on the device (``contrast.use_device_pointer()``), the host mirror is
    advanced with ``contrast.replayed(n)`` after each replay;
  * NCCL collectives inside the step are graph-capturable (NCCL >= 2.9);
  * run on a non-default stream from the start (``torch.cuda.set_stream``): autograd binds each
    leaf's gradient accumulation to the stream the leaf was first used on, and a capture must not
    touch the legacy default stream.
"""
from __future__ import annotations

from typing import Callable

import torch


class GraphedStep:
    def __init__(self, step_fn: Callable[[], torch.Tensor], contrast=None, rows_per_step: int = 0, warmup: int = 3):
        """step_fn() runs one full step (including backward) and returns the loss tensor."""
        self.contrast, self.rows = contrast, rows_per_step
        if contrast is not None:
            contrast.use_device_pointer()
        cur = torch.cuda.current_stream()
        if cur == torch.cuda.default_stream():
            raise RuntimeError("GraphedStep: make a side stream current first (torch.cuda.set_stream(torch.cuda.Stream())) "
                               "and run the step on it from its very first call")
        self.stream = cur
        for _ in range(warmup):
            step_fn()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.loss = step_fn()
        # the captured call did not execute, but its Python moved the host pointer once: undo that
        # and re-sync the device copy
        if contrast is not None:
            contrast.index = (contrast.index - rows_per_step) % contrast.K
            contrast._index_dev.fill_(contrast.index)

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        if self.contrast is not None:
            self.contrast.replayed(self.rows)
        return self.loss
