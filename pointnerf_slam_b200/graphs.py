"""CUDA-graph replay of one tracking / mapping iteration.

An iteration of the reference's optimisation loops (src/Tracker.py:127-165,
src/Mapper.py:588-653) is ~45 kernel launches plus the autograd bookkeeping
between them; on a B200 the kernels themselves finish in 1-2 ms, so eager
launching leaves the GPU waiting for the host.  The shapes do not change from
one iteration to the next (same pixel count, same sample count, same grids),
which is exactly what a CUDA graph needs: capture the iteration ONCE -- the
very same API calls, kernels, autograd backward and (on several GPUs) NCCL
collectives -- and replay it.

    step = GraphedStep(one_iteration, generators=[gen])   # one_iteration() -> loss (or a tuple of tensors)
    for _ in range(n_iters):
        loss = step()          # replays; `loss` is the same static tensor every time
        optimizer.step()       # or put the (capturable) optimizer step inside one_iteration

Rules of CUDA graphs apply: the callable must read its inputs from tensors that
stay at the same address (copy new data INTO them), must not synchronise with the
host (no .item(), no boolean-mask indexing), and gradients it leaves in ``.grad``
are overwritten by the next replay.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch


class GraphedStep:
    def __init__(self, fn: Callable[[], object], generators: Iterable[torch.Generator] = (), warmup: int = 3,
                 stream: Optional[torch.cuda.Stream] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device")
        self.fn = fn
        self.graph = None
        self.out = None
        self.launches = 0
        self._capture(list(generators), warmup, stream)

    def _capture(self, generators, warmup, stream):
        from . import _lib as L
        side = stream or torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up off the default stream, as torch.cuda.graph requires
            for _ in range(max(1, warmup)):
                self.fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        for gen in generators:
            g.register_generator_state(gen)
        n0 = L.lib().pn_launch_count()
        with torch.cuda.graph(g):
            self.out = self.fn()
        self.launches = int(L.lib().pn_launch_count() - n0)   # library kernels per replay
        self.graph = g

    def __call__(self):
        self.graph.replay()
        return self.out

    def release(self) -> None:
        """Drop the graph (required before a process group whose collectives were captured goes away)."""
        self.graph = None
        self.out = None
