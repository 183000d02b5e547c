"""Whole-step CUDA-graph capture.

The reference overlaps communication with backward using up to 15 polling host
threads per process; our exchange is stream ordered, so forward + backward +
fused push-pull/optimizer kernels can be captured ONCE and replayed with a
single launch ("CUDA streams and graphs instead of a tracing compiler").  The
cross-rank flag barriers keep their generation counters in device memory, so a
replay on every rank is self-synchronising.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from ..utils.timing import stamp


class GraphedStep:
    """Capture ``fn()`` (zero_grad + forward + backward + optimizer.step on
    static input buffers) after ``warmup`` eager runs; ``__call__`` replays it.

    ``pre_replay`` runs on the host before every replay (e.g.
    ``optimizer.refresh_hparams`` so lr schedules / Adam bias correction reach
    the kernels through their pinned hyper-parameter block).
    """

    def __init__(self, fn: Callable[[], torch.Tensor], warmup: int = 3, pre_replay: Optional[Callable] = None,
                 device=None):
        self.fn = fn
        self.pre_replay = pre_replay
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        try:   # warm-up/capture run on side streams by design; autograd's cross-stream syncs are captured too
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for i in range(max(warmup, 1)):
                out = fn()
                stamp("graph: eager warm-up %d enqueued" % i)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        stamp("graph: warm-up drained")
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            out = fn()
        self.out = out
        torch.cuda.synchronize(dev)
        stamp("graph: captured")

    def __call__(self):
        if self.pre_replay is not None:
            self.pre_replay()
        self.graph.replay()
        return self.out
