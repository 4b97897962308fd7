"""CUDA-graph capture of a whole Reduce + Connect step (forward + backward).

Every entry point of ``libtgp_b200.so`` only enqueues kernels on the caller's stream (no allocation, no
synchronisation, ``include/tgp_b200.h``), and the dense path has no data-dependent output size, so a training step on
static buffers can be captured once and replayed with a single ``cudaGraphLaunch``: the host cost of a step drops from
~10 Python / allocator / launch hops to one call (a 0.3 ms step is otherwise launch-bound on an ordinary host).

    step = GraphedStep(lambda: loss_and_backward(x, adj, s))   # x, adj, s: static CUDA tensors
    x.copy_(new_x); step.replay()                              # results land in the tensors `fn` returned
"""
from __future__ import annotations

import gc
from typing import Any, Callable

import torch

from . import _lib as L


class GraphedStep:
    """Capture ``fn()`` into one CUDA graph.  ``fn`` must read its inputs from tensors that stay alive (their
    storage is baked into the graph), may call ``torch.autograd.backward`` (set ``.grad = None`` first so the
    gradient buffers are allocated inside the capture and stay static), and must not synchronise.  Outputs of an
    earlier eager call of the same step must be released before capturing: they keep that call's autograd graph
    (and its AccumulateGrad nodes, bound to the default stream) alive, which CUDA rejects inside a capture."""

    def __init__(self, fn: Callable[[], Any], warmup: int = 3):
        if not torch.cuda.is_available():
            raise RuntimeError("tgp_b200.GraphedStep needs a CUDA device (no CPU fallback by design)")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        # Autograd graphs of earlier eager calls can sit in reference cycles until the cyclic collector runs; their
        # AccumulateGrad nodes are bound to the stream they were created on (usually the legacy default stream) and
        # would be reused inside the capture (cudaErrorStreamCaptureImplicit).  Collect before warm-up and capture.
        gc.collect()
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):  # lazy initialisation (TMA descriptors, smem opt-ins) outside capture
                out = fn()
                del out
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gc.collect()
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.kernel_launches()
        with torch.cuda.graph(self.graph):
            self.outputs = fn()
        self.kernels_per_replay = L.kernel_launches() - n0

    def replay(self):
        self.graph.replay()
        return self.outputs
