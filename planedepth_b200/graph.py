"""CUDA-graph capture of one hot-path step (forward + backward).

The fused path is four kernel launches plus a handful of tiny bookkeeping ops; at B200 speeds the
Python/autograd launch overhead is as long as the kernels, so the step is captured once into a CUDA
graph over static buffers and replayed (blackwell guide: "capture launch-bound inner loops in CUDA
graphs")."""
from __future__ import annotations

from typing import Callable, Dict, List

import torch


class GraphedStep:
    """Captures ``fn()`` (which must read only from pre-allocated tensors and leave its results in
    tensors it returns) on a side stream after a few eager warm-up runs."""

    def __init__(self, fn: Callable[[], Dict[str, torch.Tensor]], warmup: int = 3):
        self.fn = fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fn()

    def replay(self) -> Dict[str, torch.Tensor]:
        self.graph.replay()
        return self.result


def make_step(hp, inputs, outputs, leaves: List[torch.Tensor], attach=None):
    """One training-style step of the path: pred_novel_images + compute_losses + backward into the
    given leaves.  Gradients are returned (not accumulated) so that replays overwrite them."""

    def step():
        out = dict(outputs)
        if attach is not None:
            attach(out)  # differentiable views of the plane geometry, re-derived per step like the decoder does
        losses = hp.process(inputs, out)
        grads = torch.autograd.grad(losses["loss/total_loss"], leaves, allow_unused=True)
        res = {"loss": losses["loss/total_loss"].detach()}
        for i, gr in enumerate(grads):
            if gr is not None:
                res["grad%d" % i] = gr
        return res

    return step
