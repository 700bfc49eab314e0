"""CUDA-graph capture of the compressor for fixed shapes.

The compressor is a short DAG of ~25 kernels on two streams (plus one NCCL all-gather when frame-sharded); for small
batches its run time is dominated by launch latency.  ``GraphedCompressor`` captures ``forward_batched`` (or
``dist.forward_frame_sharded``) once into a ``torch.cuda.CUDAGraph`` over static input buffers and replays it — CUDA
streams and graphs instead of a tracing compiler.  Fill ``.frames_feature`` / ``.frames_embed`` / ``.guide_embed``
in place (or pass tensors to ``__call__``, which copies them in) and read ``.output`` after ``replay()``.
"""
from __future__ import annotations

from typing import Optional

import torch


class GraphedCompressor:
    def __init__(self, projector, frames_feature: torch.Tensor, frames_embed: Optional[torch.Tensor],
                 guide_embed: Optional[torch.Tensor], modal: str = "video", frame_shard_t0: Optional[int] = None,
                 group=None, warmup: int = 3, image_newline: Optional[torch.Tensor] = None, adopt_inputs: bool = False):
        """``adopt_inputs``: use the given tensors themselves as the static input buffers instead of cloning them (a
        512-video batch is 55 GB; the caller then refills them in place between replays)."""
        self.projector = projector
        self.modal = modal
        self.image_newline = image_newline  # read in place at replay (a parameter of the parent model)
        keep = (lambda t: t) if adopt_inputs else (lambda t: t.clone())
        self.frames_feature = keep(frames_feature)
        self.frames_embed = None if frames_embed is None else keep(frames_embed)
        self.guide_embed = None if guide_embed is None else keep(guide_embed)
        self._t0 = frame_shard_t0
        self._group = group
        dev = frames_feature.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        from . import ops
        n0 = ops.kernel_launch_count()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.output = self._run()
        self.kernels_per_replay = ops.kernel_launch_count() - n0

    def _run(self):
        if self._t0 is not None:
            from . import dist as hdist
            return hdist.forward_frame_sharded(self.projector, self.frames_feature, self.frames_embed,
                                               self.guide_embed, t0=self._t0, group=self._group, modal=self.modal)
        return self.projector.forward_batched(self.frames_feature, self.frames_embed, self.guide_embed, self.modal,
                                              self.image_newline)

    def replay(self):
        self.graph.replay()
        return self.output

    def __call__(self, frames_feature, frames_embed=None, guide_embed=None):
        self.frames_feature.copy_(frames_feature, non_blocking=True)
        if self.frames_embed is not None:
            self.frames_embed.copy_(frames_embed, non_blocking=True)
        if self.guide_embed is not None:
            self.guide_embed.copy_(guide_embed, non_blocking=True)
        return self.replay()


def graphed_training_forward(projector, frames_feature: torch.Tensor, frames_embed: Optional[torch.Tensor],
                             guide_embed: Optional[torch.Tensor], modal: str = "video",
                             image_newline: Optional[torch.Tensor] = None, warmup: int = 3):
    """The differentiable ``forward_batched`` (hicom_b200.autograd) with its forward AND backward captured in CUDA graphs
    for fixed shapes (``torch.cuda.make_graphed_callables`` over the library's launches).  One training step of the
    projector is ~150 launches of mostly small kernels, i.e. launch-bound when issued eagerly (4.2 ms for 8 videos x 16
    frames against 2.5 ms of kernel time); replayed it runs at kernel speed.

    Returns ``fn(frames_feature[, frames_embed][, guide_embed]) -> tokens`` taking the tensor arguments that were not
    ``None`` here, with the same shapes and dtypes; ``tokens.backward()`` fills the ``.grad`` of the projector's
    parameters (and of ``image_newline`` when it requires grad) exactly as the eager path does.  Call it on the stream
    and device it was built on, and build it BEFORE the module has run an eager backward on the default stream (PyTorch
    ties gradient accumulation to the stream of the first backward)."""
    given = [("X", frames_feature), ("E", frames_embed), ("G", guide_embed)]
    names = [n for n, t in given if t is not None]
    sample = tuple(t for _, t in given if t is not None)

    class _Step(torch.nn.Module):
        def __init__(self, proj, newline):
            super().__init__()
            self.proj = proj
            self.newline = newline  # a parameter of the parent model (or None): read in place

        def forward(self, *tensors):
            kw = dict(zip(names, tensors))
            return self.proj.forward_batched(kw.get("X"), kw.get("E"), kw.get("G"), modal, self.newline)

    step = _Step(projector, image_newline)
    return torch.cuda.make_graphed_callables(step, sample, num_warmup_iters=warmup, allow_unused_input=True)
