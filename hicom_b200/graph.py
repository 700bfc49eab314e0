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
