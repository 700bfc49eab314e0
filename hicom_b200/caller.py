"""Batched replacement for the per-sample loop of ``HIComMetaForCausalLM.encode_images_or_videos``
(hicom/model/hicom_arch.py:167-178), SURVEY §8 row f1.

The reference calls ``mm_projector`` once per sample in a Python loop.  ``compress_samples`` takes the same per-sample
inputs, groups the plain video tensors that share a shape / dtype / guide shape, runs each group through ONE
``forward_batched`` call (over a zero-copy batch VIEW when the samples are consecutive ``split`` views of the tower's
output, as in hicom_arch.py:162-164), and sends everything else (any-res image dicts, odd shapes) through the reference-signature
``forward`` — returning the per-sample token tensors in the original order, exactly what the loop produced.
"""
from __future__ import annotations

from collections import defaultdict
from typing import List, Optional, Sequence

import torch


def _key(feat, embed, guide, modal):
    return (modal, tuple(feat.shape), feat.dtype, feat.device, embed is None,
            None if guide is None else tuple(guide.shape))


def batch_view(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    """One (B, *shape) tensor over same-shape tensors WITHOUT copying when they already sit back to back in one
    allocation at a constant pitch — which is what ``tower_output.split(sizes, dim=0)`` hands the reference's loop
    (hicom_arch.py:162-164: the tower runs on all frames of the batch at once, the per-sample tensors are views of its
    output).  Falls back to ``torch.stack`` (one copy) for tensors that are scattered."""
    first = tensors[0]
    if len(tensors) == 1:
        return first.unsqueeze(0)
    same = all(t.shape == first.shape and t.stride() == first.stride() and t.dtype == first.dtype
               and t.device == first.device and t.untyped_storage().data_ptr() == first.untyped_storage().data_ptr()
               for t in tensors)
    if same and first.is_contiguous():
        pitch = tensors[1].storage_offset() - first.storage_offset()
        if pitch >= first.numel() and all(t.storage_offset() == first.storage_offset() + i * pitch
                                          for i, t in enumerate(tensors)):
            view = first.as_strided((len(tensors),) + tuple(first.shape), (pitch,) + tuple(first.stride()),
                                    first.storage_offset())
            return view
    return torch.stack(list(tensors))


def compress_samples(projector, frames_features: Sequence, frames_embeds: Optional[Sequence],
                     guide_embeds: Optional[Sequence], modalities: Sequence[str], image_newline=None,
                     min_group: int = 2) -> List[torch.Tensor]:
    """``frames_features[i]`` is a (T,H,W,d) tensor or an any-res dict, ``frames_embeds[i]`` / ``guide_embeds[i]`` the
    matching embeds (the sequences themselves may be None, hicom_arch.py:169-170).  Returns ``[tokens_i]``.
    Grad mode is the caller's: under ``torch.no_grad()`` the inference kernels run, with grad enabled (the same loop is
    the forward of a training step, train.py) the tokens carry an autograd graph (``hicom_b200/autograd.py``)."""
    n = len(frames_features)
    out: List[Optional[torch.Tensor]] = [None] * n
    groups = defaultdict(list)
    for i in range(n):
        f = frames_features[i]
        e = None if frames_embeds is None else frames_embeds[i]
        g = None if guide_embeds is None else guide_embeds[i]
        if isinstance(f, dict) or f.dim() != 4:
            out[i] = projector(f, e, g, modalities[i], image_newline)
        else:
            groups[_key(f, e, g, modalities[i])].append(i)
    for key, idx in groups.items():
        modal = key[0]
        if len(idx) < min_group:
            for i in idx:
                out[i] = projector(frames_features[i], None if frames_embeds is None else frames_embeds[i],
                                   None if guide_embeds is None else guide_embeds[i], modal, image_newline)
            continue
        # consecutive `split` views of the tower's output are viewed in place (no copy of 27 MB per video)
        X = batch_view([frames_features[i] for i in idx])
        E = None if frames_embeds is None or frames_embeds[idx[0]] is None else batch_view([frames_embeds[i] for i in idx])
        G = None if guide_embeds is None or guide_embeds[idx[0]] is None else batch_view([guide_embeds[i] for i in idx])
        tokens = projector.forward_batched(X, E, G, modal, image_newline)
        for j, i in enumerate(idx):
            out[i] = tokens[j]
    return out
