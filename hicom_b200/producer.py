"""Producer side of the compressor: what ``SiglipVisionTower.forward`` computes AFTER the SigLIP
transformer to make the compressor's key features and instruction tokens (hicom/model/encoder.py:272-286).

    image_embeds = h + head.mlp(head.layernorm(h))          encoder.py:284-285   (h = last_hidden_state)
    image_embeds -> (b, 27, 27, 1152)                        encoder.py:286       (a view, no copy)
    text_embeds  = text_model.head(last_hidden_state)        encoder.py:279-280   (``fine`` guide mode)

``head`` is SigLIP's attention-pooling head (``SiglipMultiheadAttentionPoolingHead`` of HF transformers, pinned at
4.46.3 by the reference, README.md:20): only its ``layernorm`` (eps = ``layer_norm_eps`` = 1e-6) and its ``mlp``
(``fc1`` 1152 -> 4304, tanh-form GELU ``gelu_pytorch_tanh``, ``fc2`` 4304 -> 1152) are used here, on ALL 729 tokens of
every frame, i.e. 2 * 2 * 729 * 1152 * 4304 = 14.5 GFLOP per frame — about three times the reference compressor and
nine times what the reassociated compressor executes.  The SigLIP tower itself stays reference PyTorch (out of scope).

Kernels: ``ops.layernorm`` (one warp per row), then the persistent tcgen05 linear twice — ``fc1`` with the tanh GELU
fused in its epilogue (16 epilogue warps, two MUFU per element), ``fc2`` with bias and the residual ``h`` fused.  No
PyTorch/CPU fallback.  Under autograd (stage 3 tunes ``vision_model_head``, train.py:717-721) the same kernels run
inside the Functions of ``hicom_b200/autograd.py`` (LayerNormFn, LinearFn), so the head's layernorm / MLP parameters
receive gradients; ``HICOM_AUTOGRAD=0`` restores the forward-only failure.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .projector import _IMPL, _grad_needed, _require_no_grad

__all__ = ["SiglipHeadEmbed", "text_head_embed"]


class _HeadMLP(nn.Module):
    """Parameter container with SiglipMLP's names (``fc1``, ``fc2``)."""

    def __init__(self, hidden_size: int, intermediate_size: int):
        super().__init__()
        self.fc1 = nn.Linear(hidden_size, intermediate_size)
        self.fc2 = nn.Linear(intermediate_size, hidden_size)


class SiglipHeadEmbed(nn.Module):
    """``frames_embed`` producer.  Parameter names are those of ``vision_model.head`` (``layernorm.{weight,bias}``,
    ``mlp.fc1.{weight,bias}``, ``mlp.fc2.{weight,bias}``), so ``load_state_dict(head.state_dict(), strict=False)``
    picks them up; ``from_head`` shares the tensors instead of copying them."""

    def __init__(self, hidden_size: int = 1152, intermediate_size: int = 4304, layer_norm_eps: float = 1e-6,
                 hidden_act: str = "gelu_pytorch_tanh"):
        super().__init__()
        if abs(layer_norm_eps - 1e-6) > 1e-12:
            raise NotImplementedError(f"layer_norm_eps={layer_norm_eps}: the row kernels are built for 1e-6 (SigLIP)")
        acts = {"gelu_pytorch_tanh": ops.ACT_GELU_TANH, "gelu": ops.ACT_GELU}
        if hidden_act not in acts:
            raise NotImplementedError(f"hidden_act={hidden_act!r}: expected one of {sorted(acts)}")
        self.act = acts[hidden_act]
        self.layernorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        self.mlp = _HeadMLP(hidden_size, intermediate_size)

    @classmethod
    def from_head(cls, head: nn.Module, hidden_act: str = "gelu_pytorch_tanh") -> "SiglipHeadEmbed":
        """Wrap ``vision_tower.vision_model.head`` (shares its parameters)."""
        fc1, fc2 = head.mlp.fc1, head.mlp.fc2
        act = getattr(getattr(head.mlp, "config", None), "hidden_act", hidden_act)
        self = cls(fc1.in_features, fc1.out_features, head.layernorm.eps, act)
        self.layernorm.weight, self.layernorm.bias = head.layernorm.weight, head.layernorm.bias
        self.mlp.fc1.weight, self.mlp.fc1.bias = fc1.weight, fc1.bias
        self.mlp.fc2.weight, self.mlp.fc2.bias = fc2.weight, fc2.bias
        return self

    def forward(self, last_hidden_state: torch.Tensor, num_patches_per_side: int | None = None) -> torch.Tensor:
        """(b, h*w, d) -> (b, h, w, d) = ``h + mlp(layernorm(h))`` (encoder.py:284-286)."""
        h = last_hidden_state
        if h.dim() != 3:
            raise ValueError(f"expected last_hidden_state (b, tokens, d), got {tuple(h.shape)}")
        if h.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            raise TypeError(f"hicom_b200 supports float32, bfloat16 and float16 tensors, got {h.dtype}")
        b, n, d = h.shape
        side = num_patches_per_side if num_patches_per_side is not None else int(round(n ** 0.5))
        if side * side != n:
            raise ValueError(f"{n} tokens are not a {side} x {side} grid")
        h = h.contiguous()
        if _grad_needed(self, h):
            from . import autograd as ag
            if not ag.ENABLED:
                _require_no_grad(self, h)
            ops._need_cuda(h)
            y = ag.LayerNormFn.apply(h, ag._as(self.layernorm.weight, h), ag._as(self.layernorm.bias, h))
            y = ag.linear(y, self.mlp.fc1.weight, self.mlp.fc1.bias, None, self.act)
            y = ag.linear(y, self.mlp.fc2.weight, self.mlp.fc2.bias, h)
            return y.view(b, side, side, d)
        y = ops.layernorm(h, self.layernorm.weight, self.layernorm.bias)
        y = ops.linear(y, self.mlp.fc1.weight, self.mlp.fc1.bias, None, self.act, False, _IMPL)
        y = ops.linear(y, self.mlp.fc2.weight, self.mlp.fc2.bias, h, ops.ACT_NONE, False, _IMPL)
        return y.view(b, side, side, d)


def text_head_embed(last_hidden_state: torch.Tensor, head: nn.Linear) -> torch.Tensor:
    """``fine`` guide tokens: ``guide_encoder.text_model.head(last_hidden_state)`` (encoder.py:279-280), (b, L, d)."""
    if _grad_needed(head, last_hidden_state):
        from . import autograd as ag
        if not ag.ENABLED:
            _require_no_grad(head, last_hidden_state)
        ops._need_cuda(last_hidden_state)
        return ag.linear(last_hidden_state, head.weight, head.bias)
    return ops.linear(last_hidden_state, head.weight, head.bias, None, ops.ACT_NONE, False, _IMPL)
