"""CPU oracle for the producer side of the compressor.

TEST INFRASTRUCTURE ONLY — same rules as ``oracle/hicom_oracle.py``: imported by ``tests/`` (and bench baselines)
as the checker, never by ``hicom_b200``.

Restates what ``SiglipVisionTower.forward`` does after the SigLIP transformer to build ``frames_embed`` and the
``fine`` instruction tokens (``/root/reference/hicom/model/encoder.py:272-286``).  The arithmetic lives in a
third-party dependency that is NOT vendored in the reference: HF transformers (pinned ``transformers==4.46.3`` in the
reference's README.md:20 / requirements.txt), ``models/siglip/modeling_siglip.py``:
  * ``SiglipMLP.forward``:  ``fc2(act(fc1(x)))`` with ``act = ACT2FN[config.hidden_act]``; so400m-patch14-384 uses
    ``hidden_act = "gelu_pytorch_tanh"`` (tanh-form GELU), ``hidden_size = 1152``, ``intermediate_size = 4304``;
  * ``SiglipMultiheadAttentionPoolingHead.layernorm = nn.LayerNorm(hidden_size, eps=config.layer_norm_eps)``, 1e-6.
Pinning: ``tests/test_producer_cpu.py`` runs this restatement against the transformers classes themselves (the
image ships transformers, here and on the GPU box) with the reference's own call sequence — bit-exact in fp32.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-6          # SiglipVisionConfig.layer_norm_eps
HIDDEN, INTERMEDIATE = 1152, 4304   # siglip-so400m-patch14-384


def head_mlp(sd: Dict[str, Tensor], x: Tensor, act: str = "gelu_pytorch_tanh") -> Tensor:
    """SiglipMLP.forward (modeling_siglip.py): fc1 -> activation -> fc2."""
    y = F.linear(x, sd["mlp.fc1.weight"], sd["mlp.fc1.bias"])
    y = F.gelu(y, approximate="tanh") if act == "gelu_pytorch_tanh" else F.gelu(y)
    return F.linear(y, sd["mlp.fc2.weight"], sd["mlp.fc2.bias"])


def image_embeds(sd: Dict[str, Tensor], last_hidden_state: Tensor, side: int = 27,
                 act: str = "gelu_pytorch_tanh") -> Tensor:
    """encoder.py:284-286:
        image_embeds = head.layernorm(last_hidden_state)
        image_embeds = last_hidden_state + head.mlp(image_embeds)
        image_embeds = rearrange(image_embeds, 'b (h w) d -> b h w d', h=side, w=side)
    """
    h = last_hidden_state
    d = h.shape[-1]
    y = F.layer_norm(h, (d,), sd["layernorm.weight"], sd["layernorm.bias"], LN_EPS)
    out = h + head_mlp(sd, y, act)
    return out.reshape(h.shape[0], side, side, d)


def text_embeds_fine(head_w: Tensor, head_b: Tensor, last_hidden_state: Tensor) -> Tensor:
    """encoder.py:279-280: ``guide_encoder.text_model.head(text_forward_out.last_hidden_state)`` (a Linear)."""
    return F.linear(last_hidden_state, head_w, head_b)


def synth_head_state(seed: int = 0, hidden: int = HIDDEN, inter: int = INTERMEDIATE) -> Dict[str, Tensor]:
    """Seeded stand-in for the pooling head's parameters (no checkpoint is reachable offline)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, std: std * torch.randn(*s, generator=g)
    return {
        "layernorm.weight": 1.0 + r(hidden, std=0.1), "layernorm.bias": r(hidden, std=0.1),
        "mlp.fc1.weight": r(inter, hidden, std=0.03), "mlp.fc1.bias": r(inter, std=0.05),
        "mlp.fc2.weight": r(hidden, inter, std=0.02), "mlp.fc2.bias": r(hidden, std=0.05),
    }


def synth_hidden(b: int, tokens: int, seed: int, hidden: int = HIDDEN, std: float = 0.7) -> Tensor:
    g = torch.Generator().manual_seed(1000 + seed)
    return std * torch.randn(b, tokens, hidden, generator=g)
