"""Drop-in use for inference (needs a B200): the reference's call at hicom_arch.py:212 with hicom_b200 behind it.

    python examples/inference.py
"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hicom_b200  # noqa: E402  (hicom/model/hicom_arch.py:26 would import build_vision_projector from here)

config = types.SimpleNamespace(mm_projector_type="local43_global32", mm_vision_tower="google/siglip-so400m-patch14-384",
                               mm_hidden_size=1152, hidden_size=3584, use_guide="direct", max_num_frames=64)
projector = hicom_b200.build_vision_projector(config).to(torch.bfloat16).cuda().eval()
# projector.load_state_dict(mm_projector_weights, strict=False)   # mm_projector.bin of the reference loads unchanged

T = 16
frames_feature = (0.5 * torch.randn(T, 27, 27, 1152, device="cuda")).bfloat16()   # SigLIP hidden_states[-2]
frames_embed = (0.5 * torch.randn(T, 27, 27, 1152, device="cuda")).bfloat16()     # encoder.py:284-286 (hicom_b200.producer)
guide_embed = (0.5 * torch.randn(1152, device="cuda")).bfloat16()                 # pooled SigLIP text embedding

with torch.inference_mode():
    tokens = projector(frames_feature, frames_embed, guide_embed, "video")        # (T/4*81 + 32, 3584)
    batch = projector.forward_batched(frames_feature[None].expand(4, -1, -1, -1, -1).contiguous(),
                                      frames_embed[None].expand(4, -1, -1, -1, -1).contiguous(),
                                      guide_embed[None].expand(4, -1).contiguous(), "video")
print(tuple(tokens.shape), tuple(batch.shape), float((batch[0].float() - tokens.float()).abs().max()))

# fp16 is the reference's own inference dtype (hicom/model/__init__.py:44): same call, native tcgen05 path
p16 = projector.half()
with torch.inference_mode():
    t16 = p16(frames_feature.half(), frames_embed.half(), guide_embed.half(), "video")
print("fp16", t16.dtype, float((t16.float() - tokens.float()).abs().max() / tokens.float().abs().max()))
