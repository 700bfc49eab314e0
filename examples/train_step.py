"""One optimiser step through the drop-in projector (needs a B200) — what stages 1-3 of the reference do to
mm_projector (train.py:704-738); gradients come from hicom_b200/autograd.py.

    python examples/train_step.py
"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hicom_b200  # noqa: E402
from hicom_b200.producer import SiglipHeadEmbed  # noqa: E402

config = types.SimpleNamespace(mm_projector_type="local43_adaptkv_global32", mm_vision_tower="google/siglip-so400m-patch14-384",
                               mm_hidden_size=1152, hidden_size=3584, use_guide="direct", max_num_frames=64)
projector = hicom_b200.build_vision_projector(config)
with torch.no_grad():                                   # adapters start at alpha = 0 in the reference; move them a little
    projector.local_compressor.k_alpha.fill_(0.1), projector.local_compressor.v_alpha.fill_(0.1)
projector = projector.to(torch.bfloat16).cuda().train()
head = SiglipHeadEmbed().to(torch.bfloat16).cuda().train()   # vision_model_head, tuned in stage 3 (train.py:717-721)
opt = torch.optim.AdamW(list(projector.parameters()) + list(head.parameters()), lr=1e-4)

B, T = 4, 16
frames_feature = (0.5 * torch.randn(B, T, 27, 27, 1152, device="cuda")).bfloat16()        # frozen SigLIP body
last_hidden = (0.5 * torch.randn(B * T, 729, 1152, device="cuda")).bfloat16()
guide = (0.5 * torch.randn(B, 1152, device="cuda")).bfloat16().requires_grad_(True)       # guide_encoder output

frames_embed = head(last_hidden).view(B, T, 27, 27, 1152)                                  # carries a graph to the head
tokens = projector.forward_batched(frames_feature, frames_embed, guide, "video")           # (B, 356, 3584)
loss = tokens.float().square().mean()
loss.backward()
opt.step()
missing = [k for k, p in projector.named_parameters() if p.grad is None]
print("loss", float(loss), "| parameters without a gradient:", missing, "| d(guide)", tuple(guide.grad.shape))

# Fixed shapes: forward AND backward replayed from CUDA graphs (about 1.8x faster than the eager step, which is bound by
# ~150 launches).  Build it on a module that has not run an eager backward yet.
from hicom_b200.graph import graphed_training_forward  # noqa: E402
proj2 = hicom_b200.build_vision_projector(config).to(torch.bfloat16).cuda().train()
E2 = frames_embed.detach()
step = graphed_training_forward(proj2, frames_feature, E2, guide.detach(), "video")
out = step(frames_feature, E2, guide.detach())
out.float().square().mean().backward()
print("graphed step:", tuple(out.shape), "| parameters without a gradient:",
      [k for k, p in proj2.named_parameters() if p.grad is None])
