"""One training step (forward + backward, width 3584, 8 videos x 16 frames, bf16, coarse) between
cudaProfilerStart/Stop, for `ncu --profile-from-start off --metrics gpu__time_duration.sum --csv`.
    ncu ... python tools/train_profile.py [coarse|direct]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import hicom_b200
from hicom_b200 import autograd as ag
from util import Cfg

guide = sys.argv[1] if len(sys.argv) > 1 else "coarse"
ag.enable(True)
torch.manual_seed(0)
m = hicom_b200.build_vision_projector(Cfg(use_guide=guide, hidden_size=3584, max_num_frames=16))
with torch.no_grad():
    m.global_compressor.query.normal_(std=0.02)
m = m.to(torch.bfloat16).cuda().train()
B = 8
X = (0.5 * torch.randn(B, 16, 27, 27, 1152, device="cuda")).bfloat16()
E = (0.5 * torch.randn(B, 16, 27, 27, 1152, device="cuda")).bfloat16()
G = (0.5 * torch.randn(B, 1152, device="cuda")).bfloat16()


def step():
    m.zero_grad(set_to_none=True)
    out = m.forward_batched(X, E, G, "video")
    out.float().square().mean().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
