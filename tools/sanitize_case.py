"""Small end-to-end forwards (fp32 + bf16 + fp16, coarse and fine), one training step (incl. gradients into
frames_feature and a trainable clip scale) and the frame-shard message kernels, for compute-sanitizer runs."""
import dataclasses, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle.cases import CASES_BY_NAME, materialise
from util import cuda_module_for

for name, dtype, T in (("coarse_27x27_T4", "bfloat16", 8), ("fine_T8", "bfloat16", 8), ("coarse_T7", "float32", 7),
                       ("coarse_27x27_T4", "float16", 8)):
    case = dataclasses.replace(CASES_BY_NAME[name], dtype=dtype, T=T)
    sd, X, E, g, nl = materialise(case)
    m = cuda_module_for(case, sd)
    with torch.no_grad():
        out = m(X.cuda(), E.cuda(), g.cuda(), "video")
        outb = m.forward_batched(torch.stack([X, X]).cuda(), torch.stack([E, E]).cuda(), torch.stack([g, g]).cuda(), "video")
    torch.cuda.synchronize()
    print(name, dtype, tuple(out.shape), tuple(outb.shape), bool(torch.isfinite(out.float()).all()))

# one training step: gradients into frames_feature, trainable clip scale (skinny NN, grid-pool / l2norm backward kernels)
from hicom_b200 import autograd as ag, ops
ag.enable(True)
case = dataclasses.replace(CASES_BY_NAME["coarse_27x27_T4"], dtype="bfloat16", T=8)
sd, X, E, g, nl = materialise(case)
m = cuda_module_for(case, sd).train()
m.local_logit_scale = torch.nn.Parameter(torch.tensor(1.3, device="cuda"))
m.local_logit_bias = torch.nn.Parameter(torch.tensor(-0.7, device="cuda"))
m.global_logit_scale = torch.nn.Parameter(torch.tensor(1.1, device="cuda"))
m.global_logit_bias = torch.nn.Parameter(torch.tensor(0.2, device="cuda"))
X1 = X.cuda().requires_grad_(True)
out = m(X1, E.cuda(), g.cuda(), "video")
out.float().square().mean().backward()
torch.cuda.synchronize()
print("train step", tuple(out.shape), bool(torch.isfinite(X1.grad.float()).all()), float(m.local_logit_scale.grad))
# frame-shard message + combine
gc = m.global_compressor
with torch.no_grad():
    Xb, gb = X.cuda().unsqueeze(0), g.cuda().unsqueeze(0)
    Qg = gc.injected_query(gb, 1, Xb.dtype)
    qf = gc.fold(Qg)
    msgs = torch.stack([gc.shard_message(Qg, *gc.partials(Xb[:, 4 * r:4 * r + 4].contiguous(), qf, t0=4 * r)) for r in range(2)])
    a = ops.shard_combine(msgs, Qg.shape[1], Qg.shape[2], gc.attn_layer.num_heads, Xb.dtype)
torch.cuda.synchronize()
print("shard combine", tuple(a.shape), bool(torch.isfinite(a.float()).all()))
