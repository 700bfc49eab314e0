"""fp16 probe: which tcgen05 launch shapes / operand-format combinations run (each case in its own process)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASES = ["linear_small", "linear_big", "fold", "partial_c2", "forward"]
if len(sys.argv) < 2:
    for c in CASES:
        r = subprocess.run([sys.executable, __file__, c], capture_output=True, text=True,
                           env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"))
        print(c, "rc", r.returncode, (r.stdout.strip().splitlines() or [""])[-1], (r.stderr.strip().splitlines() or [""])[-1][:300])
    sys.exit(0)
import torch
from hicom_b200 import ops
import hicom_b200.projector as P
case = sys.argv[1]
dev = "cuda"
torch.manual_seed(0)
h = torch.float16
if case.startswith("linear"):
    M = 64 if case == "linear_small" else 10368
    A = (0.5 * torch.randn(M, 1152, device=dev)).to(h); W = (0.05 * torch.randn(3584, 1152, device=dev)).to(h)
    b = (0.1 * torch.randn(3584, device=dev)).to(h)
    out = ops.linear(A, W, b, None, ops.ACT_GELU, False, ops.IMPL_AUTO)
    torch.cuda.synchronize()
    ref = torch.nn.functional.gelu(A.float() @ W.float().T + b.float())
    print("rel", float((out.float() - ref).abs().max() / ref.abs().max()))
elif case == "fold":
    q = (0.5 * torch.randn(2, 32, 1152, device=dev)).to(h); Wk = (0.05 * torch.randn(1152, 1152, device=dev)).to(h)
    out = ops.global_fold_query(q, Wk, 9, 0.088)
    torch.cuda.synchronize()
    print("ok", tuple(out.shape), out.dtype)
elif case == "partial_c2":
    X = (0.5 * torch.randn(2, 16, 27, 27, 1152, device=dev)).to(h)
    qf = (0.05 * torch.randn(2, 288, 1152, device=dev)).to(h)
    pt, ph, pw = (torch.randn(n, 1152, device=dev) for n in (16, 27, 27))
    m, l, o = ops.global_attend_partial(X, pt, ph, pw, qf, 1, ops.IMPL_AUTO)
    torch.cuda.synchronize()
    pooled = ops.softmax_merge(m, l, o, 0)
    Xp = X.float() + pt[:, None, None] + ph[None, :, None] + pw[None, None]
    S = torch.einsum("bnd,bjd->bnj", Xp.reshape(2, -1, 1152), qf.float())
    ref = torch.einsum("bnj,bnd->bjd", S.softmax(1), Xp.reshape(2, -1, 1152))
    print("rel", float((pooled - ref).abs().max() / ref.abs().max()))
else:
    import bench
    bench.DTYPE = h
    proj = bench.build_projector(3584, torch.device(dev))
    X, E, G = bench.synth_batch(2, 16, torch.device(dev), 1)
    with torch.no_grad():
        out = proj.forward_batched(X, E, G, "video")
    torch.cuda.synchronize()
    print("ok", tuple(out.shape), out.dtype, bool(torch.isfinite(out.float()).all()))
