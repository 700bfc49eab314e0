"""L2-residency probe (one B200): (1) effective L2 capacity for a buffer that is re-read by all SMs, (2) the big three
kernels of a c2 step (score pass, pooling pass, window attention) walked in chunks of a few videos so that the second
and third read of a chunk's X can hit L2.  python tools/chunk_probe.py [B T]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hicom_b200 import ops

B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 16)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
from __graft_entry__ import build
build()


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


res = {"B": B, "T": T, "l2_reread_gbs": {}}
big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
for mb in (16, 32, 48, 64, 80, 96, 112, 128, 160, 256, 1024):
    v = big[: mb << 20].view(torch.float32)
    ms = timed(lambda: v.sum(), iters=20)
    res["l2_reread_gbs"][mb] = round((mb << 20) / ms / 1e6, 0)
del big

proj = bench.build_projector(3584, dev)
X, E, G = bench.synth_batch(B, T, dev, 1)
lc, gc = proj.local_compressor, proj.global_compressor
N = T * 729
with torch.no_grad():
    Qg = gc.injected_query(G, B, X.dtype)
    qf = gc.fold(Qg)

    def walk(c, do_global=True, do_local=True, splits=None):
        for b0 in range(0, B, c):
            xs, es, gs = X[b0:b0 + c], E[b0:b0 + c], G[b0:b0 + c]
            if do_global:
                gc.partials(xs, qf[b0:b0 + c], splits=splits)
            if do_local:
                lc.attend(xs, es, gs, "video")

    for c in (32, 8, 4, 3, 2, 1):
        s = max(1, min(16, 148 // (9 * c)))
        res[f"chunk{c}"] = {
            "splits": s,
            "global_ms": timed(lambda: walk(c, True, False, s)),
            "local_ms": timed(lambda: walk(c, False, True, s)),
            "both_ms": timed(lambda: walk(c, True, True, s)),
        }
    res["chunk32_default_splits"] = {"global_ms": timed(lambda: walk(32, True, False)),
                                     "both_ms": timed(lambda: walk(32, True, True))}
    for c in (32, 2):
        s = max(1, min(16, 148 // (9 * c)))
        walk(c, True, True, s)
        with ops.KernelTimer() as kt:
            for _ in range(5):
                walk(c, True, True, s)
        res[f"kernels_us_chunk{c}"] = {k: [cnt // 5, round(ms / 5 * 1e3, 1)] for k, (cnt, ms) in
                                       sorted(kt.summary().items(), key=lambda kv: -kv[1][1])[:10]}
print(json.dumps(res, indent=1))
