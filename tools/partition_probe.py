"""SM-partition probe: the local window attention (HBM-bound) and the global attention partials (tensor-bound) alone at
several SM limits, then side by side on two streams.  python tools/partition_probe.py [B T]"""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hicom_b200 import ops
from hicom_b200.projector import default_splits

B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 16)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
from __graft_entry__ import build
build()
proj = bench.build_projector(3584, dev)
X, E, G = bench.synth_batch(B, T, dev, 1)
lc, gc = proj.local_compressor, proj.global_compressor
side = torch.cuda.Stream(dev)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


res = {"B": B, "T": T}
with torch.no_grad():
    Qg = gc.injected_query(G, B, X.dtype)
    qf = gc.fold(Qg)

    def local(limit):
        with ops.sm_limit(limit):
            return lc.attend(X, E, G, "video")

    def glob(limit):
        with ops.sm_limit(limit):
            return gc.partials(X, qf)

    for L in (0, 32, 40, 48, 56, 64, 74, 100):
        res[f"local@{L}"] = timed(lambda: local(L))
    for L in (0, 120, 108, 100, 92, 84, 74):
        res[f"global@{L}"] = timed(lambda: glob(L))

    def both(L):
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            local(L)
        glob(148 - L if L else 0)
        main.wait_stream(side)

    for L in (0, 32, 40, 48, 56, 64):
        res[f"both@{L}"] = timed(lambda: both(L))
print(json.dumps(res, indent=1))
