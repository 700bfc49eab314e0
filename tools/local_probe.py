"""Window attention alone at several SM limits (c2 shapes), ms per call.  HICOM_LOCAL_TWO_PHASE=0|1 python tools/local_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hicom_b200 import ops
from __graft_entry__ import build
build()
dev = torch.device("cuda", 0)
proj = bench.build_projector(3584, dev)
X, E, G = bench.synth_batch(32, 16, dev, 1)
lc = proj.local_compressor


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


res = {"two_phase": os.environ.get("HICOM_LOCAL_TWO_PHASE", "1")}
with torch.no_grad():
    for L in (0, 64, 48, 40, 32):
        def f():
            with ops.sm_limit(L):
                return lc.attend(X, E, G, "video")
        res[f"local@{L}"] = round(timed(f), 4)
    res["checksum"] = float(lc.attend(X, E, G, "video").float().abs().mean())
print(json.dumps(res))
