"""Per-step timing of hicom_b200.pipeline.compress_from_host on the c2 workload (diagnostic tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hicom_b200.pipeline import compress_from_host, host_affinity

dev = torch.device("cuda", 0)
hidden, T, B, _ = bench.WORKLOADS["c2"]
proj = bench.build_projector(hidden, dev)
X, E, G = bench.synth_batch(B, T, dev, 1234)
with torch.no_grad():
    out = proj.forward_batched(X, E, G, "video")
with host_affinity(dev):
    Xh, Eh, Gh = [t.cpu().pin_memory() for t in (X, E, G)]
    out_h = torch.empty(tuple(out.shape), dtype=out.dtype).pin_memory()
for chunk in (4, 8, 2):
    ts = []
    for i in range(12):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        compress_from_host(proj, Xh, Eh, Gh, "video", out=out_h, chunk=chunk, device=dev)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print("chunk", chunk, " ".join(f"{t:.1f}" for t in ts), flush=True)
