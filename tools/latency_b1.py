"""Single-video latency of the compressor (eager and CUDA-graph replay) — the reference's eval usage is batch 1."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_projector, synth_batch
from hicom_b200.graph import GraphedCompressor

dev = torch.device("cuda", 0)
for hidden, T in ((3584, 16), (3584, 32), (896, 16)):
    proj = build_projector(hidden, dev)
    X, E, G = synth_batch(1, T, dev, 7)
    with torch.no_grad():
        for _ in range(5):
            proj(X[0], E[0], G[0], "video")
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(50):
            proj(X[0], E[0], G[0], "video")
        torch.cuda.synchronize(); eager = (time.perf_counter() - t0) / 50 * 1e3
        g = GraphedCompressor(proj, X, E, G)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(200):
            g.replay()
        torch.cuda.synchronize(); graph = (time.perf_counter() - t0) / 200 * 1e3
    print(f"B=1 hidden={hidden} T={T}: eager forward {eager:.3f} ms, graph replay {graph:.3f} ms ({g.kernels_per_replay} kernels)")
