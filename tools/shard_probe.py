"""One rank's share of the frame-sharded long video (BASELINE config 4 at 8 GPUs: 64 frames, batch 1) on ONE GPU:
CUDA-graph replay time against the sum of the kernels' own durations (per-op CUDA events on one stream), i.e. how much
of the rank's time is the dependent chain of small kernels.  python tools/shard_probe.py [frames]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hicom_b200 import ops, dist as hdist
from hicom_b200.graph import GraphedCompressor
import hicom_b200.projector as P

T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
from __graft_entry__ import build
build()
proj = bench.build_projector(3584, dev)
X, E, G = bench.synth_batch(1, T, dev, 1)


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


res = {"frames": T}
with torch.no_grad():
    g = GraphedCompressor(proj, X, E, G, "video", frame_shard_t0=0)
    res["graph_replay_ms"] = timed(g.replay)
    res["kernels_per_replay"] = g.kernels_per_replay
    res["eager_ms"] = timed(lambda: hdist.forward_frame_sharded(proj, X, E, G, t0=0), 20)
    P.OVERLAP_STREAMS, hdist.SHARD_OVERLAP = False, False
    hdist.forward_frame_sharded(proj, X, E, G, t0=0)
    with ops.OpTimer() as t:
        for _ in range(10):
            hdist.forward_frame_sharded(proj, X, E, G, t0=0)
    summ = t.summary()
    res["sum_of_op_ms"] = sum(ms for _, ms in summ.values()) / 10
    res["ops"] = {k: round(ms / 10 * 1e3, 1) for k, (c, ms) in sorted(summ.items(), key=lambda kv: -kv[1][1])}
    with ops.KernelTimer() as kt:
        for _ in range(10):
            hdist.forward_frame_sharded(proj, X, E, G, t0=0)
    res["kernels_us"] = {k: round(ms / 10 * 1e3, 1) for k, (c, ms) in sorted(kt.summary().items(), key=lambda kv: -kv[1][1])}
print(json.dumps(res, indent=1))
