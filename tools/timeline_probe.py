"""Timeline of one CUDA-graph replay of the forward (torch.profiler / CUPTI): per kernel start, duration and stream, so
that overlap and idle gaps of the two chains are visible.  python tools/timeline_probe.py [B T] > timeline.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hicom_b200.graph import GraphedCompressor
from __graft_entry__ import build
build()
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 16)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
proj = bench.build_projector(3584, dev)
X, E, G = bench.synth_batch(B, T, dev, 1)
with torch.no_grad():
    g = GraphedCompressor(proj, X, E, G, "video")
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
n = len(ev) // 3
last = ev[2 * n:]
t0 = last[0].time_range.start
rows = [{"t_us": round(e.time_range.start - t0, 1), "dur_us": round(e.time_range.end - e.time_range.start, 1),
         "name": e.name[:70]} for e in last]
print(json.dumps({"B": B, "T": T, "kernels": len(last), "span_us": round(last[-1].time_range.end - t0, 1), "rows": rows}, indent=0))
