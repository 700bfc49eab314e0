"""Worker of tests/test_gpu_dist.py: one process per GPU (torchrun), NCCL.  Each rank holds the same synthetic long video,
takes its block of frames, runs ``dist.forward_frame_sharded`` (eagerly and as a captured CUDA graph) and compares its
tokens with the unsharded ``forward_batched`` of the whole video on its own GPU.  Prints ``DIST_OK`` per rank."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import hicom_oracle as O  # noqa: E402  (test infrastructure: weights / inputs / error metrics)
from oracle.cases import Case  # noqa: E402
from util import cuda_module_for  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    from hicom_b200 import dist as hdist
    from hicom_b200.graph import GraphedCompressor
    worst = 0.0
    for mode, hidden, T, dname in (("coarse", 3584, 16 * world, "bfloat16"), ("direct", 896, 8 * world, "bfloat16"),
                                   (None, 896, 8 * world, "bfloat16"), ("coarse", 896, 8 * world, "float16")):
        dt = getattr(torch, dname)
        case = Case(f"dist_{mode}", "local43_global32", mode, T, 27, 27, hidden, dname, wseed=3)
        sd = O.synth_state_dict(case.ptype, mode, hidden, seed=3, dtype=dt)
        m = cuda_module_for(case, sd, dev)
        X, E, g = O.synth_inputs(T, 27, 27, O.guide_kind_for(mode), seed=77, dtype=dt)  # same on every rank
        dv = lambda t: None if t is None else t.unsqueeze(0).to(dev)
        X, E, g = dv(X), dv(E), dv(g)
        t0, t1 = hdist.frame_shard(T, world, rank)
        Xs = X[:, t0:t1].contiguous()
        Es = None if E is None else E[:, t0:t1].contiguous()
        with torch.no_grad():
            whole = m.forward_batched(X, E, g, "video").float()
            loc, glob = hdist.forward_frame_sharded(m, Xs, Es, g, t0=t0)
            graphed = GraphedCompressor(m, Xs, Es, g, "video", frame_shard_t0=t0)
            gloc, gglob = graphed.replay()
            torch.cuda.synchronize()
        nw = loc.shape[1]
        scale = float(whole.abs().max())
        for name, got, want in (("local", loc, whole[:, rank * nw:(rank + 1) * nw]), ("global", glob, whole[:, -glob.shape[1]:]),
                                ("graph local", gloc, whole[:, rank * nw:(rank + 1) * nw]),
                                ("graph global", gglob, whole[:, -glob.shape[1]:])):
            err = float((got.float() - want).abs().max()) / scale
            worst = max(worst, err)
            assert err <= 8e-3, (mode, name, rank, err)
        # every rank finished the same global tokens
        ref = glob.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, glob), "ranks disagree on the global tokens"
        del graphed
    print(f"DIST_OK rank {rank}/{world} worst rel err {worst:.2e}", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)  # no destructor-time NCCL teardown (a captured graph holding the communicator can hang there)


if __name__ == "__main__":
    main()
