"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU paths (sharding + the one exchange)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _merge(m, l, o):
    M = m.max(1, keepdim=True).values
    w = (m - M).exp()
    L = (l * w).sum(1)
    return (o * w[..., None]).sum(1) / L[..., None]


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hicom_b200 import dist as hd
    B, S, J, d, N = 2, 2, 6, 8, 40
    g = torch.Generator().manual_seed(0)
    scores = torch.randn(B, N, J, generator=g) * 3
    vals = torch.randn(B, N, d, generator=g)
    t0, t1 = hd.frame_shard(N, world, rank, temporal_kernel=4)
    # this rank's partials over its token block, cut again in S local splits
    ms, ls, os_ = [], [], []
    for blk in torch.arange(t0, t1).chunk(S):
        s = scores[:, blk]
        m = s.max(1).values
        p = (s - m[:, None]).exp()
        ms.append(m); ls.append(p.sum(1)); os_.append(torch.einsum("bnj,bnd->bjd", p, vals[:, blk]))
    m, l, o = torch.stack(ms, 1), torch.stack(ls, 1), torch.stack(os_, 1)
    gm, gl, go = hd.gather_partials(m, l, o)
    assert gm.shape == (B, world * S, J) and go.shape == (B, world * S, J, d)
    merged = _merge(gm, gl, go)
    want = torch.einsum("bnj,bnd->bjd", torch.softmax(scores, 1), vals)
    err = float((merged - want).abs().max())
    # every rank holds the same merged result
    ref = merged.clone()
    dist.broadcast(ref, 0)
    same = float((merged - ref).abs().max())
    # the 16-bit exchange: every rank normalises on its own, applies the per-head value projection and sends
    # [rows | lse]; softmax weights of the lse combine them (what ops.shard_combine does on the GPU)
    from hicom_b200 import ops
    heads, Q = 2, 3                                      # J = heads * Q = 6 columns
    hdim = d // heads
    Wv = torch.randn(d, d, generator=g) * 0.3
    blk = torch.arange(t0, t1)
    s_loc = scores[:, blk]
    lse = torch.logsumexp(s_loc, 1)                      # (B, J)
    pooled = torch.einsum("bnj,bnd->bjd", torch.softmax(s_loc, 1), vals[:, blk])   # (B, J, d), normalised on this rank
    rows = torch.stack([torch.cat([pooled[:, h * Q + i] @ Wv[h * hdim:(h + 1) * hdim].T for h in range(heads)], -1)
                        for i in range(Q)], 1)           # (B, Q, d): head h -> channels h*hd..
    nbytes, lse_off = ops.shard_message_layout(Q, d, heads, torch.float32)
    msg = torch.zeros(B, nbytes, dtype=torch.uint8)
    msg[:, :Q * d * 4] = rows.contiguous().view(B, -1).view(torch.uint8)
    msg[:, lse_off:lse_off + J * 4] = lse.contiguous().view(torch.uint8)
    allm = hd.gather_messages(msg)
    assert allm.shape == (world, B, nbytes)
    r_rows = allm[:, :, :Q * d * 4].contiguous().view(torch.float32).view(world, B, Q, d)
    r_lse = allm[:, :, lse_off:lse_off + J * 4].contiguous().view(torch.float32).view(world, B, heads, Q)
    wgt = torch.softmax(r_lse, 0).permute(0, 1, 3, 2)    # (world, B, Q, heads)
    comb = (r_rows.view(world, B, Q, heads, hdim) * wgt[..., None]).sum(0).reshape(B, Q, d)
    want_rows = torch.stack([torch.cat([want[:, h * Q + i] @ Wv[h * hdim:(h + 1) * hdim].T for h in range(heads)], -1)
                             for i in range(Q)], 1)
    err2 = float((comb - want_rows).abs().max())
    if rank == 0:
        results.put((max(err, err2), same))
    dist.destroy_process_group()


def test_frame_sharded_merge_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    err, same = q.get()
    assert err < 1e-5 and same < 1e-6


def test_shard_arithmetic():
    from hicom_b200 import dist as hd
    for n, w in [(32, 8), (33, 8), (5, 8), (64, 3)]:
        blocks = [hd.video_shard(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [e - b for b, e in blocks]
        assert max(sizes) - min(sizes) <= 1
    assert [hd.frame_shard(512, 8, r) for r in (0, 7)] == [(0, 64), (448, 512)]
    assert hd.frame_shard(8, 4, 3) == (4, 8) or hd.frame_shard(8, 4, 3) == (8, 8)
    spans = [hd.frame_shard(24, 4, r) for r in range(4)]
    assert all((e - b) % 4 == 0 for b, e in spans) and spans[-1][1] == 24
    with pytest.raises(ValueError):
        hd.frame_shard(30, 4, 0)
    m, l, o = torch.randn(2, 3, 4), torch.rand(2, 3, 4), torch.randn(2, 3, 4, 5)
    m2, l2, o2 = hd.unpack_partials(hd.pack_partials(m, l, o))
    assert torch.equal(m, m2) and torch.equal(l, l2) and torch.equal(o, o2)


# ---- data-parallel training through the drop-in module (train.py launches one process per GPU with torchrun) ----------
def _ddp_worker(rank, world, port, results):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_ops
    import hicom_b200
    from hicom_b200 import autograd as ag, ops
    from oracle.cases import CASES_BY_NAME, materialise
    from util import cfg_for
    for name in cpu_ops.ALL:  # torch-CPU stand-ins for the kernels (tests/cpu_ops.py): this test is about the plumbing
        setattr(ops, name, getattr(cpu_ops, name))
    ag.enable(True)
    case = CASES_BY_NAME["coarse_T4"]
    sd, X, E, g, _ = materialise(case)
    m = hicom_b200.build_vision_projector(cfg_for(case))
    m.load_state_dict(sd, strict=True)
    # find_unused_parameters stays False: every trainable parameter must receive a gradient (k_proj.bias gets zeros)
    ddp = torch.nn.parallel.DistributedDataParallel(m.train())
    Xr, Er, gr = (X, E, g) if rank == 0 else (X.flip(0) * 0.9, E.flip(0) * 0.9, -g)
    out = ddp(Xr, Er, gr, "video")
    out.square().mean().backward()
    grads = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    ref = grads.clone()
    dist.broadcast(ref, 0)
    same = float((grads - ref).abs().max())
    missing = [k for k, p in m.named_parameters() if p.grad is None]
    if rank == 0:
        results.put((same, missing, float(grads.abs().sum())))
    dist.destroy_process_group()


def test_ddp_training_world2():
    """DistributedDataParallel over the training path, two ranks with different videos: the reducer sees a gradient for
    every parameter (no find_unused_parameters) and both ranks end with the same averaged gradients."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    same, missing, total = q.get()
    assert same == 0.0 and missing == [] and total > 0.0
