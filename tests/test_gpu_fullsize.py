"""GPU: the BENCHMARKED configurations (BASELINE.json configs 2-5: bf16, 27x27 patches, tcgen05 path) against the fp32
oracle — the reference algorithm of projector.py:676-708 on the same bf16-rounded weights and inputs.

Gate (north_star / SURVEY §8c): cosine >= 0.999 and max|a-b| / max|b| <= 1e-2, token order and layout exact (a permuted
row fails the tolerance).  Batches are made of DISTINCT videos; the oracle is evaluated for a few of them (it costs
0.2 s per 16-frame video and ~10 s for the 512-frame one on the host cores).
"""
import pytest
import torch

from oracle import hicom_oracle as O
from oracle.cases import Case

from util import cuda_module_for, oracle_for

pytestmark = pytest.mark.gpu

BF16_REL, BF16_COS = 1e-2, 0.999
PTYPE = "local43_global32"


def _setup(hidden, T, mode, B, wseed=11, xseed=500):
    case = Case(f"full_{hidden}_T{T}_{mode}", PTYPE, mode, T, 27, 27, hidden, "bfloat16", wseed=wseed)
    sd = O.synth_state_dict(PTYPE, mode, hidden, seed=wseed, dtype=torch.bfloat16)
    m = cuda_module_for(case, sd)
    kind = O.guide_kind_for(mode)
    vids = [O.synth_inputs(T, 27, 27, kind, seed=xseed + b, dtype=torch.bfloat16) for b in range(B)]
    X = torch.stack([v[0] for v in vids]).cuda()
    E = torch.stack([v[1] for v in vids]).cuda() if kind else None
    G = torch.stack([v[2] for v in vids]).cuda() if kind else None
    return case, sd, m, vids, X, E, G


def _check(case, sd, vids, got, which):
    orc = oracle_for(case, sd, torch.float32)
    f = lambda t: None if t is None else t.float()
    torch.set_num_threads(torch.get_num_threads())
    for b in which:
        with torch.no_grad():
            truth = orc.forward(f(vids[b][0]), f(vids[b][1]), f(vids[b][2]), "video")
        out = got[b].float().cpu()
        assert out.shape == truth.shape
        err, cos = O.rel_err(out, truth), O.cosine(out, truth)
        print(f"{case.name} video {b}: rel_err {err:.3e} cos {cos:.6f}")
        assert cos >= BF16_COS and err <= BF16_REL, (case.name, b, err, cos)


@pytest.mark.parametrize("mode", ["coarse", "direct", None])
def test_c2_width3584_T16_batch32(mode, built_library):
    """BASELINE config 2 exactly as benchmarked: width 3584, 16 frames, bf16, batch 32 (the launch shapes of the headline
    number: CTA-pair probability pass over 2944 tiles, one-split pooling, CTA-pair readouts over 10368 rows)."""
    B = 32
    case, sd, m, vids, X, E, G = _setup(3584, 16, mode, B)
    with torch.no_grad():
        got = m.forward_batched(X, E, G, "video")
    assert got.shape == (B, 4 * 81 + 32, 3584) and got.dtype == torch.bfloat16
    _check(case, sd, vids, got, (0, 13, 31))
    if mode == "direct":  # projector.py:367-368: one query -> 32 identical global rows
        glob = got[:, -32:].float()
        assert float((glob - glob[:, :1]).abs().max()) == 0.0


def test_c3_width3584_T64_batch8(built_library):
    """BASELINE config 3 per GPU at 8 GPUs: width 3584, 64 frames, bf16, 8 videos (position tables grown past
    max_num_frames, pooling split over token ranges)."""
    B = 8
    case, sd, m, vids, X, E, G = _setup(3584, 64, "coarse", B)
    with torch.no_grad():
        got = m.forward_batched(X, E, G, "video")
    assert got.shape == (B, 16 * 81 + 32, 3584)
    _check(case, sd, vids, got, (0, 7))


def test_c5_width1536_T32_batch8(built_library):
    """BASELINE config 5 shapes: Qwen2.5-1.5B width 1536, 32 frames, bf16."""
    B = 8
    case, sd, m, vids, X, E, G = _setup(1536, 32, "coarse", B)
    with torch.no_grad():
        got = m.forward_batched(X, E, G, "video")
    assert got.shape == (B, 8 * 81 + 32, 1536)
    _check(case, sd, vids, got, (0, 5))


def test_c4_width3584_T512_long_video(built_library):
    """BASELINE config 4: one 512-frame video (373 248 tokens), bf16 — unsharded, and cut into 8 frame shards on ONE GPU
    whose split-softmax partials are merged exactly as the 8 ranks of `dist.forward_frame_sharded` merge theirs (the
    NCCL exchange itself is covered by tests/test_gpu_dist.py on >= 2 GPUs)."""
    case, sd, m, vids, X, E, G = _setup(3584, 512, "coarse", 1)
    with torch.no_grad():
        got = m.forward_batched(X, E, G, "video")
    assert got.shape == (1, 128 * 81 + 32, 3584)
    _check(case, sd, vids, got, (0,))
    gc, lc = m.global_compressor, m.local_compressor
    with torch.no_grad():
        Qg = gc.injected_query(G, 1, X.dtype)
        qf = gc.fold(Qg)
        parts, local = [], []
        for r in range(8):
            Xs, Es = X[:, 64 * r:64 * (r + 1)].contiguous(), E[:, 64 * r:64 * (r + 1)].contiguous()
            from hicom_b200 import ops
            parts.append(ops.softmax_reduce(*gc.partials(Xs, qf, t0=64 * r)))
            att = lc.attend(Xs, Es, G, "video")
            from hicom_b200.projector import _run_mlp
            local.append(_run_mlp(lc.readout, att))
        mm, ll, oo = (torch.cat([p[i] for p in parts], 1) for i in range(3))
        glob = torch.empty(32, 3584, dtype=X.dtype, device="cuda")
        gc.finish(Qg, mm, ll, oo, glob, 0, 0)
        sharded = torch.cat([torch.cat(local, 1)[0], glob], 0)
    assert O.rel_err(sharded.float().cpu(), got[0].float().cpu()) <= 8e-3
    _check(case, sd, vids, sharded.unsqueeze(0), (0,))
    # the exchange the ranks actually use for 16-bit models: each shard's own normalised attention rows after the value
    # projection + the log-sum-exp of its scores (75 KB), combined with softmax weights (ops.shard_combine)
    with torch.no_grad():
        msgs = []
        for r in range(8):
            Xs = X[:, 64 * r:64 * (r + 1)].contiguous()
            msgs.append(gc.shard_message(Qg, *gc.partials(Xs, qf, t0=64 * r)))
        msgs = torch.stack(msgs, 0)
        assert msgs.shape[2] == 32 * 1152 * 2 + 288 * 4
        a = ops.shard_combine(msgs, 32, 1152, gc.attn_layer.num_heads, X.dtype)
        glob2 = torch.empty(32, 3584, dtype=X.dtype, device="cuda")
        gc.finish_attended(Qg, a, glob2, 0, 0)
    assert O.rel_err(glob2.float().cpu(), got[0, -32:].float().cpu()) <= 8e-3


def test_c2_fp16_native(built_library):
    """The reference's inference dtype is fp16 (model/__init__.py:44, projector.py:52-53): c2 shapes in fp16 against the
    fp32 oracle on the fp16-rounded weights and inputs."""
    B = 4
    mode = "coarse"
    case = Case("full_3584_T16_fp16", PTYPE, mode, 16, 27, 27, 3584, "float16", wseed=11)
    sd = O.synth_state_dict(PTYPE, mode, 3584, seed=11, dtype=torch.float16)
    m = cuda_module_for(case, sd)
    vids = [O.synth_inputs(16, 27, 27, "vec", seed=500 + b, dtype=torch.float16) for b in range(B)]
    X = torch.stack([v[0] for v in vids]).cuda()
    E = torch.stack([v[1] for v in vids]).cuda()
    G = torch.stack([v[2] for v in vids]).cuda()
    with torch.inference_mode():
        got = m.forward_batched(X, E, G, "video")
    assert got.dtype == torch.float16 and got.shape == (B, 356, 3584)
    orc = oracle_for(case, sd, torch.float32)
    for b in (0, 3):
        truth = orc.forward(vids[b][0].float(), vids[b][1].float(), vids[b][2].float(), "video")
        err = O.rel_err(got[b].float().cpu(), truth)
        print(f"fp16 c2 video {b}: rel_err {err:.3e}")
        assert err <= 2e-3
