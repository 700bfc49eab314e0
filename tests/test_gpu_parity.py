"""GPU: the CUDA path (through the torch custom ops -> C ABI) against the golden reference outputs and the oracle.

Tolerances (BASELINE.json north_star / SURVEY §8c):
  fp32 : max|a-b| / max|b| <= 1e-4 vs the reference's fp32 output (stored golden)
  bf16 : vs the fp32 oracle on the bf16-rounded weights/inputs: cosine >= 0.999 and max|a-b|/max|b| <= 1e-2
Token order/layout is checked by the same comparison (a permuted row fails the tolerance) plus shape equality.
"""
import pytest
import torch

from oracle import hicom_oracle as O
from oracle.cases import CASES, CASES_BY_NAME, materialise

from util import cuda_module_for, oracle_for, to_dev, truth_fp32

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_REL, BF16_COS = 1e-2, 0.999


def run_cuda(case, batched=False):
    sd, X, E, g, nl = materialise(case)
    m = cuda_module_for(case, sd)
    with torch.no_grad():
        out = m(to_dev(X), to_dev(E), to_dev(g), case.modal, to_dev(nl))
    torch.cuda.synchronize()
    return out.float().cpu()


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_forward_matches_reference(case, golden, built_library):
    out = run_cuda(case)
    ref = golden[case.name]
    assert tuple(out.shape) == tuple(ref.shape)
    assert torch.isfinite(out).all()
    if case.dtype == "float32":
        assert O.rel_err(out, ref) <= FP32_TOL
    else:
        truth = truth_fp32(case)
        assert O.cosine(out, truth) >= BF16_COS
        assert O.rel_err(out, truth) <= BF16_REL
        # informational: how far the reference's own bf16 forward is from the same truth
        print(f"{case.name}: ours vs truth {O.rel_err(out, truth):.2e}; reference-bf16 vs truth "
              f"{O.rel_err(ref, truth):.2e}")


@pytest.mark.parametrize("name", ["coarse_T8", "none_T8", "direct_T8", "fine_T8", "adaptkv_coarse_T8"])
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_forward_batched_equals_loop(name, dtype, built_library):
    """forward_batched == stacking forward over the batch (hicom_arch.py:167-178)."""
    import dataclasses
    case = dataclasses.replace(CASES_BY_NAME[name], dtype=dtype)
    sd, _, _, _, _ = materialise(case)
    m = cuda_module_for(case, sd)
    kind = O.guide_kind_for(case.use_guide)
    dt = getattr(torch, dtype)
    xs, es, gs = [], [], []
    for b in range(3):
        X, E, g = O.synth_inputs(case.T, case.H, case.W, kind, seed=100 + b, dtype=dt)
        xs.append(X); es.append(E); gs.append(g)
    Xb = torch.stack(xs).cuda()
    Eb = torch.stack(es).cuda() if kind else None
    Gb = torch.stack(gs).cuda() if kind else None
    with torch.no_grad():
        got = m.forward_batched(Xb, Eb, Gb, "video")
        want = torch.stack([m(Xb[b], None if Eb is None else Eb[b], None if Gb is None else Gb[b], "video")
                            for b in range(3)])
    assert got.shape == want.shape
    assert O.rel_err(got.float().cpu(), want.float().cpu()) <= (1e-5 if dtype == "float32" else 1e-2)
    # and each video matches the oracle truth
    orc = oracle_for(case, sd, torch.float32)
    for b in range(3):
        f = lambda t: None if t is None else t.float()
        truth = orc.forward(f(xs[b]), f(es[b]), f(gs[b]), "video")
        err = O.rel_err(got[b].float().cpu(), truth)
        assert err <= (FP32_TOL if dtype == "float32" else BF16_REL)


def test_anyres_dict_branch(built_library):
    """any-res image dict input (projector.py:679-698): base -> local only, patch -> local + global."""
    case = CASES_BY_NAME["image_T1_newline"]
    sd, X, E, g, nl = materialise(case)
    gen = torch.Generator().manual_seed(3)
    patch = 0.5 * torch.randn(12, 6, 1152, generator=gen)
    patch_e = 0.5 * torch.randn(12, 6, 1152, generator=gen)
    want = oracle_for(case, sd).forward({"base": X[0], "patch": patch}, {"base": E[0], "patch": patch_e}, g,
                                        "image", nl)
    m = cuda_module_for(case, sd)
    with torch.no_grad():
        got = m({"base": X[0].cuda(), "patch": patch.cuda()}, {"base": E[0].cuda(), "patch": patch_e.cuda()},
                g.cuda(), "image", nl.cuda())
    assert got.shape == want.shape
    assert O.rel_err(got.float().cpu(), want) <= FP32_TOL


def test_reference_failures_are_loud(built_library):
    """T in {5,6,9} with temporal kernel 4 raise (the reference's torch.stack fails); bad guide rank raises."""
    case = CASES_BY_NAME["coarse_T8"]
    sd, X, E, g, nl = materialise(case)
    m = cuda_module_for(case, sd)
    with torch.no_grad():
        for T in (5, 6, 9):
            x = torch.randn(T, 6, 6, 1152, device="cuda")
            with pytest.raises(RuntimeError, match="unequal windows"):
                m(x, x, g.cuda(), "video")
        with pytest.raises(ValueError):
            m(X.cuda(), E.cuda(), torch.randn(4, 1152, device="cuda"), "video")  # coarse wants a (d,) vector
        with pytest.raises(ValueError):
            m(X.cuda(), E.cuda(), None, "video")
    # autograd-enabled calls on trainable parameters: gradients (hicom_b200/autograd.py), or — with the training path
    # switched off — a loud failure; never a tensor silently cut off from the graph
    assert m(X.cuda(), E.cuda(), g.cuda(), "video").requires_grad
    from hicom_b200 import autograd as ag
    ag.enable(False)
    try:
        with pytest.raises(RuntimeError, match="forward-only"):
            m(X.cuda(), E.cuda(), g.cuda(), "video")
    finally:
        ag.enable(True)
    with torch.inference_mode():
        assert m(X.cuda(), E.cuda(), g.cuda(), "video").shape[0] == 40


def test_full_size_properties_bf16(built_library):
    """BASELINE config c2 shape (width 3584, 16 frames, bf16), one video: size-independent properties.
    direct mode => 32 identical global rows; frame-sharded partial merge == unsharded; row count."""
    import dataclasses
    base = dataclasses.replace(CASES_BY_NAME["c1_direct_896_T16"], hidden=3584, dtype="bfloat16")
    sd, X, E, g, nl = materialise(base)
    m = cuda_module_for(base, sd)
    with torch.no_grad():
        out = m(X.cuda(), E.cuda(), g.cuda(), "video")
    assert out.shape == (4 * 81 + 32, 3584) and out.dtype == torch.bfloat16
    glob = out[-32:].float()
    assert float((glob - glob[0]).abs().max()) == 0.0
    # split-softmax: 2 frame shards merged == whole video
    gc = m.global_compressor
    Xd, gd = X.cuda().unsqueeze(0), g.cuda().unsqueeze(0)
    with torch.no_grad():
        Qg = gc.injected_query(gd, 1, Xd.dtype)
        qf = gc.fold(Qg)
        whole = torch.empty(32, 3584, dtype=Xd.dtype, device="cuda")
        gc.finish(Qg, *gc.partials(Xd, qf), whole, 0, 0)
        parts = [gc.partials(Xd[:, t0:t0 + 8].contiguous(), qf, t0=t0) for t0 in (0, 8)]
        mm = torch.cat([p[0] for p in parts], 1); ll = torch.cat([p[1] for p in parts], 1)
        oo = torch.cat([p[2] for p in parts], 1)
        merged = torch.empty_like(whole)
        gc.finish(Qg, mm, ll, oo, merged, 0, 0)
    # two valid bf16 evaluations (different softmax stabilisers round P differently): within two bf16 ulps of the
    # largest output; the fp32 version of this identity is checked to 1e-5 in test_frame_shard_merge_fp32
    assert O.rel_err(merged.float().cpu(), whole.float().cpu()) <= 8e-3
    assert O.rel_err(whole.float().cpu(), out[-32:].float().cpu()) <= 8e-3


def test_frame_shard_merge_fp32(built_library):
    """fp32: partials of 4 frame shards (position rows offset by t0) merged == unsharded, to rounding noise."""
    case = CASES_BY_NAME["coarse_27x27_T4"]
    import dataclasses
    case = dataclasses.replace(case, T=16, H=9, W=9)
    sd, X, E, g, nl = materialise(case)
    m = cuda_module_for(case, sd)
    gc = m.global_compressor
    Xd, gd = X.cuda().unsqueeze(0), g.cuda().unsqueeze(0)
    with torch.no_grad():
        Qg = gc.injected_query(gd, 1, Xd.dtype)
        qf = gc.fold(Qg)
        whole = torch.empty(32, case.hidden, dtype=Xd.dtype, device="cuda")
        gc.finish(Qg, *gc.partials(Xd, qf, splits=3), whole, 0, 0)
        parts = [gc.partials(Xd[:, t0:t0 + 4].contiguous(), qf, t0=t0, splits=2) for t0 in (0, 4, 8, 12)]
        merged = torch.empty_like(whole)
        gc.finish(Qg, *(torch.cat([p[i] for p in parts], 1) for i in range(3)), merged, 0, 0)
    assert O.rel_err(merged.cpu(), whole.cpu()) <= 1e-5
    truth = oracle_for(case, sd)._global(X, g)
    assert O.rel_err(whole.cpu(), truth) <= FP32_TOL


def test_graph_replay_and_host_pipeline_match_eager(built_library):
    """CUDA-graph replay (hicom_b200.graph) and the pinned-host chunked pipeline (hicom_b200.pipeline) return the same
    tokens as the eager batched forward."""
    import dataclasses
    from hicom_b200.graph import GraphedCompressor
    from hicom_b200.pipeline import compress_from_host
    case = dataclasses.replace(CASES_BY_NAME["coarse_27x27_T4"], dtype="bfloat16", T=8)
    sd, _, _, _, _ = materialise(case)
    m = cuda_module_for(case, sd)
    xs, es, gs = zip(*[O.synth_inputs(case.T, case.H, case.W, "vec", seed=40 + b, dtype=torch.bfloat16) for b in range(5)])
    Xh, Eh, Gh = (torch.stack(t).pin_memory() for t in (xs, es, gs))
    with torch.no_grad():
        want = m.forward_batched(Xh.cuda(), Eh.cuda(), Gh.cuda(), "video")
        g = GraphedCompressor(m, Xh.cuda(), Eh.cuda(), Gh.cuda())
        got_graph = g.replay().clone()
        # new data through the same graph
        got_graph2 = g(Xh.cuda().flip(0), Eh.cuda().flip(0), Gh.cuda().flip(0)).clone()
        got_host = compress_from_host(m, Xh, Eh, Gh, "video", chunk=2)
    torch.cuda.synchronize()
    assert g.kernels_per_replay > 10
    # fp32 atomics (row sums, per-frame mass) make runs differ in the last bf16 bit, so compare with a tolerance
    assert O.rel_err(got_graph.float().cpu(), want.float().cpu()) <= 8e-3
    assert O.rel_err(got_graph2.float().cpu(), want.flip(0).float().cpu()) <= 8e-3
    assert not got_host.is_cuda and got_host.shape == want.shape
    assert O.rel_err(got_host.float(), want.float().cpu()) <= 8e-3  # chunks of 2 take other split counts than B=5


def test_batched_caller_matches_per_sample_loop(built_library):
    """hicom_b200.caller.compress_samples == the reference's per-sample loop (hicom_arch.py:167-178)
    on a mixed batch: same-shape videos (grouped), an odd-length video and an any-res image dict (per item)."""
    from hicom_b200.caller import compress_samples
    case = CASES_BY_NAME["image_T1_newline"]  # spatial_unpad + newline layouts
    sd, _, _, _, nl = materialise(case)
    m = cuda_module_for(case, sd)
    mk = lambda T, s: tuple(t.cuda() for t in O.synth_inputs(T, 6, 6, "vec", seed=s))
    v = [mk(8, 1), mk(8, 2), mk(4, 3), mk(8, 4)]
    gen = torch.Generator().manual_seed(9)
    img = {"base": (0.5 * torch.randn(6, 6, 1152, generator=gen)).cuda(),
           "patch": (0.5 * torch.randn(12, 6, 1152, generator=gen)).cuda()}
    img_e = {"base": (0.5 * torch.randn(6, 6, 1152, generator=gen)).cuda(),
             "patch": (0.5 * torch.randn(12, 6, 1152, generator=gen)).cuda()}
    feats = [v[0][0], img, v[1][0], v[2][0], v[3][0]]
    embeds = [v[0][1], img_e, v[1][1], v[2][1], v[3][1]]
    guides = [v[0][2], v[1][2], v[1][2], v[2][2], v[3][2]]
    modal = ["video", "image", "video", "video", "video"]
    nl = nl.cuda()
    with torch.no_grad():
        got = compress_samples(m, feats, embeds, guides, modal, nl)
        want = [m(f, e, g, md, nl) for f, e, g, md in zip(feats, embeds, guides, modal)]
    assert [tuple(t.shape) for t in got] == [tuple(t.shape) for t in want]
    for a, b in zip(got, want):
        assert O.rel_err(a.cpu(), b.cpu()) <= 1e-5


def test_batched_caller_views_the_tower_output_in_place(built_library, monkeypatch):
    """hicom_arch.py:162-164: the per-sample tensors are `split` views of ONE tower output.  compress_samples must hand
    forward_batched a batch VIEW of that allocation (no 27 MB-per-video stack copy) and still equal the loop."""
    from hicom_b200.caller import compress_samples
    case = CASES_BY_NAME["bf16_coarse_T8"]
    sd, _, _, _, _ = materialise(case)
    m = cuda_module_for(case, sd)
    B, T = 4, case.T
    gen = torch.Generator().manual_seed(21)
    tower_x = (0.5 * torch.randn(B * T, case.H, case.W, 1152, generator=gen)).to(torch.bfloat16).cuda()
    tower_e = (0.5 * torch.randn(B * T, case.H, case.W, 1152, generator=gen)).to(torch.bfloat16).cuda()
    guides_t = (0.5 * torch.randn(B, 1152, generator=gen)).to(torch.bfloat16).cuda()
    feats, embeds = tower_x.split([T] * B, dim=0), tower_e.split([T] * B, dim=0)
    guides = [guides_t[i] for i in range(B)]
    seen = {}
    real = m.forward_batched

    def spy(X, E, G, *a, **k):
        seen["ptrs"] = (X.data_ptr(), E.data_ptr(), G.data_ptr())
        return real(X, E, G, *a, **k)

    monkeypatch.setattr(m, "forward_batched", spy)
    with torch.no_grad():
        got = compress_samples(m, feats, embeds, guides, ["video"] * B)
        assert seen["ptrs"] == (tower_x.data_ptr(), tower_e.data_ptr(), guides_t.data_ptr())  # views, not copies
        monkeypatch.undo()
        want = [m(feats[i], embeds[i], guides[i], "video") for i in range(B)]
    for a, b in zip(got, want):
        assert O.rel_err(a.float().cpu(), b.float().cpu()) <= 1e-2


@pytest.mark.parametrize("name", ["coarse_T8", "direct_T8", "fine_T8", "image_T1_newline"])
def test_fp16_inference_dtype(name, built_library):
    """The reference's inference path is fp16 (model/__init__.py:44; mm_infer casts inputs to fp16): fp16 weights and
    inputs must work and match the fp32 oracle on the fp16-rounded values to fp16 rounding of the output."""
    import dataclasses
    case = dataclasses.replace(CASES_BY_NAME[name], dtype="float16")
    sd, X, E, g, nl = materialise(case)
    m = cuda_module_for(case, sd)
    assert next(m.parameters()).dtype == torch.float16
    with torch.inference_mode():
        out = m(to_dev(X), to_dev(E), to_dev(g), case.modal, to_dev(nl))
    assert out.dtype == torch.float16
    assert set(m.state_dict().keys()) == set(sd.keys())
    truth = truth_fp32(case)
    assert out.shape == truth.shape
    assert O.rel_err(out.float().cpu(), truth) <= 2e-3
    # eval-mode inference OUTSIDE no_grad with (still trainable) fp16 parameters, as a caller of the reference may do: the
    # same tokens, no autograd graph (fp16 is the inference dtype; the training path is bf16 / fp32)
    assert any(p.requires_grad for p in m.parameters())
    out2 = m(to_dev(X), to_dev(E), to_dev(g), case.modal, to_dev(nl))
    assert not out2.requires_grad and torch.equal(out2, out)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("guide", ["coarse", "fine", None])
def test_clip_l_tower_dims(guide, dtype, built_library):
    """CLIP-L tower: qk_dim 768, 6 heads of 128, 24x24 patches (projector.py:407-414, 577-579) — 192 score columns."""
    import hicom_b200
    from util import Cfg
    ptype, hidden, T, H, W, d = "local43_global32", 128, 8, 24, 24, 768
    sd = O.synth_state_dict(ptype, guide, hidden, seed=3, dtype=dtype, d=d)
    X, E, g = O.synth_inputs(T, H, W, O.guide_kind_for(guide), seed=4, d=d, dtype=dtype)
    m = hicom_b200.build_vision_projector(Cfg(mm_vision_tower="openai/clip-vit-large-patch14-336", mm_hidden_size=d,
                                              hidden_size=hidden, use_guide=guide, mm_projector_type=ptype,
                                              max_num_frames=8))
    m.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
    m = m.to(dtype).cuda().eval()
    f = lambda t: None if t is None else t.float()
    orc = O.OracleProjector(ptype, guide, state={k: v.float() for k, v in sd.items()}, qk_dim=d)
    with torch.no_grad():
        want = orc.forward(f(X), f(E), f(g), "video")
        got = m(to_dev(X), to_dev(E), to_dev(g), "video").float().cpu()
    assert got.shape == want.shape == (2 * 64 + 32, hidden)
    if dtype == torch.float32:
        assert O.rel_err(got, want) <= 1e-4
    else:
        assert O.cosine(got, want) >= 0.999 and O.rel_err(got, want) <= 1e-2


@pytest.mark.parametrize("name", ["coarse_T8", "direct_T8", "fine_T8", "bf16_coarse_T8", "adaptkv_coarse_T8",
                                  "adaptqkvg_coarse_T4", "adaptqkvg_fine_T8"])
def test_local_clip_scale(name, built_library):
    """use_clip_scale='local' (projector.py:527-529, 547-549): L2-normalised keys and guide, exp(logit_scale) logits.
    The SigLIP scalars are set on the module directly (the reference pulls them from the hub, :661-663).  With the key
    adapter (adaptk) the reference normalises frames_embed BEFORE the adapter mix (:528 then :533)."""
    case = CASES_BY_NAME[name]
    sd, X, E, g, nl = materialise(case)
    m = cuda_module_for(case, sd)
    ls, lb = torch.tensor(2.0), torch.tensor(-5.0)
    m.local_logit_scale, m.local_logit_bias = ls.cuda(), lb.cuda()
    f = lambda t: None if t is None else t.float()
    orc = oracle_for(case, sd, torch.float32)
    orc.local_logit = (ls, lb)
    with torch.no_grad():
        want = orc.forward(f(X), f(E), f(g), case.modal, f(nl))
        got = m(to_dev(X), to_dev(E), to_dev(g), case.modal, to_dev(nl)).float().cpu()
    if case.dtype == "float32":
        assert O.rel_err(got, want) <= FP32_TOL
    else:
        assert O.cosine(got, want) >= BF16_COS and O.rel_err(got, want) <= BF16_REL


@pytest.mark.parametrize("name", ["coarse_T8", "direct_T8", "fine_T8", "none_T8", "bf16_coarse_T8", "bf16_fine_T8"])
def test_global_clip_scale(name, built_library):
    """use_clip_scale='local,global' (projector.py:184-188): q and k L2-normalised over all 1152 channels before the head
    split, logits scaled by exp(logit_scale).  The keys are computed explicitly (k_proj + position term + row norms) and
    scored against head-masked normalised queries; the value side stays reassociated."""
    case = CASES_BY_NAME[name]
    sd, X, E, g, nl = materialise(case)
    m = cuda_module_for(case, sd)
    ls, lb = torch.tensor(2.0), torch.tensor(-5.0)
    m.global_logit_scale, m.global_logit_bias = ls.cuda(), lb.cuda()
    if E is not None:  # 'local' needs frames_embed; guide-less configs only exercise the global half
        m.local_logit_scale, m.local_logit_bias = ls.cuda(), lb.cuda()
    f = lambda t: None if t is None else t.float()
    orc = oracle_for(case, sd, torch.float32)
    orc.global_logit = (ls, lb)
    if E is not None:
        orc.local_logit = (ls, lb)
    with torch.no_grad():
        want = orc.forward(f(X), f(E), f(g), case.modal, f(nl))
        got = m(to_dev(X), to_dev(E), to_dev(g), case.modal, to_dev(nl)).float().cpu()
        B2 = m.forward_batched(torch.stack([to_dev(X)] * 2), None if E is None else torch.stack([to_dev(E)] * 2),
                               None if g is None else torch.stack([to_dev(g)] * 2), case.modal, to_dev(nl)).float().cpu()
    if case.dtype == "float32":
        assert O.rel_err(got, want) <= FP32_TOL and O.rel_err(B2[1], want) <= FP32_TOL
    else:
        assert O.cosine(got, want) >= BF16_COS and O.rel_err(got, want) <= BF16_REL
        assert O.rel_err(B2[1], want) <= BF16_REL


def test_forward_with_cuda_graph_cache(built_library):
    """enable_cuda_graphs(): the reference's per-video call pattern replays cached graphs; results equal the eager
    path, outputs are fresh tensors, entries are keyed by shape and evicted least-recently-used."""
    case = CASES_BY_NAME["bf16_coarse_T8"]
    sd, X, E, g, nl = materialise(case)
    m = cuda_module_for(case, sd)
    Xd, Ed, gd = to_dev(X), to_dev(E), to_dev(g)
    with torch.no_grad():
        want8 = m(Xd, Ed, gd, "video").clone()
        want4 = m(Xd[:4].contiguous(), Ed[:4].contiguous(), gd, "video").clone()
        m.enable_cuda_graphs(max_entries=1)
        a = m(Xd, Ed, gd, "video")
        b = m(Xd, Ed, gd, "video")                       # replay of the cached graph
        assert a.data_ptr() != b.data_ptr() and len(m.__dict__["_graphs"]) == 1
        c = m(Xd[:4].contiguous(), Ed[:4].contiguous(), gd, "video")   # second shape evicts the first
        assert len(m.__dict__["_graphs"]) == 1
        d = m(Xd * 0.5, Ed, gd, "video")                 # new values through the static buffers of a re-captured graph
        m.disable_cuda_graphs()
        want_half = m(Xd * 0.5, Ed, gd, "video")
    for got, want in ((a, want8), (b, want8), (c, want4), (d, want_half)):
        assert got.shape == want.shape
        assert O.rel_err(got.float().cpu(), want.float().cpu()) <= 1e-5


@pytest.mark.parametrize("name", ["bf16_coarse_T8", "video_grid_newline", "direct_T8", "video_one_token"])
def test_forward_batched_writes_into_padded_buffer(name, built_library):
    """Token splice (hicom_arch.py:283-373): the tokens are stored straight into rows [off, off+n) of every sample of a
    caller-owned padded (B, L, Dh) buffer; all other rows stay untouched."""
    case = CASES_BY_NAME[name]
    sd, X, E, g, nl = materialise(case)
    m = cuda_module_for(case, sd)
    B = 3
    st = lambda t: None if t is None else torch.stack([to_dev(t)] * B)
    with torch.no_grad():
        want = m.forward_batched(st(X), st(E), st(g), case.modal, to_dev(nl))
        n, Dh = want.shape[1], want.shape[2]
        off, L = 11, n + 37
        buf = torch.full((B, L, Dh), 7.0, dtype=want.dtype, device="cuda")
        got = m.forward_batched(st(X), st(E), st(g), case.modal, to_dev(nl), out=buf, out_row_offset=off)
    assert got.data_ptr() == buf[:, off:].data_ptr() and got.shape == want.shape
    assert torch.equal(buf[:, off:off + n], want)
    assert bool((buf[:, :off] == 7.0).all()) and bool((buf[:, off + n:] == 7.0).all())
    with torch.no_grad(), pytest.raises(ValueError):
        m.forward_batched(st(X), st(E), st(g), case.modal, to_dev(nl), out=buf, out_row_offset=L - n + 1)


@pytest.mark.parametrize("name,where", [("coarse_T8", "local,global"), ("fine_T8", "local,global"), ("none_T8", "global"),
                                        ("adaptkv_coarse_T8", "local")])
def test_clip_scale_fp16(name, where, built_library):
    """use_clip_scale in the reference's inference dtype: fp16 explicit keys / normalised rows through the native fp16
    path (HICOM_F16), against the fp32 oracle on the fp16-rounded values."""
    import dataclasses
    case = dataclasses.replace(CASES_BY_NAME[name], dtype="float16")
    sd, X, E, g, nl = materialise(case)
    m = cuda_module_for(case, sd)
    orc = oracle_for(case, sd, torch.float32)
    ls, lb = torch.tensor(2.0), torch.tensor(-5.0)
    for part in where.split(","):
        setattr(m, f"{part}_logit_scale", ls.cuda())
        setattr(m, f"{part}_logit_bias", lb.cuda())
        setattr(orc, f"{part}_logit", (ls, lb))
    f = lambda t: None if t is None else t.float()
    with torch.inference_mode():
        got = m(to_dev(X), to_dev(E), to_dev(g), case.modal, to_dev(nl))
    assert got.dtype == torch.float16
    with torch.no_grad():
        want = orc.forward(f(X), f(E), f(g), case.modal, f(nl))
    # the key adapter on normalised keys (entries ~0.03 feeding an MLP + LayerNorm) under exp(2)-sharpened logits amplifies
    # 16-bit rounding of the intermediate rows: measured 8e-3 in fp16 and 3.3e-2 in bf16 on this case, 3e-6 in fp32
    tol = 2e-2 if name.startswith("adaptkv") else 4e-3
    assert O.cosine(got.float().cpu(), want) >= 0.9995 and O.rel_err(got.float().cpu(), want) <= tol
