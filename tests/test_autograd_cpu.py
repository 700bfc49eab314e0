"""CPU: the backward FORMULAS and view plumbing of hicom_b200/autograd.py, with every kernel replaced
by its torch-CPU stand-in (tests/cpu_ops.py), against PyTorch autograd through the oracle — i.e. against what the
reference's training step computes (train.py:704-738 trains mm_projector through autograd).  The kernels themselves
are compared with the same stand-ins on the GPU (tests/test_gpu_autograd.py)."""
import dataclasses

import pytest
import torch

import cpu_ops
from oracle import hicom_oracle as O
from oracle.cases import CASES_BY_NAME, materialise
from util import cfg_for

TRAIN_CASES = ["none_T8", "direct_T8", "coarse_T8", "off_T4", "coarse_T7", "coarse_nondiv_7x8", "global_only_coarse_T8",
               "local22_global8_T8", "forced_guide", "video_grid_newline", "video_frame_newline", "video_one_token",
               "image_T1_newline", "none_T2_short_window",
               # fine mode (MHA over the instruction tokens) and every adapter
               "fine_T8", "local_only_fine_T8", "adaptkv_coarse_T8", "adaptqkvg_fine_T8", "adaptqkvg_coarse_T4"]


def _module(case, sd):
    import hicom_b200
    m = hicom_b200.build_vision_projector(cfg_for(case))
    m.load_state_dict(sd, strict=True)
    return m.train()


def _oracle_grads(case, sd, X, E, g, nl, probe):
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    nl_leaf = None if nl is None else nl.clone().requires_grad_(True)
    orc = O.OracleProjector(case.ptype, case.use_guide, case.merge, case.nlpos, leaf)
    out = orc.forward(X, E, g, case.modal, nl_leaf)
    if probe is None:                      # forward only (shape probing)
        return out.detach(), {}, None
    (out * probe).sum().backward()
    grads = {k: v.grad for k, v in leaf.items()}
    return out.detach(), grads, (None if nl_leaf is None else nl_leaf.grad)


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_parameter_gradients_match_reference_autograd(name, monkeypatch):
    from hicom_b200 import autograd as ag
    case = CASES_BY_NAME[name]
    sd, X, E, g, nl = materialise(case)
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(ag, "ENABLED", True)
    m = _module(case, sd)
    nl_leaf = None if nl is None else nl.clone().requires_grad_(True)
    out = m(X, E, g, case.modal, nl_leaf)                       # the reference's per-video forward signature
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    want_out, want, want_nl = _oracle_grads(case, sd, X, E, g, nl, probe)
    assert out.shape == want_out.shape
    assert O.rel_err(out.detach(), want_out) <= 2e-5
    (out * probe).sum().backward()
    got = {k: p.grad for k, p in m.named_parameters()}
    assert set(got) == set(want)
    for k in sorted(want):
        if want[k] is None or float(want[k].abs().max()) == 0.0:   # parameter unused by this mode (direct: query)
            assert got[k] is None or float(got[k].abs().max()) == 0.0, k
            continue
        assert got[k] is not None, f"no gradient reached {k}"
        if float(want[k].abs().max()) <= 1e-6:
            # a bias added to every key (k_proj.bias, the key adapter's k_norm.bias) shifts all scores of a query by one
            # constant, which the softmax cancels: the gradient is zero up to rounding noise (exactly zero for the
            # reassociated global attention)
            assert k.endswith(("k_proj.bias", "k_norm.bias")), k
            assert float(got[k].abs().max()) <= 1e-6
            if k == "global_compressor.attn_layer.k_proj.bias":
                assert float(got[k].abs().max()) == 0.0
            continue
        assert O.rel_err(got[k], want[k]) <= 2e-4, (k, O.rel_err(got[k], want[k]))
    if nl is not None:
        assert O.rel_err(nl_leaf.grad, want_nl) <= 1e-5
    kinds = {c[0] for c in cpu_ops.calls}
    assert "gemm" in kinds and "softmax_backward" in kinds or m.global_compressor is None


def test_batched_gradients_sum_over_videos(monkeypatch):
    """forward_batched under autograd == the per-video loop (hicom_arch.py:167-178): gradients add up over the batch."""
    from hicom_b200 import autograd as ag
    case = CASES_BY_NAME["coarse_T4"]
    sd, X, E, g, _ = materialise(case)
    X2, E2, g2 = (torch.stack([t, t.flip(0) * 0.9]) for t in (X, E, g))
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(ag, "ENABLED", True)
    m = _module(case, sd)
    out = m.forward_batched(X2, E2, g2, "video")
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(4))
    (out * probe).sum().backward()
    got = {k: p.grad.clone() for k, p in m.named_parameters()}
    total = None
    for b in range(2):
        _, gr, _ = _oracle_grads(case, sd, X2[b], E2[b], g2[b], None, probe[b])
        total = gr if total is None else {k: total[k] + gr[k] for k in gr}
    for k in sorted(total):
        if k == "global_compressor.attn_layer.k_proj.bias":     # exactly zero here, rounding noise in the reference
            assert float(total[k].abs().max()) <= 1e-6 and float(got[k].abs().max()) == 0.0
            continue
        assert O.rel_err(got[k], total[k]) <= 2e-4, k


def test_opt_in_and_unsupported_configs_fail_loudly(monkeypatch):
    from hicom_b200 import autograd as ag
    case = CASES_BY_NAME["coarse_T4"]
    sd, X, E, g, _ = materialise(case)
    cpu_ops.install(monkeypatch)
    m = _module(case, sd)
    monkeypatch.setattr(ag, "ENABLED", False)
    with pytest.raises(RuntimeError, match="forward-only"):
        m(X, E, g, "video")
    monkeypatch.setattr(ag, "ENABLED", True)
    with pytest.raises(NotImplementedError, match="dtype"):
        _module(case, sd).half()(X.half(), E.half(), g.half(), "video")   # fp16 is the inference dtype; train in bf16 / fp32
    # frozen projector under grad mode: nothing needs a graph -> the plain inference path (which refuses CPU tensors)
    m = _module(case, sd).requires_grad_(False)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        m(X, E, g, "video")


def test_bf16_parameter_gradients_have_parameter_dtype(monkeypatch):
    from hicom_b200 import autograd as ag
    case = dataclasses.replace(CASES_BY_NAME["coarse_T4"], dtype="bfloat16")
    sd, X, E, g, _ = materialise(case)
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(ag, "ENABLED", True)
    m = _module(case, {k: v.float() for k, v in sd.items()}).to(torch.bfloat16)
    out = m(X, E, g, "video")
    assert out.dtype == torch.bfloat16
    out.float().sum().backward()
    for k, p in m.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.bfloat16 and bool(torch.isfinite(p.grad.float()).all()), k
    f32 = {k: v.float() for k, v in sd.items()}
    _, want, _ = _oracle_grads(dataclasses.replace(case, dtype="float32"), f32, X.float(), E.float(), g.float(), None,
                               torch.ones(out.shape))
    for k in ("local_compressor.readout.2.weight", "global_compressor.attn_layer.k_proj.weight",
              "global_compressor.query", "local_compressor.guide_injector.coarse_proj.0.weight"):
        got = dict(m.named_parameters())[k].grad.float()
        assert O.cosine(got, want[k]) >= 0.99, (k, O.cosine(got, want[k]))


@pytest.mark.parametrize("name", ["direct_T8", "coarse_T8", "coarse_nondiv_7x8", "video_one_token", "adaptkv_direct_T8",
                                  "adaptqkvg_fine_T8"])
def test_stage3_gradients_reach_frames_embed_and_guide(name, monkeypatch):
    """Stage 3 of the release recipe tunes vision_model_head and guide_encoder (train.py:717-726): frames_embed (the
    local keys) and the instruction embedding then require gradients."""
    from hicom_b200 import autograd as ag
    case = (dataclasses.replace(CASES_BY_NAME["adaptkv_coarse_T8"], use_guide="direct")  # the second release recipe
            if name == "adaptkv_direct_T8" else CASES_BY_NAME[name])
    sd, X, E, g, nl = materialise(case)
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(ag, "ENABLED", True)
    m = _module(case, sd)
    E1, g1 = E.clone().requires_grad_(True), g.clone().requires_grad_(True)
    out = m(X, E1, g1, case.modal, nl)
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    (out * probe).sum().backward()
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    E2, g2 = E.clone().requires_grad_(True), g.clone().requires_grad_(True)
    want = O.OracleProjector(case.ptype, case.use_guide, case.merge, case.nlpos, leaf).forward(X, E2, g2, case.modal, nl)
    (want * probe).sum().backward()
    assert O.rel_err(E1.grad, E2.grad) <= 2e-4 and O.rel_err(g1.grad, g2.grad) <= 2e-4
    for k, p in m.named_parameters():
        if leaf[k].grad is not None and float(leaf[k].grad.abs().max()) > 1e-6:
            assert O.rel_err(p.grad, leaf[k].grad) <= 2e-4, k
    assert any(c[0] == "local_attend_backward" and c[1] and c[2] for c in cpu_ops.calls)


def test_producer_head_gradients(monkeypatch):
    """vision_model_head in mm_tunable_parts: layernorm / fc1 / fc2 of the SigLIP head receive gradients through
    frames_embed = h + mlp(layernorm(h)) (encoder.py:284-285)."""
    from hicom_b200 import autograd as ag
    from hicom_b200.producer import SiglipHeadEmbed
    from oracle import siglip_head as SH
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(ag, "ENABLED", True)
    sd = SH.synth_head_state(3, hidden=128, inter=256)
    m = SiglipHeadEmbed(128, 256)
    m.load_state_dict(sd, strict=True)
    h = SH.synth_hidden(2, 16, seed=2, hidden=128)
    out = m(h)
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(6))
    (out * probe).sum().backward()
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    (SH.image_embeds(leaf, h, side=4) * probe).sum().backward()
    for k, p in m.named_parameters():
        assert O.rel_err(p.grad, leaf[k].grad) <= 2e-4, k


def test_batched_caller_and_anyres_dict_under_autograd(monkeypatch):
    """compress_samples (the replacement of the loop of hicom_arch.py:167-178, which is also the forward of a training
    step) keeps the autograd graph: a mixed batch — grouped videos, an odd-length video, an any-res image dict with the
    newline layouts — gives the gradients of the reference-style per-sample loop through the oracle."""
    from hicom_b200 import autograd as ag
    from hicom_b200.caller import compress_samples
    case = CASES_BY_NAME["image_T1_newline"]
    sd, _, _, _, nl = materialise(case)
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(ag, "ENABLED", True)
    m = _module(case, sd)
    mk = lambda T, s: O.synth_inputs(T, 6, 6, "vec", seed=s)
    v = [mk(8, 1), mk(8, 2), mk(4, 3)]
    gen = torch.Generator().manual_seed(9)
    rnd = lambda *s: 0.5 * torch.randn(*s, generator=gen)
    img, img_e = {"base": rnd(6, 6, 1152), "patch": rnd(12, 6, 1152)}, {"base": rnd(6, 6, 1152), "patch": rnd(12, 6, 1152)}
    feats, embeds = [v[0][0], img, v[1][0], v[2][0]], [v[0][1], img_e, v[1][1], v[2][1]]
    guides, modal = [v[0][2], v[1][2], v[1][2], v[2][2]], ["video", "image", "video", "video"]
    nl_leaf = nl.clone().requires_grad_(True)
    got = compress_samples(m, feats, embeds, guides, modal, nl_leaf)
    probes = [torch.randn(t.shape, generator=torch.Generator().manual_seed(i)) for i, t in enumerate(got)]
    sum((t * p).sum() for t, p in zip(got, probes)).backward()
    leaf = {k: w.clone().requires_grad_(True) for k, w in sd.items()}
    nl2 = nl.clone().requires_grad_(True)
    orc = O.OracleProjector(case.ptype, case.use_guide, case.merge, case.nlpos, leaf)
    want = [orc.forward(f, e, g, md, nl2) for f, e, g, md in zip(feats, embeds, guides, modal)]
    sum((t * p).sum() for t, p in zip(want, probes)).backward()
    for a, b in zip(got, want):
        assert a.shape == b.shape and O.rel_err(a.detach(), b.detach()) <= 2e-5
    for k, p in m.named_parameters():
        if float(leaf[k].grad.abs().max()) <= 1e-6:
            continue
        assert O.rel_err(p.grad, leaf[k].grad) <= 2e-4, k
    assert O.rel_err(nl_leaf.grad, nl2.grad) <= 1e-5


@pytest.mark.parametrize("name", ["coarse_T4", "adaptqkvg_fine_T8"])
def test_fp32_master_weights_with_bf16_activations(name, monkeypatch):
    """HF Trainer with --bf16 and no DeepSpeed bf16 engine keeps fp32 master weights and feeds bf16 activations
    (torch.autocast semantics): parameters are cast per use, every kernel still sees one dtype, gradients come back fp32."""
    from hicom_b200 import autograd as ag
    case = CASES_BY_NAME[name]
    sd, X, E, g, _ = materialise(case)
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(ag, "ENABLED", True)
    m = _module(case, sd)                                        # fp32 parameters
    out = m(X.bfloat16(), E.bfloat16(), g.bfloat16(), case.modal)
    assert out.dtype == torch.bfloat16
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    (out.float() * probe).sum().backward()
    _, want, _ = _oracle_grads(case, sd, X.bfloat16().float(), E.bfloat16().float(), g.bfloat16().float(), None, probe)
    for k, p in m.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32, k
        if float(want[k].abs().max()) > 1e-6:
            assert O.cosine(p.grad, want[k]) >= 0.99, (k, O.cosine(p.grad, want[k]))


def test_producer_fp32_master_weights_with_bf16_activations(monkeypatch):
    from hicom_b200 import autograd as ag
    from hicom_b200.producer import SiglipHeadEmbed
    from oracle import siglip_head as SH
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(ag, "ENABLED", True)
    m = SiglipHeadEmbed(128, 256)
    m.load_state_dict(SH.synth_head_state(3, hidden=128, inter=256), strict=True)
    out = m(SH.synth_hidden(2, 16, seed=2, hidden=128).bfloat16())
    out.float().sum().backward()
    assert out.dtype == torch.bfloat16 and all(p.grad is not None and p.grad.dtype == torch.float32 for p in m.parameters())


def test_partially_frozen_projector_skips_unneeded_backward(monkeypatch):
    """Only the readouts trainable (everything else frozen): gradients reach exactly those parameters and the attention
    backward blocks are never launched (needs_input_grad prunes them)."""
    from hicom_b200 import autograd as ag
    case = CASES_BY_NAME["coarse_T4"]
    sd, X, E, g, _ = materialise(case)
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(ag, "ENABLED", True)
    m = _module(case, sd)
    for k, p in m.named_parameters():
        p.requires_grad_(".readout." in k)
    out = m(X, E, g, "video")
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    cpu_ops.calls.clear()
    (out * probe).sum().backward()
    _, want, _ = _oracle_grads(case, sd, X, E, g, None, probe)
    for k, p in m.named_parameters():
        if ".readout." in k:
            assert O.rel_err(p.grad, want[k]) <= 2e-4, k
        else:
            assert p.grad is None, k
    kinds = {c[0] for c in cpu_ops.calls}
    assert not kinds & {"softmax_backward", "local_attend_backward", "film_layernorm_backward", "global_attend_partial"}


def test_reference_errors_survive_under_autograd(monkeypatch):
    """Bad guide inputs raise the reference's exception types on the training path too (projector.py:350,362,386)."""
    from hicom_b200 import autograd as ag
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(ag, "ENABLED", True)
    for name in ("direct_T8", "coarse_T8", "fine_T8"):
        case = CASES_BY_NAME[name]
        sd, X, E, g, _ = materialise(case)
        m = _module(case, sd)
        with pytest.raises(ValueError):
            m(X, E, None, "video")                                  # guide missing
        wrong = torch.randn(3, 1152) if g.dim() == 1 else torch.randn(1152)
        with pytest.raises(ValueError):
            m(X, E, wrong, "video")                                 # wrong guide rank


def _random_train_cases(n, seed=7):
    import random
    from oracle.cases import Case
    rng = random.Random(seed)
    out = []
    for i in range(n):
        T = rng.choice([1, 2, 4, 7, 8])
        H, W = rng.choice([(6, 6), (7, 8), (5, 6), (9, 9)])
        adapt = "".join(ch for ch in "qkvg" if rng.random() < 0.4)
        ptype = rng.choice(["local43{a}_global8{g}", "local22{a}_global4{g}", "local43{a}", "global8{g}"]).format(
            a=("_adapt" + adapt) if adapt else "", g="_adaptg" if rng.random() < 0.4 else "")
        guide = rng.choice([None, "direct", "coarse", "fine"])
        modal = "image" if (T == 1 and rng.random() < 0.5) else "video"
        merge, nlpos = rng.choice([("flat", "one_token"), ("spatial_unpad", "grid"), ("spatial_unpad", "frame"),
                                   ("spatial_unpad", "one_token"), ("spatial_unpad", "no_token")])
        out.append(Case(f"fuzz{i}_{ptype}_{guide}_T{T}_{H}x{W}_{modal}_{nlpos}", ptype, guide, T, H, W, 64, "float32", modal,
                        merge, nlpos, merge != "flat", wseed=10 + i, xseed=20 + i))
    return out


@pytest.mark.parametrize("case", _random_train_cases(16), ids=lambda c: c.name)
def test_training_path_fuzz_against_reference_autograd(case, monkeypatch):
    """Seeded random configurations (grids, frame counts, guide modes, adapters, newline layouts, image modal): outputs
    and every parameter gradient equal PyTorch autograd through the oracle; shapes the reference cannot stack raise the
    same RuntimeError on both sides."""
    from hicom_b200 import autograd as ag
    sd, X, E, g, nl = materialise(case)
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(ag, "ENABLED", True)
    m = _module(case, sd)
    probe_seed = torch.Generator().manual_seed(3)
    try:
        want_out, want, _ = _oracle_grads(case, sd, X, E, g, nl, None)
    except RuntimeError:
        with pytest.raises(RuntimeError):
            m(X, E, g, case.modal, nl)
        return
    out = m(X, E, g, case.modal, nl)
    assert out.shape == want_out.shape and O.rel_err(out.detach(), want_out) <= 2e-5
    probe = torch.randn(out.shape, generator=probe_seed)
    (out * probe).sum().backward()
    _, want, _ = _oracle_grads(case, sd, X, E, g, nl, probe)
    for k, p in m.named_parameters():
        w = want[k]
        if w is None or float(w.abs().max()) <= 1e-6:
            assert p.grad is None or float(p.grad.abs().max()) <= 1e-6, k
        else:
            assert p.grad is not None and O.rel_err(p.grad, w) <= 3e-4, (k, O.rel_err(p.grad, w))


@pytest.mark.parametrize("name", ["none_T8", "direct_T8", "coarse_T8", "coarse_nondiv_7x8", "fine_T8", "adaptqkvg_coarse_T4",
                                  "global_only_coarse_T8"])
def test_gradients_into_frames_feature(name, monkeypatch):
    """mm_tunable_parts 'pure_vision_model' (train.py:712-715): the SigLIP body is tuned, so frames_feature carries a
    gradient — through the window values (and keys when frames_embed is None), the trilinear grid pooling of the query
    and the global attention's pooled operand x' = x + pos_embed."""
    from hicom_b200 import autograd as ag
    case = CASES_BY_NAME[name]
    sd, X, E, g, nl = materialise(case)
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(ag, "ENABLED", True)
    m = _module(case, sd)
    X1 = X.clone().requires_grad_(True)
    out = m(X1, E, g, case.modal, nl)
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(8))
    (out * probe).sum().backward()
    X2 = X.clone().requires_grad_(True)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = O.OracleProjector(case.ptype, case.use_guide, case.merge, case.nlpos, leaf).forward(X2, E, g, case.modal, nl)
    (want * probe).sum().backward()
    assert O.rel_err(out.detach(), want.detach()) <= 2e-5
    assert X1.grad is not None and O.rel_err(X1.grad, X2.grad) <= 2e-4, O.rel_err(X1.grad, X2.grad)
    for k, p in m.named_parameters():
        if leaf[k].grad is not None and float(leaf[k].grad.abs().max()) > 1e-6:
            assert O.rel_err(p.grad, leaf[k].grad) <= 2e-4, k


@pytest.mark.parametrize("name,where", [("coarse_T8", "local"), ("coarse_T8", "global"), ("direct_T8", "local,global"),
                                        ("none_T8", "global"), ("adaptkv_coarse_T8", "local,global"),
                                        ("fine_T8", "local,global")])
def test_clip_scale_is_differentiable(name, where, monkeypatch):
    """use_clip_scale (projector.py:184-188,527-529,548-549) with mm_tunable_parts 'attn_scale' (train.py:729-732):
    logit_scale / logit_bias are trainable parameters of the projector; gradients of every parameter, of both scalars
    and of frames_embed against PyTorch autograd through the oracle."""
    from hicom_b200 import autograd as ag
    case = CASES_BY_NAME[name]
    sd, X, E, g, nl = materialise(case)
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(ag, "ENABLED", True)
    m = _module(case, sd)
    scal = {}
    for part in where.split(","):
        ls, lb = torch.nn.Parameter(torch.tensor(1.3)), torch.nn.Parameter(torch.tensor(-0.7))
        setattr(m, f"{part}_logit_scale", ls)
        setattr(m, f"{part}_logit_bias", lb)
        scal[part] = (ls, lb)
    E1 = None if E is None else E.clone().requires_grad_(True)
    out = m(X, E1, g, case.modal, nl)
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(9))
    (out * probe).sum().backward()
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    orc = O.OracleProjector(case.ptype, case.use_guide, case.merge, case.nlpos, leaf)
    oscal = {}
    for part in where.split(","):
        oscal[part] = (torch.tensor(1.3, requires_grad=True), torch.tensor(-0.7, requires_grad=True))
        setattr(orc, f"{part}_logit", oscal[part])
    E2 = None if E is None else E.clone().requires_grad_(True)
    want = orc.forward(X, E2, g, case.modal, nl)
    (want * probe).sum().backward()
    assert O.rel_err(out.detach(), want.detach()) <= 2e-5
    for k, p in m.named_parameters():
        if k.endswith(("logit_scale", "logit_bias")):
            continue
        w = leaf[k].grad
        if w is None or float(w.abs().max()) <= 1e-6:
            continue
        assert p.grad is not None and O.rel_err(p.grad, w) <= 5e-4, (k, O.rel_err(p.grad, w))
    for part, (ls, lb) in scal.items():
        ols, olb = oscal[part]
        if ols.grad is not None and float(ols.grad.abs()) > 1e-7:
            assert ls.grad is not None and abs(float(ls.grad) - float(ols.grad)) <= 5e-4 * max(1.0, abs(float(ols.grad))), part
        assert lb.grad is None or abs(float(lb.grad)) <= 1e-6       # the softmax cancels the bias
    if E1 is not None and E2.grad is not None and E1.grad is not None:
        assert O.rel_err(E1.grad, E2.grad) <= 5e-4
