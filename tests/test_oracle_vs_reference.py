"""CPU, authoring container only: the oracle against the live reference module (skipped without /root/reference)."""
import pytest
import torch

from oracle import hicom_oracle as O
from oracle.cases import CASES, materialise
from oracle.ref_shim import load_reference, reference_available

from util import cfg_for, oracle_for

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")

SMALL = [c for c in CASES if c.H <= 9]


@pytest.mark.parametrize("case", SMALL, ids=[c.name for c in SMALL])
def test_live_reference(case):
    ref = load_reference()
    sd, X, E, g, nl = materialise(case)
    m = ref.build_vision_projector(cfg_for(case))
    m.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
    m = m.to(getattr(torch, case.dtype)).eval()
    with torch.no_grad():
        want = m(X, E, g, case.modal, nl)
        got = oracle_for(case, sd).forward(X, E, g, case.modal, nl)
    assert got.shape == want.shape
    assert O.rel_err(got.float(), want.float()) <= (2e-6 if case.dtype == "float32" else 8e-3)


@pytest.mark.parametrize("ptype", ["local43_global32_coarse", "local412_global8", "local43_adaptqk_global32_adaptg",
                                   "local43guidefine_global32", "global16", "local22", "mlp2x_gelu", "linear"])
@pytest.mark.parametrize("use_guide", [None, "direct", "coarse", "fine"])
def test_parser_and_state_dict_names(ptype, use_guide):
    from util import Cfg
    ref = load_reference()
    m = ref.build_vision_projector(Cfg(mm_projector_type=ptype, use_guide=use_guide, hidden_size=64, max_num_frames=2))
    if ptype in ("mlp2x_gelu", "linear"):
        assert O.parse_projector_type(ptype).kind in ("mlp", "linear")
        return
    want = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert O.param_shapes(ptype, use_guide, 64) == want


def test_reference_dict_branch_anyres():
    """any-res image dict input (projector.py:679-698)."""
    from oracle.cases import CASES_BY_NAME
    ref = load_reference()
    case = CASES_BY_NAME["image_T1_newline"]
    sd, X, E, g, nl = materialise(case)
    m = ref.build_vision_projector(cfg_for(case))
    m.load_state_dict(sd, strict=True)
    gen = torch.Generator().manual_seed(3)
    patch = 0.5 * torch.randn(12, 6, 1152, generator=gen)
    patch_e = 0.5 * torch.randn(12, 6, 1152, generator=gen)
    feats = {"base": X[0], "patch": patch}
    embeds = {"base": E[0], "patch": patch_e}
    with torch.no_grad():
        want = m(feats, embeds, g, "image", nl)
        got = oracle_for(case, sd).forward(feats, embeds, g, "image", nl)
    assert torch.equal(got, want)


@pytest.mark.parametrize("use_guide", [None, "coarse", "fine"])
def test_live_reference_clip_l_dims(use_guide):
    """CLIP-L tower: qk_dim 768, 6 heads (projector.py:407-414, 577-579); pins the oracle's ``qk_dim``/``d`` knobs."""
    from util import Cfg
    ref = load_reference()
    ptype, hidden, d = "local43_global32", 64, 768
    sd = O.synth_state_dict(ptype, use_guide, hidden, seed=3, d=d)
    X, E, g = O.synth_inputs(8, 6, 6, O.guide_kind_for(use_guide), seed=4, d=d)
    m = ref.build_vision_projector(Cfg(mm_vision_tower="openai/clip-vit-large-patch14-336", mm_hidden_size=d,
                                       hidden_size=hidden, use_guide=use_guide, mm_projector_type=ptype,
                                       max_num_frames=8))
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        want = m.eval()(X, E, g, "video")
        got = O.OracleProjector(ptype, use_guide, state=sd, qk_dim=d).forward(X, E, g, "video")
    assert got.shape == want.shape
    assert O.rel_err(got, want) <= 2e-6


@pytest.mark.parametrize("use_guide", ["coarse", "direct", "fine", "adaptkv_coarse", "adaptqkvg_fine"])
def test_live_reference_clip_scale(use_guide):
    """use_clip_scale='local,global' (projector.py:527-529, 547-549, 184-188) with the SigLIP scalars passed in
    (the hub weights are unavailable offline); the adapter cases pin the order normalise -> key adapter (:528, :533)."""
    from oracle.cases import CASES_BY_NAME
    ref = load_reference()
    case = CASES_BY_NAME[f"{use_guide}_T8"]
    sd, X, E, g, nl = materialise(case)
    m = ref.build_vision_projector(cfg_for(case))
    m.load_state_dict(sd, strict=True)
    ls, lb = torch.tensor(2.0), torch.tensor(-5.0)
    m.local_logit_scale, m.local_logit_bias = ls, lb
    m.global_logit_scale, m.global_logit_bias = ls.clone(), lb.clone()
    orc = oracle_for(case, sd)
    orc.local_logit, orc.global_logit = (ls, lb), (ls, lb)
    with torch.no_grad():
        want = m.eval()(X, E, g, case.modal, nl)
        got = orc.forward(X, E, g, case.modal, nl)
    assert O.rel_err(got, want) <= 2e-6


def _random_type_strings(n, seed=0):
    """Seeded samples of the type-string mini-language (projector.py:231-304), incl. junk the reference ignores."""
    import random
    rng = random.Random(seed)
    modes = ["direct", "coarse", "fine", "off"]
    out = []
    for _ in range(n):
        parts = []
        if rng.random() < 0.8:
            loc = "local" + rng.choice(["43", "22", "412", "13", "23", "41"])
            if rng.random() < 0.5:
                loc += "_adapt" + "".join(rng.sample("qkvg", rng.randint(0, 4))) + rng.choice(["", "x", "_", "qz"])
            if rng.random() < 0.3:
                loc += "guide" + rng.choice(modes)
            parts.append(loc)
        if rng.random() < 0.8 or not parts:
            glo = "global" + rng.choice(["32", "8", "16", "1", "64"])
            if rng.random() < 0.4:
                glo += "_adaptg"
            if rng.random() < 0.3:
                glo += "guide" + rng.choice(modes)
            parts.append(glo)
        if rng.random() < 0.3:
            rng.shuffle(parts)
        s = "_".join(parts) + rng.choice(["", "", "_coarse", "_fine", "_v2", "_anyres"])
        out.append((s, rng.choice([None, "off", "direct", "coarse", "fine"])))
    return out


@pytest.mark.parametrize("ptype,use_guide", _random_type_strings(48))
def test_factory_grammar_fuzz_against_reference(ptype, use_guide):
    """The product factory builds the same modules (parameter names and shapes, guide modes, kernel sizes) as the
    reference's build_vision_projector for random type strings — or fails with the same exception type."""
    import hicom_b200
    from util import Cfg
    ref = load_reference()
    mk = lambda: Cfg(mm_projector_type=ptype, use_guide=use_guide, hidden_size=64, max_num_frames=2)
    try:
        want = ref.build_vision_projector(mk())
    except Exception as exc:  # noqa: BLE001 — whatever the reference raises is the contract
        with pytest.raises(type(exc)):
            hicom_b200.build_vision_projector(mk())
        return
    got = hicom_b200.build_vision_projector(mk())
    assert {k: tuple(v.shape) for k, v in got.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in want.state_dict().items()}
    for name in ("local_compressor", "global_compressor"):
        a, b = getattr(got, name), getattr(want, name)
        assert (a is None) == (b is None)
        if a is not None:
            assert a.use_guide == b.use_guide
    if got.local_compressor is not None:
        assert (got.local_compressor.temporal_kernel_size, got.local_compressor.spatial_kernel_size) == \
               (want.local_compressor.temporal_kernel_size, want.local_compressor.spatial_kernel_size)
    if got.global_compressor is not None:
        assert got.global_compressor.query.shape == want.global_compressor.query.shape


@pytest.mark.parametrize("name", ["none_T8", "direct_T8", "coarse_T8", "fine_T8", "adaptkv_coarse_T8", "adaptqkvg_coarse_T4"])
def test_oracle_autograd_equals_reference_autograd(name):
    """The training-path tests compare the product's gradients with PyTorch autograd through the ORACLE; this pins that
    yardstick to autograd through the reference module itself (what train.py:704-738 actually optimises)."""
    from oracle.cases import CASES_BY_NAME
    ref = load_reference()
    case = CASES_BY_NAME[name]
    sd, X, E, g, nl = materialise(case)
    m = ref.build_vision_projector(cfg_for(case))
    m.load_state_dict(sd, strict=True)
    m.train()
    out = m(X, E, g, case.modal, nl)
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    (out * probe).sum().backward()
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    got = O.OracleProjector(case.ptype, case.use_guide, case.merge, case.nlpos, leaf).forward(X, E, g, case.modal, nl)
    (got * probe).sum().backward()
    for k, p in m.named_parameters():
        a, b = leaf[k].grad, p.grad
        if b is None or float(b.abs().max()) <= 1e-6:
            assert a is None or float(a.abs().max()) <= 1e-6, k
        else:
            assert O.rel_err(a, b) <= 1e-5, (k, O.rel_err(a, b))
