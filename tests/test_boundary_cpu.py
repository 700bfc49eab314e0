"""CPU: the drop-in boundary — factory grammar, state_dict contract, C-ABI exports, loud failures."""
import ctypes
import os
import re

import pytest
import torch

from oracle import hicom_oracle as O

from util import Cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ptype", ["local43_global32", "local43_global32_coarse", "local412_global8",
                                   "local43_adaptqkvg_global32_adaptg", "local43guidefine_global32guidedirect",
                                   "global16", "local22"])
@pytest.mark.parametrize("use_guide", [None, "off", "direct", "coarse", "fine"])
def test_state_dict_contract(ptype, use_guide):
    import hicom_b200
    m = hicom_b200.build_vision_projector(Cfg(mm_projector_type=ptype, use_guide=use_guide, hidden_size=64))
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == O.param_shapes(ptype, use_guide, 64)
    assert "global_compressor.pos_embed" not in got  # non-persistent in the reference (projector.py:607)
    sd = O.synth_state_dict(ptype, use_guide, 64, seed=3)
    m.load_state_dict(sd, strict=True)
    m.to(torch.bfloat16)
    assert all(v.dtype == torch.bfloat16 for v in m.state_dict().values())


def test_factory_grammar():
    import hicom_b200
    from hicom_b200 import projector as P
    m = hicom_b200.build_vision_projector(Cfg(mm_projector_type="local43_global32_coarse", use_guide="direct"))
    assert isinstance(m, P.HIComProjector)
    # the _coarse suffix is ignored: guide mode comes from config.use_guide (SURVEY finding 1)
    assert m.local_compressor.use_guide == "direct" and m.global_compressor.use_guide == "direct"
    assert (m.local_compressor.temporal_kernel_size, m.local_compressor.spatial_kernel_size) == (4, 3)
    assert m.global_compressor.query.shape == (32, 1152)
    assert float(m.global_compressor.query.detach().abs().max()) == 0.0  # zero-init (projector.py:583)
    m = hicom_b200.build_vision_projector(Cfg(mm_projector_type="local412_global8"))
    assert m.local_compressor.spatial_kernel_size == 12 and m.global_compressor.query.shape[0] == 8
    m = hicom_b200.build_vision_projector(Cfg(mm_projector_type="local43guidecoarse_global32guidedirect"))
    assert m.local_compressor.use_guide == "coarse" and m.global_compressor.use_guide == "direct"
    assert isinstance(hicom_b200.build_vision_projector(Cfg(mm_projector_type="linear")), torch.nn.Linear)
    seq = hicom_b200.build_vision_projector(Cfg(mm_projector_type="mlp2x_gelu"))
    assert isinstance(seq, torch.nn.Sequential) and len(seq) == 3
    with pytest.raises(NotImplementedError):
        hicom_b200.build_vision_projector(Cfg(mm_vision_tower="openai/other"))
    with pytest.raises(AssertionError):
        P.HIComProjector(Cfg(), None, None)
    clip = hicom_b200.build_vision_projector(Cfg(mm_vision_tower="openai/clip-vit-large-patch14-336",
                                                 mm_hidden_size=768, hidden_size=64))
    assert clip.local_compressor.qk_dim == 768 and clip.global_compressor.attn_layer.num_heads == 6


def test_parser_matches_oracle_parser():
    import hicom_b200
    for ptype in ["local43_global32", "local412_global8_adaptg", "local43_adaptkv_global32", "local22guideoff_global4"]:
        spec = O.parse_projector_type(ptype)
        m = hicom_b200.build_vision_projector(Cfg(mm_projector_type=ptype, use_guide="coarse", hidden_size=32))
        lc, gc = m.local_compressor, m.global_compressor
        assert (lc.temporal_kernel_size, lc.spatial_kernel_size) == (spec.local.temporal_kernel, spec.local.spatial_kernel)
        assert isinstance(lc.k_alpha, torch.Tensor) == spec.local.adapt_k
        assert isinstance(lc.v_alpha, torch.Tensor) == spec.local.adapt_v
        assert gc.query.shape[0] == spec.global_.num_queries


def test_position_table_matches_reference_formula():
    import hicom_b200
    from hicom_b200.projector import _axis_table
    dense = hicom_b200.get_3d_position_embedding(5, 4, 3, 1152)
    assert torch.equal(torch.from_numpy(dense).float(), O.pos_embed_3d(5, 4, 3, 1152))
    sep = (torch.from_numpy(_axis_table(5, 1152)).float()[:, None, None] +
           torch.from_numpy(_axis_table(4, 1152)).float()[None, :, None] +
           torch.from_numpy(_axis_table(3, 1152)).float()[None, None, :])
    assert float((sep - O.pos_embed_3d(5, 4, 3, 1152)).abs().max()) < 5e-7


def test_cabi_exports_every_declared_symbol(built_library):
    header = open(os.path.join(ROOT, "include", "hicom_b200.h")).read()
    declared = set(re.findall(r"\b(hicom_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    lib = ctypes.CDLL(built_library)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hicom_b200.h but not exported"
    from hicom_b200 import _cabi
    assert set(_cabi.PROTOTYPES) == declared
    assert _cabi.load().hicom_abi_version() == 1


def test_no_cpu_fallback(built_library):
    import hicom_b200
    m = hicom_b200.build_vision_projector(Cfg(hidden_size=64))
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA tensors only"):
        m(torch.randn(4, 6, 6, 1152), None, None, "video")
    with pytest.raises(RuntimeError, match="CUDA tensors only"):   # training path (autograd on): same refusal
        m(torch.randn(4, 6, 6, 1152), None, None, "video")
    from hicom_b200 import autograd as ag
    ag.enable(False)
    try:
        with pytest.raises(RuntimeError, match="forward-only"):
            m(torch.randn(4, 6, 6, 1152), None, None, "video")
    finally:
        ag.enable(True)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hicom_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"


def test_default_splits_follow_cta_waves():
    """Split-K of the pooling GEMM: tiles = 9 x B x splits walk 148 persistent CTAs in waves; fewest splits within 5 %
    of the best wave count, every range >= 512 tokens."""
    from hicom_b200.projector import default_splits
    assert default_splits(32, 11664) == 1       # 288 tiles = two full waves, no partials to merge
    assert default_splits(16, 11664) == 1       # 144 tiles = one wave
    assert default_splits(1, 373248) == 16      # 144 tiles, not 297 (three waves)
    assert default_splits(1, 729) == 1          # a single frame cannot be cut below 512 tokens
    for B, N in [(1, 11664), (3, 5832), (7, 46656), (24, 11664), (100, 23328)]:
        s = default_splits(B, N)
        assert 1 <= s <= max(1, N // 512)
        cost = lambda k: -(-(9 * B * k) // 148) / k
        assert cost(s) <= 1.05 * min(cost(k) for k in range(1, max(1, min(64, N // 512)) + 1))


def test_splice_rows_per_sample_offsets():
    """Token splice with one offset per sample (hicom_arch.py:283-373: visual tokens follow prompts of different lengths)."""
    from hicom_b200.projector import splice_rows
    g = torch.Generator().manual_seed(0)
    out = torch.zeros(3, 12, 8)
    tokens = torch.randn(3, 4, 8, generator=g)
    ret = splice_rows(out, tokens, [0, 8, 3])
    assert ret is out
    for b, o in enumerate([0, 8, 3]):
        assert torch.equal(out[b, o:o + 4], tokens[b])
        mask = torch.ones(12, dtype=torch.bool); mask[o:o + 4] = False
        assert float(out[b, mask].abs().max()) == 0.0
    out2 = torch.zeros(3, 12, 8, dtype=torch.bfloat16)
    splice_rows(out2, tokens, torch.tensor([1, 2, 3]))
    assert torch.equal(out2[1, 2:6], tokens[1].bfloat16())
    with pytest.raises(ValueError, match="out of range"):
        splice_rows(out, tokens, [0, 9, 3])
    with pytest.raises(ValueError, match="do not match"):
        splice_rows(out, tokens, [0, 1])


def test_batch_view_is_zero_copy_for_split_views():
    """caller.batch_view: consecutive `split` views of one tower output (hicom_arch.py:162-164) become one strided view;
    scattered tensors fall back to a stacked copy with the same values."""
    from hicom_b200.caller import batch_view
    big = torch.randn(5 * 4, 3, 3, 8)
    parts = big.split([4] * 5, 0)
    v = batch_view(parts)
    assert v.data_ptr() == big.data_ptr() and v.shape == (5, 4, 3, 3, 8) and torch.equal(v, torch.stack(parts))
    v2 = batch_view([parts[1], parts[3]])  # constant pitch: still a view
    assert v2.data_ptr() == parts[1].data_ptr() and torch.equal(v2, torch.stack([parts[1], parts[3]]))
    for scattered in ([parts[3], parts[1]], [parts[0], parts[1], parts[3]], [parts[0], parts[1].clone()]):
        v3 = batch_view(scattered)
        assert torch.equal(v3, torch.stack(scattered)) and v3.data_ptr() != scattered[0].data_ptr()
    g = torch.randn(5, 8)
    assert batch_view([g[i] for i in range(5)]).data_ptr() == g.data_ptr()
    assert batch_view([g[2]]).shape == (1, 8)


def test_zero3_parameters_are_registered_as_external(monkeypatch):
    """DeepSpeed ZeRO-3 (hicom_trainer.py:21-38): partitioned parameters (``ds_id``) read outside their owner's forward
    must be registered with ``deepspeed.zero.register_external_parameter``.  DeepSpeed is not in this image: the call
    pattern is checked against a stub of its API."""
    import sys
    import types
    import hicom_b200
    from util import Cfg
    seen = []
    zero = types.ModuleType("deepspeed.zero")
    zero.register_external_parameter = lambda module, p: seen.append((module, p))
    ds = types.ModuleType("deepspeed")
    ds.zero = zero
    monkeypatch.setitem(sys.modules, "deepspeed", ds)
    monkeypatch.setitem(sys.modules, "deepspeed.zero", zero)
    m = hicom_b200.build_vision_projector(Cfg(use_guide="coarse"))
    assert m.register_zero3_parameters() == 0 and seen == []          # not partitioned: nothing to do
    for i, p in enumerate(m.parameters()):
        p.ds_id = i                                                    # what zero.Init / deepspeed.initialize adds
    X = torch.zeros(1, 4, 6, 6, 1152)
    with pytest.raises(RuntimeError, match="register_zero3_parameters"):
        m.forward_batched(X, X, torch.zeros(1, 1152), "video")
    n = m.register_zero3_parameters()
    assert n == len(list(m.parameters())) == len(seen) and all(mod is m for mod, _ in seen)
    assert {id(p) for _, p in seen} == {id(p) for p in m.parameters()}
    assert m.register_zero3_parameters() == n and len(seen) == n       # idempotent
