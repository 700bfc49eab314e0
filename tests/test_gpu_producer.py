"""GPU: producer side — frames_embed = h + head.mlp(head.layernorm(h)) (encoder.py:284-286) through the
C ABI (hicom_layernorm, hicom_linear with the tanh GELU / residual epilogues) against oracle/siglip_head.py.
Tolerances as for the compressor: fp32 max|a-b|/max|b| <= 1e-4; bf16 against the fp32 oracle on the bf16-rounded
weights and inputs: cosine >= 0.999 and <= 1e-2."""
import pytest
import torch
import torch.nn.functional as F

from oracle import hicom_oracle as O
from oracle import siglip_head as SH

pytestmark = pytest.mark.gpu


def _module(sd, dtype):
    from hicom_b200.producer import SiglipHeadEmbed
    m = SiglipHeadEmbed(sd["mlp.fc1.weight"].shape[1], sd["mlp.fc1.weight"].shape[0])
    m.load_state_dict(sd, strict=True)
    return m.to(dtype).cuda().eval()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm_op(dtype, built_library):
    from hicom_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = (0.7 * torch.randn(37, 1152, generator=g) + 0.3).to(dtype)
    w, b = (1 + 0.1 * torch.randn(1152, generator=g)).to(dtype), (0.1 * torch.randn(1152, generator=g)).to(dtype)
    got = ops.layernorm(x.cuda(), w.cuda(), b.cuda()).float().cpu()
    want = F.layer_norm(x.float(), (1152,), w.float(), b.float(), 1e-6)
    assert O.rel_err(got, want) <= (2e-6 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(5, 4304, 1152), (36, 4304, 1152), (36, 1152, 4304), (1458, 4304, 1152), (2916, 4304, 1152)])
def test_linear_gelu_tanh(M, N, K, dtype, built_library):
    """tanh GELU epilogue on every tile family (skinny, 128x64, 128x256, CTA pairs), ragged N (4304 = 16*256 + 208,
    last 32-column chunk half full) and ragged K (4304 = 67*64 + 16: the TMA box reads zeros past the end)."""
    from hicom_b200 import ops
    if dtype == torch.float32 and M > 1500:
        pytest.skip("fp32 runs on the SIMT GEMM for every M; one large case is enough")
    g = torch.Generator().manual_seed(M + N)
    A = (0.5 * torch.randn(M, K, generator=g)).to(dtype)
    Wt = (0.03 * torch.randn(N, K, generator=g)).to(dtype)
    b = (0.05 * torch.randn(N, generator=g)).to(dtype)
    want = F.gelu(F.linear(A.float(), Wt.float(), b.float()), approximate="tanh")
    for impl in (ops.IMPL_SIMT, ops.IMPL_AUTO):
        got = ops.linear(A.cuda(), Wt.cuda(), b.cuda(), None, ops.ACT_GELU_TANH, False, impl).float().cpu()
        assert O.rel_err(got, want) <= (2e-5 if dtype == torch.float32 else 6e-3), impl


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("b,side", [(1, 6), (2, 27), (4, 27)])
def test_image_embeds_parity(b, side, dtype, built_library):
    if dtype == torch.float32 and b == 4:
        pytest.skip("covered by b=2 (SIMT path, same kernels)")
    sd = {k: v.to(dtype) for k, v in SH.synth_head_state(0).items()}
    h = SH.synth_hidden(b, side * side, seed=b).to(dtype)
    m = _module(sd, dtype)
    with torch.no_grad():
        got = m(h.cuda())
    assert got.shape == (b, side, side, 1152) and got.dtype == dtype
    got = got.float().cpu()
    truth = SH.image_embeds({k: v.float() for k, v in sd.items()}, h.float(), side=side)
    err, cos = O.rel_err(got, truth), O.cosine(got, truth)
    if dtype == torch.float32:
        assert err <= 1e-4, err
    else:
        assert cos >= 0.999 and err <= 1e-2, (err, cos)


def test_text_head_embed(built_library):
    from hicom_b200.producer import text_head_embed
    g = torch.Generator().manual_seed(9)
    lin = torch.nn.Linear(1152, 1152)
    x = 0.5 * torch.randn(3, 32, 1152, generator=g)
    want = SH.text_embeds_fine(lin.weight.detach(), lin.bias.detach(), x)
    lin = lin.cuda().eval()
    with torch.no_grad():
        got = text_head_embed(x.cuda(), lin).cpu()
    assert got.shape == want.shape and O.rel_err(got, want) <= 1e-4


def test_feeds_the_compressor(built_library):
    """Producer output is consumed in place as frames_embed (same (T,27,27,1152) channel-last layout, no re-layout)."""
    import hicom_b200
    from util import Cfg
    dtype = torch.bfloat16
    sd = {k: v.to(dtype) for k, v in SH.synth_head_state(1).items()}
    m = _module(sd, dtype)
    proj = hicom_b200.build_vision_projector(Cfg(use_guide="coarse", hidden_size=896, max_num_frames=4))
    torch.manual_seed(0)
    proj = proj.to(dtype).cuda().eval()
    h = SH.synth_hidden(4, 729, seed=3).to(dtype).cuda()
    X = (0.5 * torch.randn(4, 27, 27, 1152, generator=torch.Generator().manual_seed(2))).to(dtype).cuda()
    gvec = (0.5 * torch.randn(1152, generator=torch.Generator().manual_seed(4))).to(dtype).cuda()
    with torch.no_grad():
        E = m(h)
        assert E.is_contiguous() and E.shape == X.shape
        out = proj(X, E, gvec, "video")
    assert out.shape == (81 + 32, 896) and bool(torch.isfinite(out.float()).all())
