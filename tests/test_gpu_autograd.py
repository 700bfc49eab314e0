"""GPU: the backward C-ABI blocks (csrc/backward.cu) against their torch stand-ins (tests/cpu_ops.py), and the whole
training path (hicom_b200/autograd.py) against PyTorch autograd through the oracle — what the reference's training
step computes for mm_projector (train.py:704-738).  Tolerances: fp32 gradients max|a-b|/max|b| <= 1e-3 (SIMT fp32
accumulation over up to 12k-token reductions), bf16 gradients cosine >= 0.99 against the fp32 truth."""
import dataclasses

import pytest
import torch

import cpu_ops
from oracle import hicom_oracle as O
from oracle.cases import CASES_BY_NAME, materialise
from util import cfg_for

pytestmark = pytest.mark.gpu


def _r(*shape, seed=0, std=1.0, dtype=torch.float32):
    return (std * torch.randn(*shape, generator=torch.Generator().manual_seed(seed))).to(dtype)


@pytest.mark.parametrize("da,db,mode", [(torch.float32, torch.float32, "same"), (torch.float32, torch.float32, "bf16view"),
                                        (torch.bfloat16, torch.bfloat16, "same"), (torch.bfloat16, torch.bfloat16, "f32"),
                                        (torch.float32, torch.bfloat16, "f32")])
def test_gemm_strided_views(da, db, mode, built_library):
    from hicom_b200 import ops
    A = _r(3, 5, 70, 33, seed=1, dtype=da).cuda()
    B = _r(5, 90, 33, seed=2, dtype=db).cuda()                     # used transposed, broadcast over the first batch dim
    want = 0.5 * torch.matmul(A.float().cpu(), B.float().cpu().transpose(1, 2))
    if mode == "bf16view":                                         # strided destination: every second row of a buffer
        buf = torch.zeros(3, 5, 140, 90, dtype=torch.bfloat16, device="cuda")
        ops.gemm(A, B.transpose(1, 2), buf[:, :, ::2], False, 0.5)
        assert float(buf[:, :, 1::2].abs().max()) == 0.0
        got, tol = buf[:, :, ::2], 6e-3
    else:
        got = ops.gemm(A, B.transpose(1, 2), None, mode == "f32", 0.5)
        assert got.dtype == (torch.float32 if mode == "f32" or da == torch.float32 else torch.bfloat16)
        tol = 1e-5 if (da, db) == (torch.float32, torch.float32) else 6e-3
    assert got.shape == want.shape and O.rel_err(got.float().cpu(), want) <= tol
    # A stored [k, m] (dW = dYᵀ·A), a one-row column sum, and a head-permuted destination view
    At, Bm = _r(333, 70, seed=3, dtype=da).cuda(), _r(333, 20, seed=4, dtype=db).cuda()
    got = ops.gemm(At.t(), Bm, None, True, 1.0)
    assert O.rel_err(got.cpu(), At.float().cpu().t() @ Bm.float().cpu()) <= (1e-5 if tol == 1e-5 else 6e-3)
    if da == db:
        ones = torch.ones(1, 333, dtype=da, device="cuda")
        assert O.rel_err(ops.gemm(ones, Bm, None, True, 1.0).cpu()[0], Bm.float().cpu().sum(0)) <= tol
        dq = torch.zeros(2, 6, 4, 16, dtype=da, device="cuda")     # (B, Q, heads, hd) written as [b, h, i, c]
        Ah, Bh = _r(2, 4, 6, 64, seed=5, dtype=da).cuda(), _r(4, 64, 16, seed=6, dtype=db).cuda()
        ops.gemm(Ah, Bh, dq.permute(0, 2, 1, 3), False, 1.0)
        want_h = torch.matmul(Ah.float().cpu(), Bh.float().cpu()).permute(0, 2, 1, 3)
        assert O.rel_err(dq.float().cpu(), want_h) <= tol


@pytest.mark.parametrize("layout", ["NT", "NN", "TN"])
@pytest.mark.parametrize("M,N,K,batch", [(2916, 288, 1152, 2), (700, 1152, 3584, 1), (1152, 4304, 648, 1), (288, 1152, 2916, 3)])
def test_gemm_large_bf16_layouts(layout, M, N, K, batch, built_library):
    """The three operand layouts of the backward's big bf16 contractions (tcgen05 path for large problems, SIMT
    otherwise — same contract): NT S = x'·qfoldᵀ, NN dA = dY·W, TN dW = dYᵀ·A with fp32 output; ragged M/N/K tiles."""
    from hicom_b200 import ops
    A = _r(batch, M, K, seed=1, std=0.5, dtype=torch.bfloat16)
    Bm = _r(batch, K, N, seed=2, std=0.05, dtype=torch.bfloat16)
    want = torch.matmul(A.float(), Bm.float())
    if layout == "NT":
        a_op, b_op = A.cuda(), Bm.transpose(1, 2).contiguous().cuda().transpose(1, 2)     # B stored (N, K)
    elif layout == "NN":
        a_op, b_op = A.cuda(), Bm.cuda()
    else:
        a_op, b_op = A.transpose(1, 2).contiguous().cuda().transpose(1, 2), Bm.cuda()     # A stored (K, M)
    got32 = ops.gemm(a_op, b_op, None, True, 1.0)
    assert got32.dtype == torch.float32 and O.rel_err(got32.cpu(), want) <= 2e-3
    if layout != "TN":
        got16 = ops.gemm(a_op, b_op, None, False, 0.5)
        assert got16.dtype == torch.bfloat16 and O.rel_err(got16.float().cpu(), 0.5 * want) <= 6e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", [1, 2])
def test_act_backward(act, dtype, built_library):
    from hicom_b200 import ops
    pre, dy = _r(37, 130, seed=1, std=2.0), _r(37, 130, seed=2).to(dtype)
    got = ops.act_backward(pre.cuda(), dy.cuda(), act).float().cpu()
    want = cpu_ops.act_backward(pre, dy.float(), act)
    assert O.rel_err(got, want) <= (2e-6 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,pitch", [(2592, 3584, 3584), (5, 8, 8), (1000, 1152, 2304), (70000, 128, 128)])
def test_colsum(M, N, pitch, dtype, built_library):
    from hicom_b200 import ops
    buf = _r(M, pitch, seed=M + N, std=0.5).to(dtype)
    got = ops.colsum(buf.cuda()[:, :N]).cpu()
    want = buf[:, :N].double().sum(0).float()
    assert got.dtype == torch.float32 and O.rel_err(got, want) <= 1e-5
    assert not ops.colsum_supported(buf.cuda()[:, 1:N - 1])


def test_forward_batched_splices_at_per_sample_offsets(built_library):
    """Token splice with one row offset per sample (hicom_arch.py:283-373) == the one-block result, copied once."""
    case = CASES_BY_NAME["coarse_T4"]
    sd, X, E, g, _ = materialise(case)
    m = _train_module(case, sd).eval()
    Xb, Eb, gb = torch.stack([X, X.flip(0)]).cuda(), torch.stack([E, E.flip(0)]).cuda(), torch.stack([g, -g]).cuda()
    with torch.no_grad():
        block = m.forward_batched(Xb, Eb, gb, "video")
        n = block.shape[1]
        buf = torch.zeros(2, n + 9, block.shape[2], device="cuda")
        ret = m.forward_batched(Xb, Eb, gb, "video", out=buf, out_row_offset=[5, 0])
    assert O.rel_err(ret.cpu(), block.cpu()) <= 1e-5          # two runs: fp32 atomics may reorder the last bit
    assert torch.equal(buf[0, 5:5 + n], ret[0]) and torch.equal(buf[1, :n], ret[1])
    assert float(buf[0, :5].abs().max()) == 0.0 and float(buf[1, n:].abs().max()) == 0.0


@pytest.mark.parametrize("out_bf16", [False, True])
def test_softmax_backward(out_bf16, built_library):
    from hicom_b200 import ops
    B, N, J = 2, 301, 9
    S, dP = _r(B, N, J, seed=1, std=3.0), _r(B, N, J, seed=2)
    lse = torch.logsumexp(S, dim=1)
    delta = _r(B, J, seed=3)
    got = ops.softmax_backward(S.cuda(), dP.cuda(), lse.cuda(), delta.cuda(), out_bf16).float().cpu()
    want = cpu_ops.softmax_backward(S, dP, lse, delta, False)
    assert O.rel_err(got, want) <= (2e-6 if not out_bf16 else 4e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(8, 6, 6, 4, 3, False), (7, 7, 8, 4, 3, False), (1, 6, 6, 1, 3, True), (2, 6, 6, 4, 2, False)])
def test_local_attend_backward(shape, dtype, built_library):
    """dQ, dK, dV of the window attention; (7,7,8) has overlapping balanced windows (tokens in two windows: atomics)."""
    from hicom_b200 import ops
    T, H, W, kt, ks, l2 = shape
    B, d = 2, 1152
    K, V = _r(B, T, H, W, d, seed=1, std=0.5).to(dtype), _r(B, T, H, W, d, seed=2, std=0.5).to(dtype)
    nw = ops.num_windows(T, H, W, kt, ks)
    Q, dO = _r(B, nw, d, seed=3, std=0.5).to(dtype), _r(B, nw, d, seed=4).to(dtype)
    scale = 1.0 / d ** 0.5 if not l2 else 10.0
    dq, dk, dv = ops.local_attend_backward(K.cuda(), V.cuda(), Q.cuda(), dO.cuda(), kt, ks, scale, l2, True, not l2, True)
    wq, wk, wv = cpu_ops.local_attend_backward(K.float(), V.float(), Q.float(), dO.float(), kt, ks, scale, l2,
                                               True, True, True)
    tol = 2e-5 if dtype == torch.float32 else 8e-3
    assert O.rel_err(dq.float().cpu(), wq) <= tol
    assert dv.dtype == torch.float32 and O.rel_err(dv.cpu(), wv) <= tol
    if l2:
        assert dk is None
        with pytest.raises(RuntimeError, match="L2 normalisation"):
            ops.local_attend_backward(K.cuda(), V.cuda(), Q.cuda(), dO.cuda(), kt, ks, scale, True, False, True, False)
    else:
        assert dk.dtype == torch.float32 and O.rel_err(dk.cpu(), wk) <= tol
    only_k = ops.local_attend_backward(K.cuda(), V.cuda(), Q.cuda(), dO.cuda(), kt, ks, scale, False, False, True, False)
    assert only_k[0] is None and only_k[2] is None and O.rel_err(
        only_k[1].cpu(), cpu_ops.local_attend_backward(K.float(), V.float(), Q.float(), dO.float(), kt, ks, scale,
                                                       False, False, True, False)[1]) <= tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,rpg,need_dx", [(64, 32, True), (324, 162, False), (2000, 500, True), (5, 1, True)])
def test_film_layernorm_backward(rows, rpg, need_dx, dtype, built_library):
    from hicom_b200 import ops
    d = 1152
    G = -(-rows // rpg)
    x, dy = _r(rows, d, seed=1, std=0.5).to(dtype), _r(rows, d, seed=2).to(dtype)
    film = _r(G, 2 * d, seed=3, std=0.3)
    w = (1 + _r(d, seed=4, std=0.1)).to(dtype)
    dx, dfilm, dw, db = ops.film_layernorm_backward(x.cuda(), film.cuda(), w.cuda(), dy.cuda(), rpg, need_dx)
    wdx, wfilm, wdw, wdb = cpu_ops.film_layernorm_backward(x.float(), film, w.float(), dy.float(), rpg, True)
    tol = 2e-5 if dtype == torch.float32 else 8e-3
    if need_dx:
        assert O.rel_err(dx.float().cpu(), wdx) <= tol
    else:
        assert dx.numel() == 0
    assert O.rel_err(dfilm.cpu(), wfilm) <= tol and O.rel_err(dw.cpu(), wdw) <= tol and O.rel_err(db.cpu(), wdb) <= tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,need_dx", [(64, True), (700, False), (5, True)])
def test_mix_layernorm_backward(rows, need_dx, dtype, built_library):
    from hicom_b200 import ops
    d = 1152
    x, y, dout = (_r(rows, d, seed=s, std=0.5).to(dtype) for s in (1, 2, 3))
    w, b = (1 + _r(d, seed=4, std=0.1)).to(dtype), _r(d, seed=5, std=0.1).to(dtype)
    alpha = torch.tensor([0.3]).to(dtype)
    dx, dy, dw, db, da = ops.mix_layernorm_backward(x.cuda(), y.cuda(), w.cuda(), b.cuda(), alpha.cuda(), dout.cuda(), need_dx)
    wdx, wdy, wdw, wdb, wda = cpu_ops.mix_layernorm_backward(x.float(), y.float(), w.float(), b.float(), alpha.float(),
                                                             dout.float(), True)
    tol = 2e-5 if dtype == torch.float32 else 8e-3
    assert (dx is None) == (not need_dx)
    if need_dx:
        assert O.rel_err(dx.float().cpu(), wdx) <= tol
    assert O.rel_err(dy.float().cpu(), wdy) <= tol and O.rel_err(dw.cpu(), wdw) <= tol and O.rel_err(db.cpu(), wdb) <= tol
    assert abs(float(da.cpu()) - float(wda)) <= tol * max(1.0, abs(float(wda)))


# ---- the whole training path -------------------------------------------------------------------------------------
def _oracle_grads(case, sd, X, E, g, nl, probe):
    leaf = {k: v.float().clone().requires_grad_(True) for k, v in sd.items()}
    f = lambda t: None if t is None else t.float()
    orc = O.OracleProjector(case.ptype, case.use_guide, case.merge, case.nlpos, leaf)
    out = orc.forward(f(X), f(E), f(g), case.modal, f(nl))
    (out * probe).sum().backward()
    return out.detach(), {k: v.grad for k, v in leaf.items()}


def _train_module(case, sd):
    import hicom_b200
    m = hicom_b200.build_vision_projector(cfg_for(case))
    m.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
    return m.to(getattr(torch, case.dtype)).cuda().train()


@pytest.fixture
def autograd_on():
    from hicom_b200 import autograd as ag
    old = ag.ENABLED
    ag.enable(True)
    yield ag
    ag.enable(old)


RELEASE_ADAPTKV = dataclasses.replace(CASES_BY_NAME["adaptkv_coarse_T8"], name="adaptkv_direct_T8", use_guide="direct")


@pytest.mark.parametrize("name", ["none_T8", "direct_T8", "coarse_T8", "coarse_T7", "coarse_nondiv_7x8",
                                  "global_only_coarse_T8", "video_grid_newline", "coarse_27x27_T4",
                                  "fine_T8", "adaptkv_coarse_T8", "adaptqkvg_fine_T8", "adaptkv_direct_T8"])
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_training_step_gradients(name, dtype, built_library, autograd_on):
    """`adaptkv_direct_T8` is the second release recipe (scripts/qwen2.5_7B/release/directg_local43_adaptkv_global32.sh)."""
    base = RELEASE_ADAPTKV if name == "adaptkv_direct_T8" else CASES_BY_NAME[name]
    case = dataclasses.replace(base, dtype=dtype)
    if dtype == "bfloat16" and name not in ("coarse_T8", "direct_T8", "coarse_27x27_T4", "adaptkv_direct_T8", "fine_T8"):
        pytest.skip("bf16 is covered on five representative cases")
    sd, X, E, g, nl = materialise(case)
    m = _train_module(case, sd)
    dev = lambda t: None if t is None else t.cuda()
    out = m(dev(X), dev(E), dev(g), case.modal, dev(nl))
    assert out.requires_grad
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    want_out, want = _oracle_grads(case, sd, X, E, g, nl, probe)
    fp32 = dtype == "float32"
    assert O.rel_err(out.detach().float().cpu(), want_out) <= (1e-4 if fp32 else 1e-2)
    (out.float() * probe.cuda()).sum().backward()
    for k, p in m.named_parameters():
        w = want[k]
        if w is None or float(w.abs().max()) <= 1e-6:   # unused parameter, or a key bias the softmax cancels (noise)
            assert p.grad is None or float(p.grad.float().abs().max()) <= (1e-6 if fp32 else 1e-3), k
            continue
        assert p.grad is not None and p.grad.dtype == p.dtype, k
        got = p.grad.float().cpu()
        if fp32:
            assert O.rel_err(got, w) <= 1e-3, (k, O.rel_err(got, w))
        else:
            assert O.cosine(got, w) >= 0.99, (k, O.cosine(got, w))


def test_optimizer_step_moves_every_parameter(built_library, autograd_on):
    """A real AdamW step through the drop-in module (what stage 1-3 training does to mm_projector)."""
    case = CASES_BY_NAME["coarse_T4"]
    sd, X, E, g, _ = materialise(case)
    m = _train_module(case, sd)
    before = {k: p.detach().clone() for k, p in m.named_parameters()}
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=0.0)
    out = m.forward_batched(torch.stack([X, X.flip(0)]).cuda(), torch.stack([E, E.flip(0)]).cuda(),
                            torch.stack([g, -g]).cuda(), "video")
    out.square().mean().backward()
    opt.step()
    moved = {k: float((p.detach() - before[k]).abs().max()) for k, p in m.named_parameters()}
    still = [k for k, v in moved.items() if v == 0.0 and not k.endswith("attn_layer.k_proj.bias")]
    assert not still, still


def test_switch_off_restores_forward_only(built_library):
    from hicom_b200 import autograd as ag
    case = CASES_BY_NAME["coarse_T4"]
    sd, X, E, g, _ = materialise(case)
    m = _train_module(case, sd)
    old = ag.ENABLED
    ag.enable(False)
    try:
        with pytest.raises(RuntimeError, match="forward-only"):
            m(X.cuda(), E.cuda(), g.cuda(), "video")
    finally:
        ag.enable(old)
    # what is not trainable stays loud with the path on: fp16 is the inference dtype (train in bf16 / fp32)
    with pytest.raises(NotImplementedError, match="dtype"):
        m.half()(X.cuda().half(), E.cuda().half(), g.cuda().half(), "video")


@pytest.mark.parametrize("name,dtype", [("direct_T8", "float32"), ("coarse_nondiv_7x8", "float32"),
                                        ("direct_T8", "bfloat16")])
def test_stage3_gradients_reach_frames_embed_and_guide(name, dtype, built_library, autograd_on):
    """vision_model_head + guide_encoder tuned (train.py:717-726): frames_embed and the instruction embedding carry
    gradients, compared with PyTorch autograd through the oracle."""
    case = dataclasses.replace(CASES_BY_NAME[name], dtype=dtype)
    sd, X, E, g, nl = materialise(case)
    m = _train_module(case, sd)
    E1, g1 = E.cuda().requires_grad_(True), g.cuda().requires_grad_(True)
    out = m(X.cuda(), E1, g1, case.modal, None)
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    (out.float() * probe.cuda()).sum().backward()
    leaf = {k: v.float().clone().requires_grad_(True) for k, v in sd.items()}
    E2, g2 = E.float().clone().requires_grad_(True), g.float().clone().requires_grad_(True)
    want = O.OracleProjector(case.ptype, case.use_guide, case.merge, case.nlpos, leaf).forward(X.float(), E2, g2,
                                                                                               case.modal, None)
    (want * probe).sum().backward()
    assert E1.grad.dtype == E1.dtype and g1.grad.dtype == g1.dtype
    if dtype == "float32":
        assert O.rel_err(E1.grad.cpu(), E2.grad) <= 1e-3 and O.rel_err(g1.grad.cpu(), g2.grad) <= 1e-3
    else:
        assert O.cosine(E1.grad.float().cpu(), E2.grad) >= 0.99 and O.cosine(g1.grad.float().cpu(), g2.grad) >= 0.99


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_producer_head_gradients(dtype, built_library, autograd_on):
    from hicom_b200.producer import SiglipHeadEmbed
    from oracle import siglip_head as SH
    sd = SH.synth_head_state(3)
    m = SiglipHeadEmbed()
    m.load_state_dict(sd, strict=True)
    m = m.to(dtype).cuda().train()
    h = SH.synth_hidden(2, 36, seed=2).to(dtype)
    out = m(h.cuda())
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(6))
    (out.float() * probe.cuda()).sum().backward()
    leaf = {k: v.to(dtype).float().clone().requires_grad_(True) for k, v in sd.items()}
    (SH.image_embeds(leaf, h.float(), side=6) * probe).sum().backward()
    for k, p in m.named_parameters():
        got = p.grad.float().cpu()
        assert p.grad.dtype == dtype
        if dtype == torch.float32:
            assert O.rel_err(got, leaf[k].grad) <= 1e-3, k
        else:
            assert O.cosine(got, leaf[k].grad) >= 0.99, k


def test_gemm_tn_batched_single_launch(built_library):
    from hicom_b200 import ops
    A = _r(5, 2916, 288, seed=1, std=0.5, dtype=torch.bfloat16).cuda()      # stored (K, M) per batch entry
    Bm = _r(5, 2916, 1152, seed=2, std=0.05, dtype=torch.bfloat16).cuda()
    got = ops.gemm(A.transpose(1, 2), Bm, None, True, 1.0)
    want = torch.matmul(A.float().cpu().transpose(1, 2), Bm.float().cpu())
    assert O.rel_err(got.cpu(), want) <= 2e-3


def test_gemm_nt_batched_single_launch(built_library):
    """S[b] = x'[b]·qfold[b]ᵀ per video: every batch entry of a K-major x K-major contraction in ONE tcgen05 launch
    (the kernel's batch axis walks A, B and C), ragged M tiles included."""
    from hicom_b200 import ops
    A = _r(3, 1000, 1152, seed=3, std=0.5, dtype=torch.bfloat16).cuda()
    Bm = _r(3, 288, 1152, seed=4, std=0.05, dtype=torch.bfloat16).cuda()
    n0 = ops.kernel_launch_count()
    got = ops.gemm(A, Bm.transpose(1, 2), None, True, 1.0)
    assert ops.kernel_launch_count() - n0 == 1
    want = torch.matmul(A.float().cpu(), Bm.float().cpu().transpose(1, 2))
    assert got.shape == (3, 1000, 288) and O.rel_err(got.cpu(), want) <= 2e-3


@pytest.mark.parametrize("adt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(8, 2304, 2304), (1, 1152, 2304), (5, 1028, 300)])
def test_gemm_skinny_nn(M, N, K, adt, built_library):
    """dA = dpre·W on a handful of rows (the per-video FiLM MLPs): the streaming skinny kernel (fp32 atomics into a
    zeroed C) against fp32 torch."""
    from hicom_b200 import ops
    A = _r(M, K, seed=5, std=0.5, dtype=adt).cuda()
    W = _r(K, N, seed=6, std=0.05, dtype=torch.bfloat16).cuda()
    got = ops.gemm(A, W, None, True, 0.5)
    want = 0.5 * torch.matmul(A.float().cpu(), W.float().cpu())
    assert got.dtype == torch.float32 and O.rel_err(got.cpu(), want) <= 1e-4


@pytest.mark.parametrize("mode", ["coarse", "direct"])
def test_graphed_training_step_matches_eager(mode, built_library, autograd_on):
    """forward + backward captured in CUDA graphs (hicom_b200.graph.graphed_training_forward): same tokens and the same
    parameter gradients as the eager training path, over two replays with different inputs."""
    import copy
    from hicom_b200.graph import graphed_training_forward
    case = dataclasses.replace(CASES_BY_NAME[f"{mode}_T8"], dtype="bfloat16")
    sd, X, E, g, nl = materialise(case)
    m = _train_module(case, sd)
    ref = copy.deepcopy(m)
    Xc, Ec, gc = X.cuda(), E.cuda(), g.cuda()
    fn = graphed_training_forward(m, Xc.unsqueeze(0), Ec.unsqueeze(0), gc.unsqueeze(0), case.modal)
    for scale in (1.0, 0.7):
        xs, es, gs = (Xc * scale).unsqueeze(0), (Ec * scale).unsqueeze(0), (gc * scale).unsqueeze(0)
        m.zero_grad(set_to_none=True)
        out = fn(xs, es, gs)
        out.float().square().mean().backward()
        ref.zero_grad(set_to_none=True)
        want = ref.forward_batched(xs, es, gs, case.modal)
        want.float().square().mean().backward()
        assert O.rel_err(out.float().cpu(), want.float().cpu()) <= 1e-2
        for (k, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
            assert (p.grad is None) == (q.grad is None), k
            if p.grad is not None and float(q.grad.float().abs().max()) > 0:
                assert O.cosine(p.grad.float().cpu(), q.grad.float().cpu()) >= 0.999, k


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_grid_pool_and_l2norm_backward(dtype, built_library):
    import cpu_ops
    from hicom_b200 import ops
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for (B, T, H, W, d, kt, ks) in [(2, 8, 6, 6, 128, 4, 3), (1, 7, 7, 8, 256, 4, 3), (1, 1, 6, 6, 128, 1, 2)]:
        nw = ops.num_windows(T, H, W, kt, ks)
        dq = _r(B, nw, d, seed=11, dtype=dtype)
        got = ops.grid_pool_backward(dq.cuda(), T, H, W, kt, ks)
        want = cpu_ops.grid_pool_backward(dq, T, H, W, kt, ks)
        assert got.dtype == torch.float32 and O.rel_err(got.cpu(), want) <= (1e-5 if dtype == torch.float32 else 1e-5)
    x, dy = _r(37, 1152, seed=12, dtype=dtype), _r(37, 1152, seed=13, dtype=dtype)
    got = ops.l2norm_rows_backward(x.cuda(), dy.cuda())
    assert got.dtype == dtype and O.rel_err(got.float().cpu(), cpu_ops.l2norm_rows_backward(x.float(), dy.float())) <= tol


@pytest.mark.parametrize("name,dtype", [("coarse_T8", "float32"), ("none_T8", "float32"), ("direct_T8", "bfloat16"),
                                        ("coarse_nondiv_7x8", "float32")])
def test_gradients_reach_frames_feature(name, dtype, built_library, autograd_on):
    """'pure_vision_model' tuned (train.py:712-715): frames_feature carries a gradient, compared with PyTorch autograd
    through the oracle."""
    case = dataclasses.replace(CASES_BY_NAME[name], dtype=dtype)
    sd, X, E, g, nl = materialise(case)
    m = _train_module(case, sd)
    dv = lambda t: None if t is None else t.cuda()
    X1 = X.cuda().requires_grad_(True)
    out = m(X1, dv(E), dv(g), case.modal, None)
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    (out.float() * probe.cuda()).sum().backward()
    f = lambda t: None if t is None else t.float()
    leaf = {k: v.float().clone().requires_grad_(True) for k, v in sd.items()}
    X2 = X.float().clone().requires_grad_(True)
    want = O.OracleProjector(case.ptype, case.use_guide, case.merge, case.nlpos, leaf).forward(X2, f(E), f(g), case.modal, None)
    (want * probe).sum().backward()
    assert X1.grad is not None and X1.grad.dtype == X1.dtype
    if dtype == "float32":
        assert O.rel_err(X1.grad.cpu(), X2.grad) <= 1e-3, O.rel_err(X1.grad.cpu(), X2.grad)
    else:
        assert O.cosine(X1.grad.float().cpu(), X2.grad) >= 0.99


@pytest.mark.parametrize("name,where,dtype", [("coarse_T8", "local,global", "float32"), ("direct_T8", "local,global", "bfloat16"),
                                              ("adaptkv_coarse_T8", "local", "float32"), ("none_T8", "global", "float32")])
def test_clip_scale_training_step(name, where, dtype, built_library, autograd_on):
    """use_clip_scale with trainable logit_scale / logit_bias ('attn_scale', train.py:729-732): every gradient against
    PyTorch autograd through the oracle."""
    case = dataclasses.replace(CASES_BY_NAME[name], dtype=dtype)
    sd, X, E, g, nl = materialise(case)
    m = _train_module(case, sd)
    scal = {}
    for part in where.split(","):
        ls = torch.nn.Parameter(torch.tensor(1.3, device="cuda"))
        lb = torch.nn.Parameter(torch.tensor(-0.7, device="cuda"))
        setattr(m, f"{part}_logit_scale", ls)
        setattr(m, f"{part}_logit_bias", lb)
        scal[part] = (ls, lb)
    dv = lambda t: None if t is None else t.cuda()
    out = m(dv(X), dv(E), dv(g), case.modal, None)
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(9))
    (out.float() * probe.cuda()).sum().backward()
    f = lambda t: None if t is None else t.float()
    leaf = {k: v.float().clone().requires_grad_(True) for k, v in sd.items()}
    orc = O.OracleProjector(case.ptype, case.use_guide, case.merge, case.nlpos, leaf)
    oscal = {}
    for part in where.split(","):
        oscal[part] = (torch.tensor(1.3, requires_grad=True), torch.tensor(-0.7, requires_grad=True))
        setattr(orc, f"{part}_logit", oscal[part])
    want = orc.forward(f(X), f(E), f(g), case.modal, None)
    (want * probe).sum().backward()
    if dtype == "float32":
        assert O.rel_err(out.detach().float().cpu(), want.detach()) <= 1e-4
    for k, p in m.named_parameters():
        if k.endswith(("logit_scale", "logit_bias")):
            continue
        w = leaf[k].grad
        if w is None or float(w.abs().max()) <= 1e-6:
            continue
        assert p.grad is not None, k
        if dtype == "float32":
            assert O.rel_err(p.grad.cpu(), w) <= 2e-3, (k, O.rel_err(p.grad.cpu(), w))
        else:
            # the key bias of the normalised keys is a column sum over every token of gradients that nearly cancel (1e-5
            # against 1e-2 for the weights): bf16 rounding of the per-token gradients is visible in it
            assert O.cosine(p.grad.float().cpu(), w) >= (0.8 if k.endswith("attn_layer.k_proj.bias") else 0.99), k
    for part, (ls, lb) in scal.items():
        ols, _ = oscal[part]
        if ols.grad is not None and abs(float(ols.grad)) > 1e-6:
            tol = 2e-3 if dtype == "float32" else 5e-2
            assert ls.grad is not None and abs(float(ls.grad) - float(ols.grad)) <= tol * max(1.0, abs(float(ols.grad))), \
                (part, float(ls.grad), float(ols.grad))


def test_col_logsumexp(built_library):
    from hicom_b200 import ops
    S = _r(3, 5000, 288, seed=21, std=4.0)
    S[1, 77, 5] = 90.0                                  # one dominant score
    got = ops.col_logsumexp(S.cuda())
    assert got.shape == (3, 288) and O.rel_err(got.cpu(), torch.logsumexp(S, dim=1)) <= 1e-6
    assert torch.equal(S, S.clone())                    # input untouched (the kernel reads only)
