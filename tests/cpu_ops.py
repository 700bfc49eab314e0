"""TEST INFRASTRUCTURE: torch-CPU stand-ins for the ``hicom_b200.ops`` entry points used by the training path.

``tests/test_autograd_cpu.py`` patches them into ``hicom_b200.ops`` so that the backward FORMULAS and index plumbing of
``hicom_b200/autograd.py`` (which op is called with which strided view) can be checked on a machine without a GPU
against PyTorch autograd through the oracle.  The kernels themselves are checked on the GPU
(``tests/test_gpu_autograd.py``).  Each stand-in restates the contract written in ``include/hicom_b200.h``."""
import math

import torch
import torch.nn.functional as F

from oracle import hicom_oracle as O

ACT_NONE, ACT_GELU, ACT_GELU_TANH = 0, 1, 2
Q_POOLED, Q_FILM_LN, Q_VECTOR, Q_EXPLICIT = 0, 1, 2, 3
calls = []  # (name, ...) log so tests can assert which kernels the path would have launched


def _act(x, act):
    return F.gelu(x) if act == ACT_GELU else (F.gelu(x, approximate="tanh") if act == ACT_GELU_TANH else x)


def _same_dtype(*ts):
    """The kernels take ONE storage dtype per call (ops._check_linear & co. raise otherwise): the stand-ins insist too."""
    dts = {t.dtype for t in ts if t is not None}
    assert len(dts) <= 1, f"mixed dtypes reach a kernel: {dts}"


def linear(A, W, bias, residual, act, out_fp32, impl):
    calls.append(("linear", tuple(A.shape), tuple(W.shape), act))
    _same_dtype(A, W, bias, residual)
    y = _act(F.linear(A.float(), W.float(), None if bias is None else bias.float()), act)
    if residual is not None:
        y = y + residual.float().reshape(y.shape)
    return y if out_fp32 else y.to(A.dtype)


def gemm(A, B, out, out_fp32, alpha):
    calls.append(("gemm", tuple(A.shape), tuple(B.shape)))
    C = alpha * torch.matmul(A.float(), B.float())
    if out is None:
        return C if (out_fp32 or A.dtype == torch.float32) else C.to(A.dtype)
    assert out.stride(-1) == 1 or out.shape[-1] == 1
    out.copy_(C.reshape(out.shape))
    return out


def colsum(x2):
    calls.append(("colsum",))
    return x2.float().sum(0)


def act_backward(pre, dy, act):
    calls.append(("act_backward", act))
    x = pre.detach().float().requires_grad_(True)
    with torch.enable_grad():
        (g,) = torch.autograd.grad(_act(x, act), x, dy.float())
    return g.to(dy.dtype)


def softmax_backward(S, dP, lse, delta, out_bf16):
    calls.append(("softmax_backward",))
    dS = torch.exp(S - lse[:, None, :]) * (dP - delta[:, None, :])
    return dS.bfloat16() if out_bf16 else dS


def global_fold_query(q, Wk, heads, alpha):
    _same_dtype(q, Wk)
    B, Q, d = q.shape
    hd = d // heads
    qf = torch.einsum("bihc,hck->bhik", q.float().view(B, Q, heads, hd), Wk.float().view(heads, hd, d)) * alpha
    return qf.reshape(B, heads * Q, d).to(q.dtype)


def posadd(X, pt, ph, pw):
    return (X.float() + pt[None, :, None, None, :] + ph[None, None, :, None, :] + pw[None, None, None, :, :]).to(X.dtype)


def global_attend_partial(X, pt, ph, pw, qfold, splits, impl):
    calls.append(("global_attend_partial", splits))
    B, T, H, W, d = X.shape
    N = T * H * W
    Xp = posadd(X, pt, ph, pw).float().view(B, N, d)
    S = torch.matmul(Xp, qfold.float().transpose(1, 2))  # (B,N,J)
    chunk = -(-N // splits)
    ms, ls, os_ = [], [], []
    for s in range(splits):
        a, b = s * chunk, min(N, (s + 1) * chunk)
        m = S[:, a:b].max(dim=1).values
        p = torch.exp(S[:, a:b] - m[:, None, :])
        ms.append(m); ls.append(p.sum(1)); os_.append(torch.einsum("bnj,bnd->bjd", p, Xp[:, a:b]))
    return torch.stack(ms, 1), torch.stack(ls, 1), torch.stack(os_, 1)


def softmax_reduce(m, l, o):
    M = m.max(dim=1, keepdim=True).values
    w = torch.exp(m - M)
    return M, (l * w).sum(1, keepdim=True), (o * w[..., None]).sum(1, keepdim=True)


def softmax_merge(m, l, o, out_bf16):
    M, L, Osum = softmax_reduce(m, l, o)
    pooled = (Osum / L[..., None])[:, 0]
    return pooled.bfloat16() if out_bf16 else pooled


def softmax_merge_lse(m, l, o, out_bf16, lse_out=None):
    M, L, Osum = softmax_reduce(m, l, o)
    pooled = (Osum / L[..., None])[:, 0]
    lse = (M + torch.log(L))[:, 0]
    if lse_out is not None:
        lse_out.copy_(lse)
    return (pooled.bfloat16() if out_bf16 else pooled), lse


def out_code(dtype):
    return 1 if dtype == torch.bfloat16 else (2 if dtype == torch.float16 else 0)


def global_attend_partial_keys(X, Kscore, pt, ph, pw, qfold, splits, impl):
    calls.append(("global_attend_partial_keys", splits))
    B, T, H, W, d = X.shape
    N = T * H * W
    Xp = posadd(X, pt, ph, pw).float().view(B, N, d)
    S = torch.matmul(Kscore.float().view(B, N, d), qfold.float().transpose(1, 2))
    chunk = -(-N // splits)
    ms, ls, os_ = [], [], []
    for s in range(splits):
        a, b = s * chunk, min(N, (s + 1) * chunk)
        m = S[:, a:b].max(dim=1).values
        p = torch.exp(S[:, a:b] - m[:, None, :])
        ms.append(m); ls.append(p.sum(1)); os_.append(torch.einsum("bnj,bnd->bjd", p, Xp[:, a:b]))
    return torch.stack(ms, 1), torch.stack(ls, 1), torch.stack(os_, 1)


def col_logsumexp(S):
    return torch.logsumexp(S, dim=1)


def l2norm_rows(X):
    return (X.float() / X.float().norm(p=2, dim=-1, keepdim=True)).to(X.dtype)


def l2norm_rows_backward(X, dY):
    x, g = X.float(), dY.float()
    n = x.norm(p=2, dim=-1, keepdim=True)
    y = x / n
    return ((g - y * (y * g).sum(-1, keepdim=True)) / n).to(X.dtype)


def grid_pool_backward(dQ, T, H, W, kt, ks):
    B, nw, d = dQ.shape
    ds = (math.ceil(T / kt), math.ceil(H / ks), math.ceil(W / ks))
    x = torch.zeros((B, d, T, H, W), dtype=torch.float32, requires_grad=True)
    with torch.enable_grad():
        q = F.interpolate(x, size=ds, mode="trilinear")
        (gx,) = torch.autograd.grad(q, x, dQ.float().view(B, *ds, d).permute(0, 4, 1, 2, 3))
    return gx.permute(0, 2, 3, 4, 1).contiguous()


def global_value_proj(pooled, Wv, bv, Q, heads):
    _same_dtype(pooled, Wv, bv)
    B, J, d = pooled.shape
    hd = d // heads
    a = torch.einsum("bhik,hck->bihc", pooled.float().view(B, heads, Q, d), Wv.float().view(heads, hd, d))
    a = a.reshape(B, Q, d)
    if bv is not None:
        a = a + bv.float()
    return a.to(pooled.dtype)


def grid_pool(X, kt, ks):
    B, T, H, W, d = X.shape
    ds = (math.ceil(T / kt), math.ceil(H / ks), math.ceil(W / ks))
    q = F.interpolate(X.float().permute(0, 4, 1, 2, 3), size=ds, mode="trilinear")
    return q.permute(0, 2, 3, 4, 1).reshape(B, -1, d).to(X.dtype)


def film_layernorm(x, film, ln_w, ln_b, rows_per_group):
    _same_dtype(x, ln_w, ln_b)
    assert film.dtype == torch.float32
    d = x.shape[-1]
    rows = x.reshape(-1, d).float()
    g = torch.arange(rows.shape[0]) // rows_per_group
    u = rows * (1 + film[g, :d]) + film[g, d:]
    return F.layer_norm(u, (d,), ln_w.float(), ln_b.float(), 1e-6).to(x.dtype).reshape(x.shape)


def film_layernorm_backward(x, film, ln_w, dy, rows_per_group, need_dx):
    calls.append(("film_layernorm_backward",))
    xx = x.detach().float().requires_grad_(True)
    ff = film.detach().clone().requires_grad_(True)
    ww = ln_w.detach().float().requires_grad_(True)
    bb = torch.zeros_like(ww).requires_grad_(True)
    with torch.enable_grad():
        y = film_layernorm(xx, ff, ww, bb, rows_per_group)
        dx, dfilm, dw, db = torch.autograd.grad(y, (xx, ff, ww, bb), dy.float())
    return (dx.to(x.dtype) if need_dx else x.new_empty(0)), dfilm, dw, db


def local_attend(K, V, P, q_aux, film, ln_w, ln_b, kt, ks, qmode, logit_scale, k_l2norm):
    calls.append(("local_attend", qmode))
    B, T, H, W, d = V.shape
    if qmode == Q_POOLED:
        Q = grid_pool(P, kt, ks).float()
    elif qmode == Q_VECTOR:
        nw = math.ceil(T / kt) * math.ceil(H / ks) * math.ceil(W / ks)
        Q = q_aux.float()[:, None, :].expand(B, nw, d)
    elif qmode == Q_EXPLICIT:
        Q = q_aux.float()
    else:
        q0 = grid_pool(P, kt, ks)
        Q = film_layernorm(q0, film, ln_w, ln_b, q0.shape[1]).float()
    outs = []
    for b in range(B):
        k = K[b].float()
        if k_l2norm:
            k = k / k.norm(dim=-1, keepdim=True)
        rk = O.window_gather(k, (kt, ks, ks))
        rv = O.window_gather(V[b].float(), (kt, ks, ks))
        s = torch.einsum("nd,nmd->nm", Q[b], rk) * logit_scale
        outs.append(torch.einsum("nm,nmd->nd", torch.softmax(s, dim=-1), rv))
    return torch.stack(outs).to(V.dtype)


def local_attend_backward(K, V, Q, dO, kt, ks, logit_scale, k_l2norm, need_q, need_k, need_v):
    calls.append(("local_attend_backward", need_q, need_k, need_v))
    q = Q.detach().float().requires_grad_(True)
    k = K.detach().float().requires_grad_(True)
    v = V.detach().float().requires_grad_(True)
    with torch.enable_grad():
        o = local_attend(k, v, v, q, None, None, None, kt, ks, Q_EXPLICIT, logit_scale, k_l2norm)
        dq, dk, dv = torch.autograd.grad(o, (q, k, v), dO.float())
    return (dq.to(Q.dtype) if need_q else None), (dk if need_k else None), (dv if need_v else None)


def mix_layernorm(x, y, ln_w, ln_b, alpha):
    _same_dtype(x, y, ln_w, ln_b, alpha)
    a = alpha.float()
    ln = F.layer_norm(y.float(), (y.shape[-1],), ln_w.float(), ln_b.float(), 1e-6)
    return ((1 - a) * x.float() + a * ln).to(x.dtype)


def mix_layernorm_backward(x, y, ln_w, ln_b, alpha, dout, need_dx):
    calls.append(("mix_layernorm_backward",))
    leaves = [t.detach().float().requires_grad_(True) for t in (x, y, ln_w, ln_b, alpha)]
    with torch.enable_grad():
        out = mix_layernorm(*leaves)
        dx, dy, dw, db, da = torch.autograd.grad(out, leaves, dout.float())
    return (dx.to(x.dtype) if need_dx else None), dy.to(y.dtype), dw, db, da.reshape(1)


def add_layernorm(a, b, ln_w, ln_b):
    _same_dtype(a, b, ln_w, ln_b)
    return F.layer_norm(a.float() + b.float(), (a.shape[-1],), ln_w.float(), ln_b.float(), 1e-6).to(a.dtype)


def guide_attend(q, k, v, heads, scale):
    _same_dtype(q, k, v)
    B, n, d = q.shape
    hd = d // heads
    sp = lambda t: t.float().view(B, t.shape[1], heads, hd).transpose(1, 2)
    p = torch.softmax(torch.matmul(sp(q), sp(k).transpose(2, 3)) * scale, dim=-1)
    return torch.matmul(p, sp(v)).transpose(1, 2).reshape(B, n, d).to(q.dtype)


def layernorm(x, w, b):
    _same_dtype(x, w, b)
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), 1e-6).to(x.dtype)


def _need_cuda(*ts):
    for t in ts:
        if t is not None:
            return t.device
    return None


ALL = ["linear", "gemm", "colsum", "act_backward", "softmax_backward", "global_fold_query", "posadd", "global_attend_partial",
       "softmax_reduce", "softmax_merge", "global_value_proj", "grid_pool", "film_layernorm", "film_layernorm_backward",
       "local_attend", "local_attend_backward", "layernorm", "mix_layernorm", "mix_layernorm_backward", "add_layernorm",
       "guide_attend", "_need_cuda", "softmax_merge_lse", "out_code", "global_attend_partial_keys", "l2norm_rows",
       "l2norm_rows_backward", "grid_pool_backward", "col_logsumexp"]


def install(monkeypatch):
    """Patch every stand-in into hicom_b200.ops (the modules look ops up by attribute at call time)."""
    import sys
    from hicom_b200 import ops
    me = sys.modules[__name__]
    for name in ALL:
        monkeypatch.setattr(ops, name, getattr(me, name))
    calls.clear()
