"""Training path of the compressor: autograd for the custom ops.

The reference trains ``mm_projector`` in all three stages (train.py:704-738, ``mm_tunable_parts``) through PyTorch
autograd.  ``frames_feature`` always comes from the frozen SigLIP body (encoder.py:235) and never needs a gradient; stage
3 of the release recipe additionally tunes ``vision_model_head`` and ``guide_encoder`` (train.py:717-726), so
``frames_embed`` (the keys of the local attention, made by the head: ``hicom_b200.producer``) and the instruction
embedding may require gradients too — both are supported.

``forward_batched_train`` is the differentiable twin of ``HIComProjector.forward_batched``.  Its forward runs the same
sm_100a kernels as inference; every op that has a trainable parameter upstream is wrapped in a
``torch.autograd.Function`` whose backward is composed from the C-ABI blocks of ``csrc/backward.cu``
(``hicom_gemm`` for every contraction, ``hicom_act_backward``, ``hicom_softmax_backward``,
``hicom_local_attend_backward``, ``hicom_film_layernorm_backward``, ``hicom_mix_layernorm_backward``).  PyTorch only supplies plumbing: views,
``cat``/``expand`` of token rows, dtype casts, and reductions over tensors of a few hundred rows.

Global attention backward in the reassociated form (forward: projector.py:180-226 as ``pooled = softmax(x'·qfold)ᵀ x'``):
    dP = x'·dpooledᵀ,   delta_j = pooled_j·dpooled_j,   dS = P ∘ (dP − delta),   dqfold = dSᵀ·x'
so no gradient is ever formed for the N x 1152 keys/values the reference materialises.

Status: first correct CUDA path (SIMT GEMMs, fp32 accumulation), validated on a B200 against PyTorch autograd through
the oracle (tests/test_gpu_autograd.py).  Supported: every ``use_guide`` mode (None/off/direct/coarse/fine) and every
adapter (``adaptq/k/v/g``); ``use_clip_scale`` and a ``frames_feature`` that requires grad raise
``NotImplementedError`` — nothing ever returns a tensor silently cut off from the graph.  ``HICOM_AUTOGRAD=0`` / ``hicom_b200.autograd.enable(False)`` restores
the forward-only behaviour (a forward that needs gradients raises, ``projector._require_no_grad``).
"""
from __future__ import annotations

import math
import os
import torch
import torch.nn as nn
from torch.autograd import Function

from . import ops

ENABLED = os.environ.get("HICOM_AUTOGRAD", "1") == "1"


def enable(on: bool = True) -> None:
    """Route grad-enabled projector forwards through ``forward_batched_train`` instead of raising."""
    global ENABLED
    ENABLED = bool(on)


def _impl():
    from . import projector
    return projector._IMPL


# ------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------
def _gemm_to(A: torch.Tensor, B: torch.Tensor, want: torch.dtype, alpha: float = 1.0) -> torch.Tensor:
    """alpha * A @ B on hicom_gemm (strided views, fp32 accumulation), result cast to ``want``."""
    if A.dtype == torch.bfloat16 and B.dtype == torch.float32:
        A = A.float()
    # a transposed A (dW = dYᵀ·A, dqfold = dSᵀ·x') accumulates over thousands of rows into a small result: keep it fp32
    # (that is also the form the tensor-core path serves) and round once
    out_fp32 = A.dtype != B.dtype or (A.dtype == torch.bfloat16 and (want == torch.float32 or A.stride(-2) == 1))
    C = ops.gemm(A, B, None, out_fp32, alpha)
    return C if C.dtype == want else C.to(want)


def _colsum_to(x2: torch.Tensor, want: torch.dtype) -> torch.Tensor:
    """Column sums of (M, N) — bias gradients (db = 1ᵀ·dpre): hicom_colsum, or 1ᵀ·x on hicom_gemm for odd shapes."""
    if ops.colsum_supported(x2):
        s = ops.colsum(x2)
        return s if s.dtype == want else s.to(want)
    ones = torch.ones((1, x2.shape[0]), dtype=x2.dtype, device=x2.device)
    return _gemm_to(ones, x2, want).reshape(-1)


# ------------------------------------------------------------------------------------------
# nn.Linear / build_mlp stage                                           projector.py:307-312
# ------------------------------------------------------------------------------------------
class LinearFn(Function):
    """y = act(A·Wᵀ + bias) [+ residual]; backward: dpre = dy ∘ act'(pre), dA = dpre·W, dW = dpreᵀ·A, db = 1ᵀ·dpre."""

    @staticmethod
    def forward(ctx, A, W, bias, residual, act, out_fp32):
        ctx.save_for_backward(A, W, bias)
        ctx.act = act
        ctx.res_dtype = None if residual is None else residual.dtype
        return ops.linear(A, W, bias, residual, act, out_fp32, _impl())

    @staticmethod
    def backward(ctx, dY):
        A, W, bias = ctx.saved_tensors
        N, K = W.shape
        A2 = A.reshape(-1, K)
        M = A2.shape[0]
        dY2 = dY.reshape(M, N)
        need_A, need_W, need_b, need_R = ctx.needs_input_grad[:4]
        dR = dY.to(ctx.res_dtype) if (need_R and ctx.res_dtype is not None) else None
        if ctx.act != ops.ACT_NONE:
            pre = ops.linear(A2, W, bias, None, ops.ACT_NONE, True, _impl())  # fp32 pre-activations, recomputed
            dpre = ops.act_backward(pre, dY2, ctx.act)
        else:
            dpre = dY2
        dA = _gemm_to(dpre, W, A.dtype).reshape(A.shape) if need_A else None
        dW = _gemm_to(dpre.t(), A2, W.dtype) if need_W else None
        db = _colsum_to(dpre, bias.dtype) if (need_b and bias is not None) else None
        return dA, dW, db, dR, None, None


def _as(p, ref: torch.Tensor):
    """Parameter ``p`` in the dtype of the activations ``ref``.  fp32 master weights under bf16 activations (HF Trainer
    with ``--bf16`` and no DeepSpeed bf16 engine: ``torch.autocast`` semantics) are cast per use, as autocast does for
    ``F.linear``; the cast is an autograd op, so the gradient arrives in the parameter's own dtype."""
    return p if (p is None or p.dtype == ref.dtype) else p.to(ref.dtype)


def linear(A, W, bias=None, residual=None, act=ops.ACT_NONE, out_fp32=False):
    return LinearFn.apply(A, _as(W, A), _as(bias, A), _as(residual, A), act, out_fp32)


def mlp(seq: nn.Sequential, x: torch.Tensor, out_fp32: bool = False):
    """build_mlp Sequential (projector.py:307-312) on differentiable linears (GELU fused in all but the last)."""
    linears = [m for m in seq if isinstance(m, nn.Linear)]
    for i, lin in enumerate(linears):
        last = i == len(linears) - 1
        x = linear(x, lin.weight, lin.bias, None, ops.ACT_NONE if last else ops.ACT_GELU, out_fp32 and last)
    return x


# ------------------------------------------------------------------------------------------
# global attention, reassociated                                         projector.py:180-226
# ------------------------------------------------------------------------------------------
class FoldQueryFn(Function):
    """qfold[b, h*Q+i, :] = alpha * q[b, i, head h] · Wk[head h rows, :]  (ops.global_fold_query).
    ``bk`` (k_proj.bias) shifts every score of a column by one constant, which the softmax cancels: it takes no part in
    the forward and its gradient is exactly zero — returned as zeros (not None) so that every trainable parameter
    receives a gradient, as under the reference's autograd (DDP / ZeRO reducers expect that)."""

    @staticmethod
    def forward(ctx, q, Wk, bk, heads, alpha):
        ctx.save_for_backward(q, Wk)
        ctx.heads, ctx.alpha = heads, alpha
        ctx.bk_meta = None if bk is None else (bk.shape, bk.dtype, bk.device)
        return ops.global_fold_query(q, Wk, heads, alpha)

    @staticmethod
    def backward(ctx, dqfold):
        q, Wk = ctx.saved_tensors
        heads, alpha = ctx.heads, ctx.alpha
        B, Q, d = q.shape
        hd = d // heads
        dqf = dqfold.contiguous().view(B, heads, Q, d)                      # [b, h, i, k]
        dq = dWk = None
        if ctx.needs_input_grad[0]:
            if dqf.dtype == Wk.dtype:
                # dq[b, i, head h] = alpha * Wk[head h rows] · dqfold[b, h*Q+i]: the value projection's formula with Wk in
                # Wv's place -- one dense tensor-core GEMM whose epilogue keeps the diagonal head blocks
                dq = ops.global_value_proj(dqf.view(B, heads * Q, d), Wk, None, Q, heads).mul_(alpha).to(q.dtype)
            else:
                dq = torch.empty_like(q)
                wk_t = Wk.view(heads, hd, d).transpose(1, 2).float()        # [h, k, c]
                dq_view = dq.view(B, Q, heads, hd).permute(0, 2, 1, 3)      # [b, h, i, c]
                dq_view.copy_(ops.gemm(dqf.float(), wk_t, None, True, alpha))
        if ctx.needs_input_grad[1]:
            q_h = q.contiguous().view(B * Q, heads, hd).permute(1, 2, 0)    # [h, c, (b,i)]
            dqf_h = dqf.permute(1, 0, 2, 3).reshape(heads, B * Q, d)        # [h, (b,i), k]  (small copy)
            dWk = _gemm_to(q_h, dqf_h, Wk.dtype, alpha).reshape(d, d)
        dbk = None
        if ctx.needs_input_grad[2] and ctx.bk_meta is not None:
            shape, dtype, device = ctx.bk_meta
            dbk = torch.zeros(shape, dtype=dtype, device=device)
        return dq, dWk, dbk, None, None


class ValueProjFn(Function):
    """attn[b, i, head h] = Wv[head h rows] · pooled[b, h*Q+i] + bv  (ops.global_value_proj)."""

    @staticmethod
    def forward(ctx, pooled, Wv, bv, Q, heads):
        ctx.save_for_backward(pooled, Wv, bv)
        ctx.Q, ctx.heads = Q, heads
        return ops.global_value_proj(pooled, Wv, bv, Q, heads)

    @staticmethod
    def backward(ctx, dattn):
        pooled, Wv, bv = ctx.saved_tensors
        Q, heads = ctx.Q, ctx.heads
        B, J, d = pooled.shape
        hd = d // heads
        da = dattn.contiguous()
        dpooled = dWv = dbv = None
        if ctx.needs_input_grad[0]:
            if da.dtype == Wv.dtype:
                # dpooled[b, h*Q+i, :] = da[b, i, head h] · Wv[head h rows, :]: the query fold's formula with Wv in Wk's place
                dpooled = ops.global_fold_query(da, Wv, heads, 1.0).to(pooled.dtype)
            else:
                da_h = da.view(B, Q, heads, hd).permute(0, 2, 1, 3).float()  # [b, h, i, c]
                wv_h = Wv.view(heads, hd, d).float()                         # [h, c, k]
                dpooled = ops.gemm(da_h, wv_h, None, False, 1.0).reshape(B, J, d).to(pooled.dtype)
        if ctx.needs_input_grad[1]:
            da_t = da.view(B * Q, heads, hd).permute(1, 2, 0)               # [h, c, (b,i)]
            p_h = pooled.view(B, heads, Q, d).permute(1, 0, 2, 3).reshape(heads, B * Q, d)
            dWv = _gemm_to(da_t, p_h, Wv.dtype).reshape(d, d)
        if ctx.needs_input_grad[2] and bv is not None:
            dbv = _colsum_to(da.view(B * Q, d), bv.dtype)
        return dpooled, dWv, dbv, None, None


class GlobalPoolFn(Function):
    """pooled[b, j, :] = sum_n softmax_n(x'_n · qfold[b, j]) x'_n,  x' = X + pos_embed  (split-softmax kernels + merge).
    Differentiable in ``qfold`` and — when the SigLIP body is tuned (train.py:712-715) — in ``X``:
    dx'_n = sum_j P[n,j] dpooled[j] + sum_j dS[n,j] qfold[j]."""

    @staticmethod
    def forward(ctx, X, qfold, gc, t0, splits):
        m, l, o = gc.partials(X, qfold, t0, splits)
        pooled, lse = ops.softmax_merge_lse(m, l, o, ops.out_code(X.dtype))  # lse (B, J) fp32
        ctx.save_for_backward(X, qfold, pooled, lse)
        ctx.gc, ctx.t0 = gc, t0
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        X, qfold, pooled, lse = ctx.saved_tensors
        need_x, need_q = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_x or need_q):
            return None, None, None, None, None
        B, T, H, W, d = X.shape
        N = T * H * W
        pt, ph, pw = ctx.gc.pos_tables(ctx.t0, T, H, W, X.device)
        Xp = ops.posadd(X, pt, ph, pw).view(B, N, d)                        # explicit x' (projector.py:636-640)
        dpl = dpooled.contiguous().to(X.dtype)
        S = ops.gemm(Xp, qfold.transpose(1, 2), None, True, 1.0)            # (B, N, J) fp32 scores
        if X.dtype != torch.float32:
            # the 16-bit forward applied the position terms in higher precision than this x' (rounded to 16 bits): take
            # the log-sum-exp of the backward's OWN scores so that exp(S - lse) sums to one exactly (no systematic bias
            # in dS, hence in the q_proj / k_proj / query / guide gradients)
            lse = ops.col_logsumexp(S)
        dP = ops.gemm(Xp, dpl.transpose(1, 2), None, True, 1.0)             # (B, N, J) fp32
        delta = (pooled.float() * dpl.float()).sum(-1)                      # (B, J)
        P = None
        if need_x:  # the probabilities themselves (softmax_backward overwrites nothing, but S is consumed below)
            P = torch.exp(S - lse[:, None, :]).to(X.dtype)
        dS = ops.softmax_backward(S, dP, lse, delta, X.dtype == torch.bfloat16)
        dqfold = _gemm_to(dS.transpose(1, 2), Xp, qfold.dtype) if need_q else None   # (B, J, d)
        dX = None
        if need_x:
            dX = ops.gemm(P, dpl, None, True, 1.0)                                  # (B, N, d) fp32
            dX += ops.gemm(dS.to(X.dtype), qfold, None, True, 1.0)
            dX = dX.view(X.shape).to(X.dtype)
        return dX, dqfold, None, None, None


class GlobalPoolKeysFn(Function):
    """use_clip_scale (projector.py:184-188): pooled[b, j, :] = sum_n softmax_n(kn_n · qs[b, j]) x'_n with EXPLICIT
    normalised keys ``Kn`` (B,T,H,W,d) and score columns ``qs`` (B,J,d) = exp(logit_scale) x the normalised query masked
    to its head's channels.  Differentiable in Kn (-> k_proj), qs (-> q_proj, logit_scale) and X."""

    @staticmethod
    def forward(ctx, X, Kn, qs, gc, t0, splits):
        B, T, H, W, d = X.shape
        pt, ph, pw = gc.pos_tables(t0, T, H, W, X.device)
        if splits is None:
            from .projector import default_splits
            splits = default_splits(B, T * H * W, d)
        m, l, o = ops.global_attend_partial_keys(X, Kn, pt, ph, pw, qs, splits, _impl())
        pooled, lse = ops.softmax_merge_lse(m, l, o, ops.out_code(X.dtype))
        ctx.save_for_backward(X, Kn, qs, pooled, lse)
        ctx.gc, ctx.t0 = gc, t0
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        X, Kn, qs, pooled, lse = ctx.saved_tensors
        need_x, need_k, need_q = ctx.needs_input_grad[:3]
        if not (need_x or need_k or need_q):
            return None, None, None, None, None, None
        B, T, H, W, d = X.shape
        N = T * H * W
        pt, ph, pw = ctx.gc.pos_tables(ctx.t0, T, H, W, X.device)
        Xp = ops.posadd(X, pt, ph, pw).view(B, N, d)
        Kf = Kn.contiguous().view(B, N, d)
        dpl = dpooled.contiguous().to(X.dtype)
        S = ops.gemm(Kf, qs.transpose(1, 2), None, True, 1.0)               # (B, N, J) fp32
        dP = ops.gemm(Xp, dpl.transpose(1, 2), None, True, 1.0)
        delta = (pooled.float() * dpl.float()).sum(-1)
        P = torch.exp(S - lse[:, None, :]).to(X.dtype) if need_x else None
        dS = ops.softmax_backward(S, dP, lse, delta, X.dtype == torch.bfloat16)
        dSc = dS.to(X.dtype)
        dqs = _gemm_to(dS.transpose(1, 2), Kf, qs.dtype) if need_q else None
        dKn = ops.gemm(dSc, qs, None, True, 1.0).view(X.shape).to(Kn.dtype) if need_k else None
        dX = ops.gemm(P, dpl, None, True, 1.0).view(X.shape).to(X.dtype) if need_x else None
        return dX, dKn, dqs, None, None, None


# ------------------------------------------------------------------------------------------
# coarse injector and window attention                          projector.py:369-372, 546-553
# ------------------------------------------------------------------------------------------
class FilmLayerNormFn(Function):
    """LN(x*(1+scale)+shift)*w + b with film = [scale | shift] (G, 2d) fp32, row r using film[r // rows_per_group]."""

    @staticmethod
    def forward(ctx, x, film, ln_w, ln_b, rows_per_group):
        ctx.save_for_backward(x, film, ln_w)
        ctx.rpg = rows_per_group
        ctx.b_dtype = ln_b.dtype
        return ops.film_layernorm(x, film, ln_w, ln_b, rows_per_group)

    @staticmethod
    def backward(ctx, dy):
        x, film, ln_w = ctx.saved_tensors
        need_x = ctx.needs_input_grad[0]
        dx, dfilm, dw, db = ops.film_layernorm_backward(x, film, ln_w, dy.contiguous().to(x.dtype), ctx.rpg, need_x)
        return (dx.view(x.shape) if need_x else None, dfilm if ctx.needs_input_grad[1] else None,
                dw.to(ln_w.dtype) if ctx.needs_input_grad[2] else None,
                db.to(ctx.b_dtype) if ctx.needs_input_grad[3] else None, None)


class GridPoolFn(Function):
    """Trilinear grid pooling of frames_feature (projector.py:539-540), differentiable in X (only needed when the SigLIP
    body is tuned, train.py:712-715)."""

    @staticmethod
    def forward(ctx, X, kt, ks):
        ctx.geom = (X.shape, X.dtype, kt, ks)
        return ops.grid_pool(X, kt, ks)

    @staticmethod
    def backward(ctx, dq):
        (B, T, H, W, d), dtype, kt, ks = ctx.geom
        return ops.grid_pool_backward(dq.contiguous().to(dtype), T, H, W, kt, ks).to(dtype), None, None


class L2NormFn(Function):
    """x / |x| over the last dim (use_clip_scale, projector.py:184-188,527-529)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.l2norm_rows(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.l2norm_rows_backward(x, dy.contiguous().to(x.dtype))


class LocalAttendFn(Function):
    """Window attention on explicit query rows (ops.local_attend, Q_EXPLICIT); differentiable in the query rows, the
    keys (``frames_embed`` / the key adapter's output) and the values (only behind a trainable value adapter — plain
    values are ``frames_feature`` of the frozen SigLIP body)."""

    @staticmethod
    def forward(ctx, K, V, Qrows, kt, ks, scale, k_l2norm, logit_scale=None):
        """``logit_scale`` (0-dim tensor, use_clip_scale): scores = exp(logit_scale) * q·k (projector.py:549); ``scale`` is
        its value as a float.  d/d(logit_scale) = sum over windows of q·dq (dq already carries the factor)."""
        ctx.save_for_backward(K, V, Qrows)
        ctx.geom = (kt, ks, scale, k_l2norm)
        ctx.ls_meta = None if logit_scale is None else (logit_scale.shape, logit_scale.dtype)
        return ops.local_attend(K, V, V, Qrows, None, None, None, kt, ks, ops.Q_EXPLICIT, scale, k_l2norm)

    @staticmethod
    def backward(ctx, dO):
        K, V, Qrows = ctx.saved_tensors
        kt, ks, scale, k_l2norm = ctx.geom
        need_k, need_v, need_q = ctx.needs_input_grad[:3]
        need_ls = len(ctx.needs_input_grad) > 7 and ctx.needs_input_grad[7] and ctx.ls_meta is not None
        dQ = dK = dV = dls = None
        if need_q or need_k or need_v or need_ls:
            dQ, dK, dV = ops.local_attend_backward(K, V, Qrows, dO.contiguous().to(Qrows.dtype), kt, ks, scale,
                                                   k_l2norm, need_q or need_ls, need_k, need_v)
            dK = None if dK is None else dK.to(K.dtype)
            dV = None if dV is None else dV.to(V.dtype)
            if need_ls:
                shape, dtype = ctx.ls_meta
                dls = (Qrows.float() * dQ.float()).sum().reshape(shape).to(dtype)
            if not need_q:
                dQ = None
        return dK, dV, dQ, None, None, None, None, dls


class LayerNormFn(Function):
    """LN(x)*w + b (ops.layernorm, eps 1e-6) — the SigLIP head's layernorm in the producer (encoder.py:284).  Backward
    on hicom_film_layernorm_backward with a zero FiLM (scale = shift = 0 leaves u = x)."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.b_dtype = b.dtype
        return ops.layernorm(x, w, b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        d = x.shape[-1]
        rows = x.numel() // d
        film = torch.zeros((1, 2 * d), dtype=torch.float32, device=x.device)
        need_x = ctx.needs_input_grad[0]
        dx, _, dw, db = ops.film_layernorm_backward(x, film, w, dy.contiguous().to(x.dtype), max(rows, 1), need_x)
        return (dx.view(x.shape) if need_x else None, dw.to(w.dtype) if ctx.needs_input_grad[1] else None,
                db.to(ctx.b_dtype) if ctx.needs_input_grad[2] else None)


class MixLayerNormFn(Function):
    """(1 - alpha) * x + alpha * LN(y) — the adapter mixes (ops.mix_layernorm; projector.py:365,533-534,541)."""

    @staticmethod
    def forward(ctx, x, y, ln_w, ln_b, alpha):
        alpha = alpha.to(x.dtype)
        ctx.save_for_backward(x, y, ln_w, ln_b, alpha)
        return ops.mix_layernorm(x, y, ln_w, ln_b, alpha)

    @staticmethod
    def backward(ctx, dout):
        x, y, ln_w, ln_b, alpha = ctx.saved_tensors
        need_x = ctx.needs_input_grad[0]
        dx, dy, dw, db, da = ops.mix_layernorm_backward(x, y, ln_w, ln_b, alpha, dout.contiguous().to(x.dtype), need_x)
        return (dx, dy if ctx.needs_input_grad[1] else None, dw.to(ln_w.dtype) if ctx.needs_input_grad[2] else None,
                db.to(ln_b.dtype) if ctx.needs_input_grad[3] else None,
                da.to(alpha.dtype).reshape(alpha.shape) if ctx.needs_input_grad[4] else None)


class AddLayerNormFn(Function):
    """LN(a + b) — the fine injector's residual (ops.add_layernorm, projector.py:392); backward through the FiLM-LN
    backward kernel with a zero FiLM."""

    @staticmethod
    def forward(ctx, a, b, ln_w, ln_b):
        ctx.save_for_backward(a, b, ln_w)
        ctx.b_dtype = ln_b.dtype
        return ops.add_layernorm(a, b, ln_w, ln_b)

    @staticmethod
    def backward(ctx, dy):
        a, b, ln_w = ctx.saved_tensors
        d = a.shape[-1]
        u = a + b                                                            # small: query rows only
        film = torch.zeros((1, 2 * d), dtype=torch.float32, device=a.device)
        du, _, dw, db = ops.film_layernorm_backward(u, film, ln_w, dy.contiguous().to(a.dtype), max(u.numel() // d, 1),
                                                    True)
        du = du.view(a.shape)
        return (du if ctx.needs_input_grad[0] else None, du if ctx.needs_input_grad[1] else None,
                dw.to(ln_w.dtype) if ctx.needs_input_grad[2] else None,
                db.to(ctx.b_dtype) if ctx.needs_input_grad[3] else None)


class GuideAttendFn(Function):
    """Multi-head attention of query rows over the L instruction tokens (ops.guide_attend; projector.py:391 ->
    :193-224).  Backward composed from hicom_gemm over per-head strided views and hicom_softmax_backward (the softmax
    runs over the L keys: scores are laid out (b*h, L, rows) so that the key axis is the one the kernel normalises)."""

    @staticmethod
    def forward(ctx, q, k, v, heads, scale):
        ctx.save_for_backward(q, k, v)
        ctx.heads, ctx.scale = heads, scale
        return ops.guide_attend(q, k, v, heads, scale)

    @staticmethod
    def backward(ctx, dout):
        q, k, v = ctx.saved_tensors
        heads, scale = ctx.heads, ctx.scale
        B, n, d = q.shape
        L = k.shape[1]
        hd = d // heads
        split = lambda t, rows: t.contiguous().view(B, rows, heads, hd).permute(0, 2, 1, 3)   # [b, h, row, c] view
        qh, kh, vh, doh = split(q, n), split(k, L), split(v, L), split(dout.to(q.dtype), n)
        St = ops.gemm(kh, qh.transpose(-1, -2), None, True, scale).view(B * heads, L, n)      # scale * k_l·q_i
        dPt = ops.gemm(vh, doh.transpose(-1, -2), None, True, 1.0).view(B * heads, L, n)      # v_l·do_i
        lse = torch.logsumexp(St, dim=1)                                                       # (b*h, rows), small
        zero = torch.zeros_like(lse)
        P = ops.softmax_backward(St, torch.ones_like(St), lse, zero, False)                   # exp(S - lse)
        delta = (P * dPt).sum(dim=1)
        dSt = ops.softmax_backward(St, dPt, lse, delta, False).view(B, heads, L, n)
        P = P.view(B, heads, L, n)
        merge = lambda t, rows: t.permute(0, 2, 1, 3).reshape(B, rows, d)
        dq = dk = dv = None
        if ctx.needs_input_grad[0]:
            dq = merge(ops.gemm(dSt.transpose(-1, -2), kh, None, True, scale), n).to(q.dtype)
        if ctx.needs_input_grad[1]:
            dk = merge(ops.gemm(dSt, qh, None, True, scale), L).to(k.dtype)
        if ctx.needs_input_grad[2]:
            dv = merge(ops.gemm(P, doh, None, True, 1.0), L).to(v.dtype)
        return dq, dk, dv, None, None


# ------------------------------------------------------------------------------------------
# the differentiable forward
# ------------------------------------------------------------------------------------------
def _is_param(x) -> bool:
    return isinstance(x, torch.Tensor)


def check_supported(proj, X, E, G) -> None:
    if X.dtype not in (torch.float32, torch.bfloat16):
        raise NotImplementedError(f"hicom_b200.autograd: dtype {X.dtype} (train in fp32 or bf16)")
    for comp in (proj.local_compressor, proj.global_compressor):
        if comp is not None and comp.use_guide not in (None, "off", "direct", "coarse", "fine"):
            raise NotImplementedError  # projector.py:350


def _mix(x, proj, norm, alpha):
    """(1-α)x + α·LN(proj(x)) (projector.py:365,533-534,541), differentiable; the identity when not adapting."""
    if not _is_param(alpha):
        return x
    y = mlp(proj, x) if isinstance(proj, nn.Sequential) else linear(x, proj.weight, proj.bias)
    return MixLayerNormFn.apply(x, y, _as(norm.weight, x), _as(norm.bias, x), alpha)


def _prepared_guide(inj, G):
    """GuideInjector.prepared_guide with gradients: text2qk projection + guide adapter (projector.py:364-365,388-389)."""
    if isinstance(inj.text2qk_proj, nn.Sequential):
        G = mlp(inj.text2qk_proj, G)
    return _mix(G, inj.guide_proj, inj.guide_norm, inj.guide_alpha)


def _inject(inj, mode, rows, G):
    """GuideInjector.forward on explicit rows (B, n, d) (projector.py:344-397), differentiable."""
    if mode in (None, "off"):
        return rows
    inj.check_guide(G, 0)
    B, n, d = rows.shape
    g = _prepared_guide(inj, G)
    if mode == "direct":                                                   # :367-368: the rows' content is discarded
        return g.unsqueeze(1).expand(B, n, d)                              # autograd sums over the rows
    if mode == "coarse":                                                   # :369-372
        film = mlp(inj.coarse_proj, g, out_fp32=True)
        return FilmLayerNormFn.apply(rows.contiguous(), film, _as(inj.coarse_norm.weight, rows),
                                     _as(inj.coarse_norm.bias, rows), n)
    mha = inj.fine_proj                                                    # fine, :374-397
    q = linear(rows, mha.q_proj.weight, mha.q_proj.bias)
    k = linear(g, mha.k_proj.weight, mha.k_proj.bias)
    v = linear(g, mha.v_proj.weight, mha.v_proj.bias)
    a = GuideAttendFn.apply(q, k, v, mha.num_heads, mha.scale)
    a = linear(a, mha.out_proj.weight, mha.out_proj.bias)
    return AddLayerNormFn.apply(rows.contiguous(), a, _as(inj.fine_norm.weight, rows), _as(inj.fine_norm.bias, rows))


def _local_tokens(proj, X, E, G, modal, image_newline, is_anyres):
    lc = proj.local_compressor
    B, T, H, W, d = X.shape
    tk, grid = lc.output_grid(T, H, W, modal)
    sk = lc.spatial_kernel_size
    n_local, plan = proj._local_rows(grid, modal, image_newline, is_anyres)
    if plan != "plain" and image_newline is None:
        raise ValueError("this mm_newline_position needs image_newline")
    mode = lc.use_guide
    ls, lb = proj.local_logit_scale, proj.local_logit_bias                        # use_clip_scale 'local' (:667-668)
    ls_grad = torch.is_tensor(ls) and ls.requires_grad
    adapting = any(_is_param(a) for a in (lc.q_alpha, lc.k_alpha, lc.v_alpha))
    upstream = (mode in ("coarse", "fine") or adapting or ls_grad or X.requires_grad
                or (mode == "direct" and ((G is not None and G.requires_grad)
                                          or _is_param(lc.guide_injector.guide_alpha)))
                or (E is not None and E.requires_grad))
    if not upstream:
        # nothing trainable in front of the readout (query = pooled feature or the frozen instruction vector, keys and
        # values from frozen towers): the fused inference kernel, as a constant
        with torch.no_grad():
            att = lc.attend(X, E, G, modal, ls, lb)
    else:
        if E is not None and ls is not None:                                       # :527-529
            E = L2NormFn.apply(E)
            if G is not None:
                G = G / G.norm(p=2, dim=-1, keepdim=True)
        K = _mix(X if E is None else E, lc.k_proj, lc.k_norm, lc.k_alpha)          # projector.py:532-533
        V = _mix(X, lc.v_proj, lc.v_norm, lc.v_alpha)                              # :534
        if mode == "direct":
            rows = torch.empty((B, ops.num_windows(T, H, W, tk, sk), d), dtype=X.dtype, device=X.device)  # shape only
        else:
            if X.requires_grad:                                                    # SigLIP body tuned (train.py:712-715)
                rows = GridPoolFn.apply(X, tk, sk)
            else:
                with torch.no_grad():
                    rows = ops.grid_pool(X, tk, sk)                                # :539-540 (no parameter)
            rows = _mix(rows, lc.q_proj, lc.q_norm, lc.q_alpha)                    # :541
        rows = _inject(lc.guide_injector, mode, rows, G)                           # :542
        if ls is None:
            att = LocalAttendFn.apply(K, V, rows, tk, sk, 1.0 / math.sqrt(lc.qk_dim), False)   # :544-558
        else:                                                                      # :548-549, logit_bias cancels
            from .projector import _exp_scalar
            att = LocalAttendFn.apply(K, V, rows.contiguous(), tk, sk, _exp_scalar(ls), False, ls if ls_grad else None)
            if torch.is_tensor(lb) and lb.requires_grad:  # softmax ignores the bias: an exact zero gradient, not None
                att = att + (lb * 0).to(att.dtype)
    tokens = mlp(lc.readout, att)                                                  # (B, Nw, Dh), projector.py:559
    Dh = tokens.shape[-1]
    t1, h1, w1 = grid
    if plan == "plain":
        return tokens
    nl = image_newline.to(tokens.dtype)
    if plan == "tail":                                                             # mm_utils.py:115-117
        return torch.cat([tokens, nl.expand(B, 1, Dh)], dim=1)
    if plan == "grid":                                                             # mm_utils.py:101-107
        blk = torch.cat([tokens.view(B, t1 * h1, w1, Dh), nl.expand(B, t1 * h1, 1, Dh)], dim=2)
    else:                                                                          # "frame", mm_utils.py:108-114
        blk = torch.cat([tokens.view(B, t1, h1 * w1, Dh), nl.expand(B, t1, 1, Dh)], dim=2)
    return blk.reshape(B, n_local, Dh)


def _masked_heads(qn: torch.Tensor, heads: int) -> torch.Tensor:
    """(B, Q, d) -> (B, heads*Q, d): score column (h, i) keeps head h's channels of query i, zeros elsewhere
    (differentiable: a product with a constant block mask)."""
    B, Q, d = qn.shape
    hd = d // heads
    mask = torch.zeros((heads, 1, heads, 1), dtype=qn.dtype, device=qn.device)
    ar = torch.arange(heads, device=qn.device)
    mask[ar, 0, ar, 0] = 1
    return (qn.view(B, 1, Q, heads, hd) * mask.view(1, heads, 1, heads, 1)).reshape(B, heads * Q, d)


def _global_tokens(proj, X, G, splits=None, t0=0):
    gc = proj.global_compressor
    attn = gc.attn_layer
    B = X.shape[0]
    nq, d = gc.query.shape
    if gc.use_guide == "direct":          # projector.py:367-368: every query row is the guide — one distinct row
        gc.guide_injector.check_guide(G, 0)
        Qg = _prepared_guide(gc.guide_injector, G).unsqueeze(1).contiguous()        # (B, 1, d)
    else:
        rows = gc.query.to(X.dtype).unsqueeze(0).expand(B, nq, d).contiguous()      # autograd sums over the batch
        Qg = _inject(gc.guide_injector, gc.use_guide, rows, G)                      # projector.py:642
    nrows = Qg.shape[1]
    q = linear(Qg, attn.q_proj.weight, attn.q_proj.bias)                            # projector.py:180
    ls, lb = proj.global_logit_scale, proj.global_logit_bias                        # use_clip_scale 'global' (:669-670)
    if ls is None:
        qfold = FoldQueryFn.apply(q, _as(attn.k_proj.weight, q), _as(attn.k_proj.bias, q), attn.num_heads,
                                  attn.scale)                                       # :181 + :197 folded
        pooled = GlobalPoolFn.apply(X, qfold, gc, t0, splits)                       # :197-215
    else:
        # :184-188 — queries and keys L2-normalised over all d channels before the head split: the key norm cannot be
        # folded into the queries, so the keys are explicit: k = k_proj(x + pos_embed)
        heads, hd = attn.num_heads, d // attn.num_heads
        B_, T, H, W, _ = X.shape
        pt, ph, pw = gc.pos_tables(t0, T, H, W, X.device)
        if X.requires_grad:  # differentiable x' = x + pos_embed (:636-640)
            Xp = X + (pt[None, :, None, None, :] + ph[None, None, :, None, :] + pw[None, None, None, :, :]).to(X.dtype)
        else:
            with torch.no_grad():
                Xp = ops.posadd(X, pt, ph, pw)
        Kn = L2NormFn.apply(linear(Xp.view(B_, T * H * W, d), attn.k_proj.weight, attn.k_proj.bias)).view(X.shape)
        scale = torch.as_tensor(ls, device=q.device).exp().to(q.dtype)
        qn = L2NormFn.apply(q) * scale
        qs = _masked_heads(qn, heads)
        pooled = GlobalPoolKeysFn.apply(X, Kn, qs, gc, t0, splits)
        if torch.is_tensor(lb) and lb.requires_grad:  # the bias shifts every score of a row equally: exact zero gradient
            pooled = pooled + (lb * 0).to(pooled.dtype)
    a = ValueProjFn.apply(pooled, _as(attn.v_proj.weight, pooled), _as(attn.v_proj.bias, pooled), nrows,
                          attn.num_heads)                                           # :182, :223-224
    x = linear(a, attn.out_proj.weight, attn.out_proj.bias, Qg)                     # :226 + residual of :646
    tokens = mlp(gc.readout, x)                                                     # (B, nrows, Dh)
    if nrows != nq:
        tokens = tokens.expand(B, nq, tokens.shape[-1])
    return tokens


def forward_batched_train(proj, X, E, G, modal, image_newline=None, is_anyres=False, base=None, with_global=True):
    """Differentiable ``HIComProjector.forward_batched``: same arguments, same ``(B, n_tokens, Dh)`` result, with an
    autograd graph reaching every projector parameter (and ``image_newline`` / ``base`` / ``frames_embed`` / the
    instruction embedding when they require grad)."""
    check_supported(proj, X, E, G)
    ops._need_cuda(X, E, G)  # raises for CPU tensors: there is no CPU fallback
    parts = [] if base is None else [base]
    if proj.local_compressor is not None:
        parts.append(_local_tokens(proj, X, E, G, modal, image_newline, is_anyres))
    if proj.global_compressor is not None and with_global:
        parts.append(_global_tokens(proj, X, G))
    if not parts:
        return None
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
