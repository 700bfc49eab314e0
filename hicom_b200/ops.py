"""Thin torch custom-op layer over the C ABI (include/hicom_b200.h).

Each op validates its tensors, allocates the result with torch's caching allocator, and passes raw
device pointers plus the CURRENT CUDA stream to ``libhicom_b200.so`` — nothing here computes.  The ops
are registered as ``torch.ops.hicom_b200.*`` (no autograd formula is registered on the ops themselves: gradients are
the business of ``hicom_b200/autograd.py``, which wraps them in ``torch.autograd.Function``s and composes their
backward from the backward blocks at the end of this file).  There is no CPU or PyTorch
fallback: CPU tensors raise, a missing library raises.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _cabi
from ._cabi import ACT_GELU, ACT_GELU_TANH, ACT_NONE, IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05  # noqa: F401 (re-exported)
from ._cabi import Q_EXPLICIT, Q_FILM_LN, Q_POOLED, Q_VECTOR  # noqa: F401

_DT = {torch.float32: _cabi.F32, torch.bfloat16: _cabi.BF16, torch.float16: _cabi.F16}
_CODE_DT = {v: k for k, v in _DT.items()}


def _dt(t: Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"hicom_b200 supports float32, bfloat16 and float16 tensors, got {t.dtype}") from None


def _need_cuda(*ts: Optional[Tensor]):
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("hicom_b200 ops run on CUDA tensors only (no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def _ptr(t: Optional[Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(dev) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _c(t: Optional[Tensor]) -> Optional[Tensor]:
    return None if t is None else t.contiguous()


# ------------------------------------------------------------------------------------------
# local compressor
# ------------------------------------------------------------------------------------------
def num_windows(T: int, H: int, W: int, kt: int, ks: int) -> int:
    return math.ceil(T / kt) * math.ceil(H / ks) * math.ceil(W / ks)


def _impl_grid_pool(X: Tensor, kt: int, ks: int) -> Tensor:
    """(B,T,H,W,d) -> (B,Nw,d) trilinear grid pooling — projector.py:536-540."""
    dev = _need_cuda(X)
    X = X.contiguous()
    B, T, H, W, d = X.shape
    out = torch.empty((B, num_windows(T, H, W, kt, ks), d), dtype=X.dtype, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_grid_pool(_ptr(X), _ptr(out), B, T, H, W, d, kt, ks, _dt(X), _stream(dev))
    _cabi.check(rc, "hicom_grid_pool")
    return out


def _impl_local_attend_into(K: Tensor, V: Tensor, P: Tensor, q_aux: Optional[Tensor], film: Optional[Tensor],
                            ln_w: Optional[Tensor], ln_b: Optional[Tensor], kt: int, ks: int, qmode: int,
                            logit_scale: float, k_l2norm: bool, out: Tensor) -> None:
    """Fused pool -> inject -> window softmax -> A·V — projector.py:536-558 — written into ``out`` (B,Nw,d), which may
    be a leading-dim slice of a larger buffer (the forward computes the batch in two halves)."""
    dev = _need_cuda(K, V, P, q_aux, film, ln_w, ln_b, out)
    same_kv = K.data_ptr() == V.data_ptr()
    same_pv = P.data_ptr() == V.data_ptr()
    V = V.contiguous()
    K = V if same_kv else K.contiguous()
    P = V if same_pv else P.contiguous()
    if not (K.shape == V.shape == P.shape) or V.dim() != 5:
        raise ValueError(f"local_attend: K/V/P must share a (B,T,H,W,d) shape, got {K.shape} {V.shape} {P.shape}")
    if not (K.dtype == V.dtype == P.dtype):
        raise TypeError("local_attend: K/V/P dtypes differ")
    B, T, H, W, d = V.shape
    nw = num_windows(T, H, W, kt, ks)
    q_aux, ln_w, ln_b = _c(q_aux), _c(ln_w), _c(ln_b)
    if film is not None:
        film = film.contiguous()
        if film.dtype != torch.float32 or film.shape != (B, 2 * d):
            raise ValueError("local_attend: film must be fp32 (B, 2d)")
    for t in (q_aux, ln_w, ln_b):
        if t is not None and t.dtype != V.dtype:
            raise TypeError("local_attend: q_aux / ln params must have the feature dtype")
    if qmode == Q_VECTOR and (q_aux is None or q_aux.shape != (B, d)):
        raise ValueError("local_attend: Q_VECTOR needs q_aux (B,d)")
    if qmode == Q_EXPLICIT and (q_aux is None or q_aux.shape != (B, nw, d)):
        raise ValueError(f"local_attend: Q_EXPLICIT needs q_aux (B,{nw},{d})")
    if out.shape != (B, nw, d) or out.dtype != V.dtype or not out.is_contiguous():
        raise ValueError(f"local_attend: out must be a contiguous (B,{nw},{d}) tensor in the feature dtype")
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_local_attend(_ptr(K), _ptr(V), _ptr(P), _ptr(q_aux), _ptr(film), _ptr(ln_w),
                                             _ptr(ln_b), _ptr(out), B, T, H, W, d, kt, ks, qmode,
                                             float(logit_scale), int(k_l2norm), _dt(V), _stream(dev))
    _cabi.check(rc, "hicom_local_attend")


def _impl_local_attend(K: Tensor, V: Tensor, P: Tensor, q_aux: Optional[Tensor], film: Optional[Tensor],
                       ln_w: Optional[Tensor], ln_b: Optional[Tensor], kt: int, ks: int, qmode: int,
                       logit_scale: float, k_l2norm: bool) -> Tensor:
    """Fused pool -> inject -> window softmax -> A·V — projector.py:536-558.  Returns (B,Nw,d)."""
    B, T, H, W, d = V.shape
    out = torch.empty((B, num_windows(T, H, W, kt, ks), d), dtype=V.dtype, device=V.device)
    _impl_local_attend_into(K, V, P, q_aux, film, ln_w, ln_b, kt, ks, qmode, logit_scale, k_l2norm, out)
    return out


# ------------------------------------------------------------------------------------------
# dense layers
# ------------------------------------------------------------------------------------------
def _rows2d(t: Tensor) -> Tensor:
    t2 = t.reshape(-1, t.shape[-1])
    if t2.stride(-1) != 1:
        t2 = t2.contiguous()
    return t2


def _linear_call(A2, W, bias, R2, C, ldc, M, N, K, act, out_dtype_code, rpg, gstride, impl, dev):
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_linear(_ptr(A2), A2.stride(0), _ptr(W), W.stride(0), _ptr(bias), _ptr(R2),
                                       0 if R2 is None else R2.stride(0), _ptr(C), ldc, M, N, K, act,
                                       _dt(A2), out_dtype_code, rpg, gstride, impl, _stream(dev))
    _cabi.check(rc, "hicom_linear")


def _check_linear(A, W, bias, residual):
    dev = _need_cuda(A, W, bias, residual)
    if W.dim() != 2 or A.shape[-1] != W.shape[1]:
        raise ValueError(f"linear: A (...,{A.shape[-1]}) does not match W {tuple(W.shape)}")
    if W.dtype != A.dtype or (bias is not None and bias.dtype != A.dtype):
        raise TypeError("linear: A, W, bias must share a dtype")
    if W.stride(1) != 1:
        W = W.contiguous()
    return dev, W


def _impl_linear(A: Tensor, W: Tensor, bias: Optional[Tensor], residual: Optional[Tensor], act: int,
           out_fp32: bool, impl: int) -> Tensor:
    """act(A·Wᵀ + bias) [+ residual] — nn.Linear / build_mlp stage (projector.py:180-182,226,307-312)."""
    dev, W = _check_linear(A, W, bias, residual)
    A2 = _rows2d(A)
    M, K = A2.shape
    N = W.shape[0]
    R2 = None
    if residual is not None:
        if residual.dtype != A.dtype:
            raise TypeError("linear: residual dtype differs")
        R2 = _rows2d(residual)
        if R2.shape != (M, N):
            raise ValueError("linear: residual shape mismatch")
    odt = torch.float32 if out_fp32 else A.dtype
    C = torch.empty((M, N), dtype=odt, device=dev)
    _linear_call(A2, W, _c(bias), R2, C, N, M, N, K, act, _DT[odt], max(M, 1), 0, impl, dev)
    return C.reshape(*A.shape[:-1], N)


def _impl_linear_into(A: Tensor, W: Tensor, bias: Optional[Tensor], residual: Optional[Tensor], act: int,
                out: Tensor, row_offset: int, rows_per_group: int, group_stride_rows: int, impl: int) -> None:
    """Same as ``linear`` but writes row r to ``out[(r // rows_per_group) * group_stride_rows +
    r % rows_per_group + row_offset]`` — the readouts write straight into the concatenated token block
    of every video (projector.py:707), no ``cat``."""
    dev, W = _check_linear(A, W, bias, residual)
    _need_cuda(out)
    A2 = _rows2d(A)
    M, K = A2.shape
    N = W.shape[0]
    if out.dim() != 2 or out.shape[1] != N or out.stride(1) != 1:
        raise ValueError("linear_into: out must be (rows, N) with unit inner stride")
    if M > 0:
        last = ((M - 1) // rows_per_group) * group_stride_rows + (M - 1) % rows_per_group + row_offset
        if last >= out.shape[0] or row_offset < 0:
            raise ValueError("linear_into: destination rows out of range")
    R2 = _rows2d(residual) if residual is not None else None
    Cview = out[row_offset:]
    _linear_call(A2, W, _c(bias), R2, Cview, out.stride(0), M, N, K, act, _dt(out), rows_per_group,
                 group_stride_rows, impl, dev)


def _impl_layernorm(x: Tensor, ln_w: Tensor, ln_b: Tensor) -> Tensor:
    """LN(x) over the last dim, eps 1e-6 — the SigLIP head's layernorm in front of the frames_embed MLP
    (encoder.py:284)."""
    dev = _need_cuda(x, ln_w, ln_b)
    x = x.contiguous()
    d = x.shape[-1]
    if ln_w.shape != (d,) or ln_b.shape != (d,) or ln_w.dtype != x.dtype or ln_b.dtype != x.dtype:
        raise ValueError("layernorm: weight/bias must be (d,) in the input dtype")
    out = torch.empty_like(x)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_layernorm(_ptr(x), _ptr(_c(ln_w)), _ptr(_c(ln_b)), _ptr(out), x.numel() // d, d,
                                          _dt(x), _stream(dev))
    _cabi.check(rc, "hicom_layernorm")
    return out


def _impl_film_layernorm(x: Tensor, film: Tensor, ln_w: Tensor, ln_b: Tensor, rows_per_group: int) -> Tensor:
    """LN(x*(1+scale)+shift), film (G,2d) fp32 — coarse injector on explicit rows (projector.py:369-372)."""
    dev = _need_cuda(x, film, ln_w, ln_b)
    x = x.contiguous()
    d = x.shape[-1]
    rows = x.numel() // d
    film = film.contiguous()
    if film.dtype != torch.float32 or film.shape[-1] != 2 * d:
        raise ValueError("film_layernorm: film must be fp32 (G, 2d)")
    if rows > film.shape[0] * rows_per_group:
        raise ValueError("film_layernorm: not enough film rows")
    out = torch.empty_like(x)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_film_layernorm(_ptr(x), _ptr(film), _ptr(_c(ln_w)), _ptr(_c(ln_b)), _ptr(out),
                                               rows, d, rows_per_group, _dt(x), _stream(dev))
    _cabi.check(rc, "hicom_film_layernorm")
    return out


def _impl_add_layernorm(a: Tensor, b: Tensor, ln_w: Tensor, ln_b: Tensor) -> Tensor:
    """LN(a + b) — fine injector residual (projector.py:392)."""
    dev = _need_cuda(a, b, ln_w, ln_b)
    a, b = a.contiguous(), b.contiguous()
    if a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError("add_layernorm: operands differ")
    d = a.shape[-1]
    out = torch.empty_like(a)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_add_layernorm(_ptr(a), _ptr(b), _ptr(_c(ln_w)), _ptr(_c(ln_b)), _ptr(out),
                                              a.numel() // d, d, _dt(a), _stream(dev))
    _cabi.check(rc, "hicom_add_layernorm")
    return out


def _impl_mix_layernorm(x: Tensor, y: Tensor, ln_w: Tensor, ln_b: Tensor, alpha: Tensor) -> Tensor:
    """(1-alpha)*x + alpha*LN(y) — adapter mixes (projector.py:365,533-534,541)."""
    dev = _need_cuda(x, y, ln_w, ln_b, alpha)
    x, y = x.contiguous(), y.contiguous()
    if x.shape != y.shape or x.dtype != y.dtype or alpha.dtype != x.dtype:
        raise ValueError("mix_layernorm: operands differ")
    d = x.shape[-1]
    out = torch.empty_like(x)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_mix_layernorm(_ptr(x), _ptr(y), _ptr(_c(ln_w)), _ptr(_c(ln_b)),
                                              _ptr(alpha.contiguous()), _ptr(out), x.numel() // d, d, _dt(x),
                                              _stream(dev))
    _cabi.check(rc, "hicom_mix_layernorm")
    return out


def _impl_guide_attend(q: Tensor, k: Tensor, v: Tensor, heads: int, scale: float) -> Tensor:
    """MHA of (G,Mq,d) queries over (G,L,d) instruction tokens — projector.py:391 -> :193-224."""
    dev = _need_cuda(q, k, v)
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    G, Mq, d = q.shape
    if k.shape != v.shape or k.shape[0] != G or k.shape[2] != d:
        raise ValueError("guide_attend: shape mismatch")
    out = torch.empty_like(q)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_guide_attend(_ptr(q), _ptr(k), _ptr(v), _ptr(out), G, Mq, k.shape[1], d, heads,
                                             float(scale), _dt(q), _stream(dev))
    _cabi.check(rc, "hicom_guide_attend")
    return out


# ------------------------------------------------------------------------------------------
# global compressor
# ------------------------------------------------------------------------------------------
def _impl_global_fold_query(q: Tensor, Wk: Tensor, heads: int, alpha: float) -> Tensor:
    """qfold[b,h*Q+i,:] = alpha * q[b,i,head h] · Wk[head h rows, :] — folds k_proj (:181) and the scale (:197)."""
    dev = _need_cuda(q, Wk)
    q, Wk = q.contiguous(), Wk.contiguous()
    B, Q, d = q.shape
    out = torch.empty((B, heads * Q, d), dtype=q.dtype, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_global_fold_query(_ptr(q), _ptr(Wk), _ptr(out), B, Q, d, heads, float(alpha),
                                                  _dt(q), _stream(dev))
    _cabi.check(rc, "hicom_global_fold_query")
    return out


def _impl_global_attend_partial(X: Tensor, pos_t: Tensor, pos_h: Tensor, pos_w: Tensor, qfold: Tensor, splits: int,
                          impl: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Split-softmax partials (m,l,o) of the global attention over X's frames — projector.py:636-640,197,213,215."""
    dev = _need_cuda(X, pos_t, pos_h, pos_w, qfold)
    X, qfold = X.contiguous(), qfold.contiguous()
    B, T, H, W, d = X.shape
    J = qfold.shape[1]
    for name, tab, n in (("pos_t", pos_t, T), ("pos_h", pos_h, H), ("pos_w", pos_w, W)):
        if tab.dtype != torch.float32 or tab.shape != (n, d) or not tab.is_contiguous():
            raise ValueError(f"global_attend_partial: {name} must be contiguous fp32 ({n},{d})")
    if qfold.shape != (B, J, d) or qfold.dtype != X.dtype:
        raise ValueError("global_attend_partial: qfold must be (B,J,d) in the feature dtype")
    lib = _cabi.load()
    ws_bytes = lib.hicom_global_attend_workspace_bytes(B, T, H, W, d, J, splits, _dt(X), impl)
    ws = torch.empty((max(ws_bytes, 256),), dtype=torch.uint8, device=dev)
    m = torch.empty((B, splits, J), dtype=torch.float32, device=dev)
    l = torch.empty((B, splits, J), dtype=torch.float32, device=dev)
    o = torch.empty((B, splits, J, d), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.hicom_global_attend_partial(_ptr(X), _ptr(pos_t), _ptr(pos_h), _ptr(pos_w), _ptr(qfold), _ptr(m),
                                             _ptr(l), _ptr(o), B, T, H, W, d, J, splits, _dt(X), _ptr(ws),
                                             ws.numel(), impl, _stream(dev))
    _cabi.check(rc, "hicom_global_attend_partial")
    return m, l, o


def _impl_global_attend_partial_keys(X: Tensor, Kscore: Tensor, pos_t: Tensor, pos_h: Tensor, pos_w: Tensor,
                                     qfold: Tensor, splits: int, impl: int) -> Tuple[Tensor, Tensor, Tensor]:
    """global_attend_partial whose scores come from explicit (normalised) keys ``Kscore`` (B,T,H,W,d) while the pooled
    operand stays X + pos_embed — the use_clip_scale variant, projector.py:184-188."""
    dev = _need_cuda(X, Kscore, pos_t, pos_h, pos_w, qfold)
    X, Kscore, qfold = X.contiguous(), Kscore.contiguous(), qfold.contiguous()
    B, T, H, W, d = X.shape
    J = qfold.shape[1]
    if Kscore.shape != X.shape or Kscore.dtype != X.dtype:
        raise ValueError("global_attend_partial_keys: Kscore must match X in shape and dtype")
    for name, tab, n in (("pos_t", pos_t, T), ("pos_h", pos_h, H), ("pos_w", pos_w, W)):
        if tab.dtype != torch.float32 or tab.shape != (n, d) or not tab.is_contiguous():
            raise ValueError(f"global_attend_partial_keys: {name} must be contiguous fp32 ({n},{d})")
    if qfold.shape != (B, J, d) or qfold.dtype != X.dtype:
        raise ValueError("global_attend_partial_keys: qfold must be (B,J,d) in the feature dtype")
    lib = _cabi.load()
    ws_bytes = lib.hicom_global_attend_workspace_bytes(B, T, H, W, d, J, splits, _dt(X), impl)
    ws = torch.empty((max(ws_bytes, 256),), dtype=torch.uint8, device=dev)
    m = torch.empty((B, splits, J), dtype=torch.float32, device=dev)
    l = torch.empty((B, splits, J), dtype=torch.float32, device=dev)
    o = torch.empty((B, splits, J, d), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.hicom_global_attend_partial_keys(_ptr(X), _ptr(Kscore), _ptr(pos_t), _ptr(pos_h), _ptr(pos_w),
                                                  _ptr(qfold), _ptr(m), _ptr(l), _ptr(o), B, T, H, W, d, J, splits,
                                                  _dt(X), _ptr(ws), ws.numel(), impl, _stream(dev))
    _cabi.check(rc, "hicom_global_attend_partial_keys")
    return m, l, o


def _impl_posadd(X: Tensor, pos_t: Tensor, pos_h: Tensor, pos_w: Tensor) -> Tensor:
    """X (B,T,H,W,d) + pos_t[t] + pos_h[h] + pos_w[w] (fp32 per-axis tables) — projector.py:636-640, separable form."""
    dev = _need_cuda(X, pos_t, pos_h, pos_w)
    X = X.contiguous()
    B, T, H, W, d = X.shape
    for name, tab, n in (("pos_t", pos_t, T), ("pos_h", pos_h, H), ("pos_w", pos_w, W)):
        if tab.dtype != torch.float32 or tab.shape != (n, d) or not tab.is_contiguous():
            raise ValueError(f"posadd: {name} must be contiguous fp32 ({n},{d})")
    out = torch.empty_like(X)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_posadd(_ptr(X), _ptr(out), _ptr(pos_t), _ptr(pos_h), _ptr(pos_w), B, T, H, W, d,
                                       _dt(X), _stream(dev))
    _cabi.check(rc, "hicom_posadd")
    return out


def _impl_l2norm_rows(X: Tensor) -> Tensor:
    """Rows of X (…, d) divided by their L2 norm — projector.py:184-186, :527-529."""
    dev = _need_cuda(X)
    X = X.contiguous()
    out = torch.empty_like(X)
    d = X.shape[-1]
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_l2norm_rows(_ptr(X), _ptr(out), X.numel() // d, d, _dt(X), _stream(dev))
    _cabi.check(rc, "hicom_l2norm_rows")
    return out


def out_code(dtype: torch.dtype) -> int:
    """dtype code of a 16-bit model dtype (1 = bf16, 2 = fp16), 0 (fp32 output) otherwise — the `out_bf16` argument of
    softmax_merge (a bool in the first ABI: True == 1 == bf16)."""
    return _DT[dtype] if dtype in (torch.bfloat16, torch.float16) else 0


def _impl_softmax_merge(m: Tensor, l: Tensor, o: Tensor, out_bf16: int) -> Tensor:
    """Combine (m,l,o) partials over dim 1 (token splits and/or frame shards) -> pooled (B,J,d); ``out_bf16``: 0/False
    fp32, 1/True bf16, 2 fp16 (`out_code`)."""
    dev = _need_cuda(m, l, o)
    m, l, o = m.contiguous(), l.contiguous(), o.contiguous()
    B, P, J, d = o.shape
    if m.shape != (B, P, J) or l.shape != (B, P, J) or o.dtype != torch.float32:
        raise ValueError("softmax_merge: shape mismatch")
    odt = _CODE_DT[int(out_bf16)]
    out = torch.empty((B, J, d), dtype=odt, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_softmax_merge(_ptr(m), _ptr(l), _ptr(o), B, P, J, d, _ptr(out), _DT[odt],
                                              _stream(dev))
    _cabi.check(rc, "hicom_softmax_merge")
    return out


def _impl_softmax_merge_lse(m: Tensor, l: Tensor, o: Tensor, out_bf16: int,
                            lse_out: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """softmax_merge that also returns lse (B,J) fp32 = log sum_n exp(score) over the tokens of these partials
    (written into ``lse_out`` when given: a contiguous (B,J) fp32 view, e.g. the tail of a frame-shard message)."""
    dev = _need_cuda(m, l, o, lse_out)
    m, l, o = m.contiguous(), l.contiguous(), o.contiguous()
    B, P, J, d = o.shape
    if m.shape != (B, P, J) or l.shape != (B, P, J) or o.dtype != torch.float32:
        raise ValueError("softmax_merge_lse: shape mismatch")
    out = torch.empty((B, J, d), dtype=_CODE_DT[int(out_bf16)], device=dev)
    if lse_out is not None and (lse_out.shape != (B, J) or lse_out.dtype != torch.float32 or not lse_out.is_contiguous()):
        raise ValueError("softmax_merge_lse: lse_out must be a contiguous (B,J) fp32 tensor")
    lse = torch.empty((B, J), dtype=torch.float32, device=dev) if lse_out is None else lse_out
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_softmax_merge_lse(_ptr(m), _ptr(l), _ptr(o), B, P, J, d, _ptr(out), _dt(out),
                                                  _ptr(lse), _stream(dev))
    _cabi.check(rc, "hicom_softmax_merge_lse")
    return out, lse


def shard_message_layout(Q: int, d: int, heads: int, dtype: torch.dtype) -> Tuple[int, int]:
    """(bytes per video, byte offset of the lse block) of the frame-shard message [attn (Q,d) | lse (heads*Q) fp32]."""
    rows = Q * d * torch.empty((), dtype=dtype).element_size()
    lse_off = -(-rows // 16) * 16
    return -(-(lse_off + heads * Q * 4) // 16) * 16, lse_off


def _impl_shard_combine(msgs: Tensor, Q: int, d: int, heads: int, dtype: torch.dtype) -> Tensor:
    """msgs (R, B, nbytes) uint8, the gathered frame-shard messages -> attn (B, Q, d): the ranks' attention rows combined
    with softmax weights of their log-sum-exps (hicom_shard_combine)."""
    dev = _need_cuda(msgs)
    if msgs.dtype != torch.uint8 or msgs.dim() != 3 or not msgs.is_contiguous():
        raise ValueError("shard_combine: msgs must be a contiguous (R, B, nbytes) uint8 tensor")
    R, B, nbytes = msgs.shape
    need, lse_off = shard_message_layout(Q, d, heads, dtype)
    if nbytes != need:
        raise ValueError(f"shard_combine: message of {nbytes} bytes, expected {need}")
    out = torch.empty((B, Q, d), dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_shard_combine(_ptr(msgs), B * nbytes, nbytes, lse_off, R, B, Q, d, heads, _ptr(out),
                                              _DT[dtype], _stream(dev))
    _cabi.check(rc, "hicom_shard_combine")
    return out


def _impl_softmax_reduce(m: Tensor, l: Tensor, o: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Reduce the P partials along dim 1 to ONE un-normalised partial (m,l,o) with P=1 — done by every rank before
    the frame-shard exchange so the message does not grow with the number of token splits."""
    dev = _need_cuda(m, l, o)
    m, l, o = m.contiguous(), l.contiguous(), o.contiguous()
    B, P, J, d = o.shape
    if m.shape != (B, P, J) or l.shape != (B, P, J) or o.dtype != torch.float32:
        raise ValueError("softmax_reduce: shape mismatch")
    mo = torch.empty((B, 1, J), dtype=torch.float32, device=dev)
    lo = torch.empty((B, 1, J), dtype=torch.float32, device=dev)
    oo = torch.empty((B, 1, J, d), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_softmax_reduce(_ptr(m), _ptr(l), _ptr(o), B, P, J, d, _ptr(mo), _ptr(lo), _ptr(oo),
                                               _stream(dev))
    _cabi.check(rc, "hicom_softmax_reduce")
    return mo, lo, oo


def global_value_proj_into(pooled: Tensor, Wv: Tensor, bv: Optional[Tensor], Q: int, heads: int, out: Tensor) -> Tensor:
    """global_value_proj written into ``out`` (B, Q, d), a contiguous view (e.g. the row block of a frame-shard message)."""
    dev = _need_cuda(pooled, Wv, bv, out)
    pooled, Wv = pooled.contiguous(), Wv.contiguous()
    B, J, d = pooled.shape
    if J != Q * heads or out.shape != (B, Q, d) or out.dtype != pooled.dtype or out.stride(-1) != 1 \
            or out.stride(1) != d:
        raise ValueError("global_value_proj_into: bad shapes")
    if B > 1 and out.stride(0) != Q * d:  # the kernel writes one dense (B*Q, d) block: go through a temporary
        out.copy_(_impl_global_value_proj(pooled, Wv, bv, Q, heads))
        return out
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_global_value_proj(_ptr(pooled), _ptr(Wv), _ptr(_c(bv)), _ptr(out), B, Q, d, heads,
                                                  _dt(pooled), _stream(dev))
    _cabi.check(rc, "hicom_global_value_proj")
    return out


def _impl_global_value_proj(pooled: Tensor, Wv: Tensor, bv: Optional[Tensor], Q: int, heads: int) -> Tensor:
    """attn[b,i,head h] = Wv[head h rows] · pooled[b,h*Q+i] + bv — v_proj (:182) after pooling + head merge (:223-224)."""
    dev = _need_cuda(pooled, Wv, bv)
    pooled, Wv = pooled.contiguous(), Wv.contiguous()
    B, J, d = pooled.shape
    if J != Q * heads:
        raise ValueError("global_value_proj: J != Q*heads")
    out = torch.empty((B, Q, d), dtype=pooled.dtype, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_global_value_proj(_ptr(pooled), _ptr(Wv), _ptr(_c(bv)), _ptr(out), B, Q, d, heads,
                                                  _dt(pooled), _stream(dev))
    _cabi.check(rc, "hicom_global_value_proj")
    return out


# ------------------------------------------------------------------------------------------
# backward building blocks (composed by hicom_b200/autograd.py)
# ------------------------------------------------------------------------------------------
_GEMM_TRIPLES = {(torch.float32, torch.float32, torch.float32), (torch.float32, torch.float32, torch.bfloat16),
                 (torch.bfloat16, torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.bfloat16, torch.float32),
                 (torch.float32, torch.bfloat16, torch.float32)}


def _impl_gemm(A: Tensor, B: Tensor, out: Optional[Tensor], out_fp32: bool, alpha: float) -> Tensor:
    """C = alpha * A @ B over strided VIEWS (no copies): A (..., M, K), B (..., K, N) with up to two leading batch
    dims each (broadcast dims may have stride 0, e.g. from ``expand``), any element strides; ``out`` (..., M, N) may
    be a strided view with unit inner stride (written in place), else a new contiguous tensor in A's dtype (fp32 when
    ``out_fp32``).  The contraction behind every backward formula (dA = dY·W, dW = dYᵀ·A, per-head folds, …)."""
    dev = _need_cuda(A, B, out)
    if A.dim() < 2 or B.dim() < 2 or A.dim() > 4 or B.dim() > 4:
        raise ValueError("gemm: operands must have 2 to 4 dims")
    M, K = A.shape[-2:]
    K2, N = B.shape[-2:]
    if K != K2:
        raise ValueError(f"gemm: inner dims differ ({K} vs {K2})")
    batch = torch.broadcast_shapes(A.shape[:-2], B.shape[:-2])
    batch = (1,) * (2 - len(batch)) + tuple(batch)
    A4, B4 = A.expand(*batch, M, K), B.expand(*batch, K, N)
    if out is None:
        odt = torch.float32 if out_fp32 else A.dtype
        out = torch.empty((*batch, M, N), dtype=odt, device=dev)
        ret = out.reshape(*torch.broadcast_shapes(A.shape[:-2], B.shape[:-2]), M, N)
    else:
        ret = out
        if tuple(out.shape[-2:]) != (M, N) or (N > 1 and out.stride(-1) != 1):
            raise ValueError("gemm: out must be (..., M, N) with unit inner stride")
    C4 = out.expand(*batch, M, N) if out.dim() < 4 else out
    if tuple(C4.shape) != (*batch, M, N) or any(C4.stride(i) == 0 and batch[i] > 1 for i in (0, 1)):
        raise ValueError("gemm: out batch dims do not match the operands")
    if (A.dtype, B.dtype, out.dtype) not in _GEMM_TRIPLES:
        raise TypeError(f"gemm: dtype combination {A.dtype}, {B.dtype} -> {out.dtype} is not built")
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_gemm(_ptr(A4), A4.stride(2), A4.stride(3), A4.stride(0), A4.stride(1),
                                     _ptr(B4), B4.stride(2), B4.stride(3), B4.stride(0), B4.stride(1),
                                     _ptr(C4), C4.stride(2) if M > 1 else max(C4.stride(2), N), C4.stride(0),
                                     C4.stride(1), M, N, K, batch[0], batch[1], float(alpha), _dt(A), _dt(B),
                                     _dt(out), _stream(dev))
    _cabi.check(rc, "hicom_gemm")
    return ret


def colsum_supported(x2: Tensor) -> bool:
    return (x2.dim() == 2 and x2.shape[1] % 4 == 0 and x2.stride(1) == 1 and x2.stride(0) % 4 == 0
            and x2.stride(0) >= x2.shape[1] and x2.data_ptr() % 16 == 0 and x2.dtype in _DT)


def _impl_colsum(x2: Tensor) -> Tensor:
    """Column sums of (M, N) -> fp32 (N) — bias gradients (db = 1ᵀ·dpre)."""
    dev = _need_cuda(x2)
    if not colsum_supported(x2):
        raise ValueError("colsum: needs a 2-D row-major view with N and the row pitch multiples of 4, 16-byte aligned")
    M, N = x2.shape
    out = torch.zeros((N,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_colsum(_ptr(x2), x2.stride(0), _ptr(out), M, N, _dt(x2), _stream(dev))
    _cabi.check(rc, "hicom_colsum")
    return out


def _impl_act_backward(pre: Tensor, dy: Tensor, act: int) -> Tensor:
    """dy * act'(pre) — backward of the GELU between the layers of build_mlp (projector.py:310)."""
    dev = _need_cuda(pre, dy)
    pre, dy = pre.contiguous(), dy.contiguous()
    if pre.shape != dy.shape:
        raise ValueError("act_backward: shape mismatch")
    out = torch.empty_like(dy)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_act_backward(_ptr(pre), _ptr(dy), _ptr(out), dy.numel(), act, _dt(pre), _dt(dy),
                                             _stream(dev))
    _cabi.check(rc, "hicom_act_backward")
    return out


def col_logsumexp(S: Tensor) -> Tensor:
    """log sum_n exp(S[b, n, j]) of fp32 scores (B,N,J) -> (B,J), S untouched (hicom_col_stats over token ranges, then
    a tiny combine)."""
    dev = _need_cuda(S)
    if S.dtype != torch.float32 or S.dim() != 3 or not S.is_contiguous():
        raise ValueError("col_logsumexp: S must be a contiguous fp32 (B,N,J) tensor")
    B, N, J = S.shape
    splits = max(1, min(64, N // 256))
    m = torch.empty((B, splits, J), dtype=torch.float32, device=dev)
    l = torch.empty_like(m)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_col_stats(_ptr(S), _ptr(m), _ptr(l), B, N, J, splits, _stream(dev))
    _cabi.check(rc, "hicom_col_stats")
    return torch.logsumexp(m + torch.log(l), dim=1)


def _impl_softmax_backward(S: Tensor, dP: Tensor, lse: Tensor, delta: Tensor, out_bf16: bool) -> Tensor:
    """dS = exp(S - lse) * (dP - delta): S, dP (B,N,J) fp32, lse/delta (B,J) fp32 — softmax of projector.py:213."""
    dev = _need_cuda(S, dP, lse, delta)
    S, dP, lse, delta = S.contiguous(), dP.contiguous(), lse.contiguous(), delta.contiguous()
    B, N, J = S.shape
    if dP.shape != S.shape or lse.shape != (B, J) or delta.shape != (B, J):
        raise ValueError("softmax_backward: shape mismatch")
    if any(t.dtype != torch.float32 for t in (S, dP, lse, delta)):
        raise TypeError("softmax_backward: fp32 inputs expected")
    odt = torch.bfloat16 if out_bf16 else torch.float32
    out = torch.empty((B, N, J), dtype=odt, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_softmax_backward(_ptr(S), _ptr(dP), _ptr(lse), _ptr(delta), _ptr(out), B, N, J,
                                                 _DT[odt], _stream(dev))
    _cabi.check(rc, "hicom_softmax_backward")
    return out


def _impl_local_attend_backward(K: Tensor, V: Tensor, Q: Tensor, dO: Tensor, kt: int, ks: int, logit_scale: float,
                                k_l2norm: bool, need_q: bool, need_k: bool, need_v: bool):
    """Backward of local_attend (projector.py:546-553): K, V (B,T,H,W,d); Q, dO (B,Nw,d).  Returns
    (dQ in Q's dtype | None, dK fp32 (B,T,H,W,d) | None, dV fp32 | None)."""
    dev = _need_cuda(K, V, Q, dO)
    V = V.contiguous()
    K = V if K.data_ptr() == V.data_ptr() else K.contiguous()
    Q, dO = Q.contiguous(), dO.contiguous()
    B, T, H, W, d = V.shape
    nw = num_windows(T, H, W, kt, ks)
    if K.shape != V.shape or Q.shape != (B, nw, d) or dO.shape != (B, nw, d):
        raise ValueError("local_attend_backward: shape mismatch")
    if not (K.dtype == V.dtype == Q.dtype == dO.dtype):
        raise TypeError("local_attend_backward: dtypes differ")
    dQ = torch.empty_like(Q) if need_q else None
    dK = torch.zeros(V.shape, dtype=torch.float32, device=dev) if need_k else None
    dV = torch.zeros(V.shape, dtype=torch.float32, device=dev) if need_v else None
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_local_attend_backward(_ptr(K), _ptr(V), _ptr(Q), _ptr(dO), _ptr(dQ), _ptr(dK),
                                                      _ptr(dV), B, T, H, W, d, kt, ks, float(logit_scale),
                                                      int(k_l2norm), _dt(V), _stream(dev))
    _cabi.check(rc, "hicom_local_attend_backward")
    return dQ, dK, dV


def _impl_grid_pool_backward(dQ: Tensor, T: int, H: int, W: int, kt: int, ks: int) -> Tensor:
    """Backward of grid_pool: dQ (B,Nw,d) -> dX (B,T,H,W,d) fp32 (gradients into frames_feature)."""
    dev = _need_cuda(dQ)
    dQ = dQ.contiguous()
    B, nw, d = dQ.shape
    if nw != num_windows(T, H, W, kt, ks):
        raise ValueError("grid_pool_backward: dQ does not match the window grid")
    dX = torch.zeros((B, T, H, W, d), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_grid_pool_backward(_ptr(dQ), _ptr(dX), B, T, H, W, d, kt, ks, _dt(dQ), _stream(dev))
    _cabi.check(rc, "hicom_grid_pool_backward")
    return dX


def _impl_l2norm_rows_backward(X: Tensor, dY: Tensor) -> Tensor:
    """Backward of l2norm_rows: dX = (dY - Y (Y·dY)) / |X| per row of the last dim."""
    dev = _need_cuda(X, dY)
    X, dY = X.contiguous(), dY.contiguous()
    if dY.shape != X.shape or dY.dtype != X.dtype:
        raise ValueError("l2norm_rows_backward: operands differ")
    d = X.shape[-1]
    dX = torch.empty_like(X)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_l2norm_rows_backward(_ptr(X), _ptr(dY), _ptr(dX), X.numel() // d, d, _dt(X),
                                                     _stream(dev))
    _cabi.check(rc, "hicom_l2norm_rows_backward")
    return dX


def _impl_film_layernorm_backward(x: Tensor, film: Tensor, ln_w: Tensor, dy: Tensor, rows_per_group: int,
                                  need_dx: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Backward of film_layernorm: returns (dx | empty, dfilm (G,2d) fp32, dw (d) fp32, dbias (d) fp32)."""
    dev = _need_cuda(x, film, ln_w, dy)
    x, film, dy = x.contiguous(), film.contiguous(), dy.contiguous()
    d = x.shape[-1]
    rows = x.numel() // d
    if dy.shape != x.shape or dy.dtype != x.dtype or ln_w.dtype != x.dtype:
        raise ValueError("film_layernorm_backward: operands differ")
    if film.dtype != torch.float32 or film.shape[-1] != 2 * d or rows > film.shape[0] * rows_per_group:
        raise ValueError("film_layernorm_backward: film must be fp32 (G, 2d) covering every row")
    dx = torch.empty_like(x) if need_dx else torch.empty((0,), dtype=x.dtype, device=dev)
    dfilm = torch.zeros_like(film)
    dw = torch.zeros((d,), dtype=torch.float32, device=dev)
    db = torch.zeros((d,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_film_layernorm_backward(_ptr(x), _ptr(film), _ptr(_c(ln_w)), _ptr(dy),
                                                        _ptr(dx) if need_dx else None, _ptr(dfilm), _ptr(dw),
                                                        _ptr(db), rows, d, rows_per_group, _dt(x), _stream(dev))
    _cabi.check(rc, "hicom_film_layernorm_backward")
    return dx, dfilm, dw, db


def _impl_mix_layernorm_backward(x: Tensor, y: Tensor, ln_w: Tensor, ln_b: Tensor, alpha: Tensor, dout: Tensor,
                                 need_dx: bool):
    """Backward of mix_layernorm: returns (dx | None, dy, dw fp32 (d), dbias fp32 (d), dalpha fp32 (1))."""
    dev = _need_cuda(x, y, ln_w, ln_b, alpha, dout)
    x, y, dout = x.contiguous(), y.contiguous(), dout.contiguous()
    d = x.shape[-1]
    if y.shape != x.shape or dout.shape != x.shape or not (x.dtype == y.dtype == dout.dtype == alpha.dtype == ln_w.dtype):
        raise ValueError("mix_layernorm_backward: operands differ")
    dx = torch.empty_like(x) if need_dx else None
    dy = torch.empty_like(y)
    dw = torch.zeros((d,), dtype=torch.float32, device=dev)
    db = torch.zeros((d,), dtype=torch.float32, device=dev)
    da = torch.zeros((1,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.load().hicom_mix_layernorm_backward(_ptr(x), _ptr(y), _ptr(_c(ln_w)), _ptr(_c(ln_b)),
                                                       _ptr(alpha.contiguous()), _ptr(dout), _ptr(dx), _ptr(dy),
                                                       _ptr(dw), _ptr(db), _ptr(da), x.numel() // d, d, _dt(x),
                                                       _stream(dev))
    _cabi.check(rc, "hicom_mix_layernorm_backward")
    return dx, dy, dw, db, da


def set_sm_limit(sms: int) -> int:
    """Kernels enqueued by this thread from now on size their persistent grids for at most ``sms`` SMs (0 = whole
    device).  Returns the previous limit.  Used to run the HBM-bound local chain and the tensor-bound global chain side
    by side on disjoint SM sets (``projector.SM_SPLIT``)."""
    return int(_cabi.load().hicom_set_sm_limit(int(sms)))


class sm_limit:
    """``with ops.sm_limit(n): ...`` — scoped ``set_sm_limit``."""

    def __init__(self, sms: int):
        self.sms = int(sms)

    def __enter__(self):
        self.old = set_sm_limit(self.sms)
        return self

    def __exit__(self, *exc):
        set_sm_limit(self.old)


def kernel_launch_count() -> int:
    """Kernels enqueued by libhicom_b200 since load (bench.py's ``gpu_launches``)."""
    return int(_cabi.load().hicom_kernel_launch_count())


class KernelTimer:
    """Per-kernel CUDA-event timing inside libhicom_b200 (tcgen05 GEMMs and the local kernel).
    ``with KernelTimer() as kt: ...; kt.summary()`` -> {label: (launches, total_ms)}."""

    def __enter__(self):
        _cabi.load().hicom_kernel_timing_enable(1)
        return self

    def __exit__(self, *exc):
        lib = _cabi.load()
        n = lib.hicom_kernel_timing_collect(None, 0)
        buf = ctypes.create_string_buffer(int(n) + 16)
        lib.hicom_kernel_timing_collect(buf, len(buf))
        self._table = {}
        for line in buf.value.decode().splitlines():
            label, count, ms = line.split("\t")
            self._table[label] = (int(count), float(ms))
        lib.hicom_kernel_timing_enable(0)

    def summary(self):
        return dict(self._table)


class OpTimer:
    """Times every hicom_b200 op with CUDA events on the launching stream (no synchronisation while
    recording).  ``with OpTimer() as t: ...; t.summary()`` -> {op: (calls, total_ms)}."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        global _ACTIVE_TIMER
        _ACTIVE_TIMER = self
        return self

    def __exit__(self, *exc):
        global _ACTIVE_TIMER
        _ACTIVE_TIMER = None

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.records:
            calls, ms = out.get(name, (0, 0.0))
            out[name] = (calls + 1, ms + a.elapsed_time(b))
        return out


_ACTIVE_TIMER = None


def timed(name: str):
    """Context manager used by the modules around each op call; free when no OpTimer is active."""
    return _Timed(name)


class _Timed:
    __slots__ = ("name", "start")

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _ACTIVE_TIMER is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()
        return self

    def __exit__(self, *exc):
        if _ACTIVE_TIMER is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            _ACTIVE_TIMER.records.append((self.name, self.start, end))


_SM_COUNT = {}


def sm_count(device=None) -> int:
    """SMs of ``device`` (cached)."""
    key = str(device)
    if key not in _SM_COUNT:
        _SM_COUNT[key] = device_info(device)[0]
    return _SM_COUNT[key]


def device_info(device=None):
    """(sm_count, cc_major, cc_minor) of the current CUDA device; raises unless it is sm_100."""
    sm, ma, mi = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    with torch.cuda.device(device):
        rc = _cabi.load().hicom_device_info(ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi))
    _cabi.check(rc, "hicom_device_info")
    return sm.value, ma.value, mi.value


# ------------------------------------------------------------------------------------------
# registration + dispatch
#
# Every op is registered as ``torch.ops.hicom_b200.<name>`` (a forward-only torch custom op).  The
# torch.library dispatcher costs ~45 us of host time per call on this stack, which is more than most of
# these kernels take, so by default the modules call the SAME Python implementation directly; set
# HICOM_VIA_TORCH_OPS=1 to route every call through the dispatcher instead (tests do both).
# ------------------------------------------------------------------------------------------
import os as _os

VIA_TORCH_OPS = _os.environ.get("HICOM_VIA_TORCH_OPS", "0") == "1"
_REGISTERED = {}


def _register(name, fn, mutates):
    op = torch.library.custom_op(f"hicom_b200::{name}", mutates_args=mutates, device_types="cuda")(fn)
    _REGISTERED[name] = op
    return op


def _wrap(name, impl, mutates, label):
    op = _register(name, impl, mutates)

    def call(*args):
        target = op if VIA_TORCH_OPS else impl
        if _ACTIVE_TIMER is None:
            return target(*args)
        with _Timed(label(*args)):
            return target(*args)

    call.__name__ = name
    call.__doc__ = impl.__doc__
    call.op, call.impl = op, impl
    return call


def _rows(t):
    return t.numel() // t.shape[-1]


_lin_label = lambda A, W, *a: f"linear M={_rows(A)} N={W.shape[0]} K={W.shape[1]}"
grid_pool = _wrap("grid_pool", _impl_grid_pool, (), lambda *a: "grid_pool")
local_attend = _wrap("local_attend", _impl_local_attend, (), lambda *a: "local_attend")
local_attend_into = _wrap("local_attend_into", _impl_local_attend_into, ("out",), lambda *a: "local_attend")
linear = _wrap("linear", _impl_linear, (), _lin_label)
linear_into = _wrap("linear_into", _impl_linear_into, ("out",), _lin_label)
layernorm = _wrap("layernorm", _impl_layernorm, (), lambda *a: "layernorm")
film_layernorm = _wrap("film_layernorm", _impl_film_layernorm, (), lambda *a: "film_layernorm")
add_layernorm = _wrap("add_layernorm", _impl_add_layernorm, (), lambda *a: "add_layernorm")
mix_layernorm = _wrap("mix_layernorm", _impl_mix_layernorm, (), lambda *a: "mix_layernorm")
guide_attend = _wrap("guide_attend", _impl_guide_attend, (), lambda *a: "guide_attend")
global_fold_query = _wrap("global_fold_query", _impl_global_fold_query, (), lambda *a: "global_fold_query")
global_attend_partial = _wrap("global_attend_partial", _impl_global_attend_partial, (),
                              lambda *a: "global_attend_partial")
global_attend_partial_keys = _wrap("global_attend_partial_keys", _impl_global_attend_partial_keys, (),
                                   lambda *a: "global_attend_partial_keys")
posadd = _wrap("posadd", _impl_posadd, (), lambda *a: "posadd")
l2norm_rows = _wrap("l2norm_rows", _impl_l2norm_rows, (), lambda *a: "l2norm_rows")
softmax_merge = _wrap("softmax_merge", _impl_softmax_merge, (), lambda *a: "softmax_merge")
softmax_reduce = _wrap("softmax_reduce", _impl_softmax_reduce, (), lambda *a: "softmax_reduce")
softmax_merge_lse = _impl_softmax_merge_lse  # optional output view: called directly
shard_combine = _impl_shard_combine  # takes a torch.dtype: called directly
global_value_proj = _wrap("global_value_proj", _impl_global_value_proj, (), lambda *a: "global_value_proj")
# backward blocks are called directly (``gemm`` writes into strided views, which torch.library cannot describe)
gemm, act_backward, softmax_backward, colsum = _impl_gemm, _impl_act_backward, _impl_softmax_backward, _impl_colsum
local_attend_backward = _impl_local_attend_backward
grid_pool_backward, l2norm_rows_backward = _impl_grid_pool_backward, _impl_l2norm_rows_backward
film_layernorm_backward = _impl_film_layernorm_backward
mix_layernorm_backward = _impl_mix_layernorm_backward
