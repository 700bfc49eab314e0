/*
 * hicom_b200 — C ABI of the B200-native HICom compressor kernels (sm_100a).
 *
 * The reference (lntzm/HICom) is 100 % Python and defines no FFI; its hot path is
 * `HIComProjector.forward` (hicom/model/projector.py:676-708).  This header is the seam a
 * maintainer binds instead of the ATen calls on that path.  Every entry point names the reference
 * lines it replaces.  All pointers are DEVICE pointers (plain `void*`, no torch types), tensors are
 * dense row-major unless a leading dimension is given, `stream` is a `cudaStream_t` passed as
 * `void*`; nothing here synchronises the device.  Functions return 0 on success and a non-zero
 * code otherwise — `hicom_last_error()` then holds a message (thread-local).  Nothing aborts.
 *
 * dtype codes: HICOM_F32 tensors are float, HICOM_BF16 tensors are __nv_bfloat16, HICOM_F16 tensors are __half
 * (the reference's inference dtype, hicom/model/__init__.py:44; forward entry points only — the backward blocks take
 * F32 / BF16); accumulation is always fp32, softmax statistics are fp32.
 */
#ifndef HICOM_B200_H
#define HICOM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HICOM_ABI_VERSION 1

enum { HICOM_F32 = 0, HICOM_BF16 = 1, HICOM_F16 = 2 };
/* GELU = exact erf form (nn.GELU(), projector.py:310); GELU_TANH = the tanh form ("gelu_pytorch_tanh") of the SigLIP
 * pooling-head MLP that produces frames_embed (encoder.py:284-285) */
enum { HICOM_ACT_NONE = 0, HICOM_ACT_GELU = 1, HICOM_ACT_GELU_TANH = 2 };
enum {
  HICOM_Q_POOLED = 0,   /* query = grid-pooled feature                 (use_guide None/off)        */
  HICOM_Q_FILM_LN = 1,  /* query = LN(pooled*(1+scale)+shift)           (coarse, projector.py:369-372) */
  HICOM_Q_VECTOR = 2,   /* query = per-video vector, pooled discarded  (direct, projector.py:367-368) */
  HICOM_Q_EXPLICIT = 3  /* query rows supplied by the caller           (fine / adapt_q paths)       */
};
enum { HICOM_IMPL_AUTO = 0, HICOM_IMPL_SIMT = 1, HICOM_IMPL_TCGEN05 = 2 };

/* ---- library ---------------------------------------------------------------------------- */
int hicom_abi_version(void);
const char* hicom_last_error(void);
/* Number of CUDA kernels this library has enqueued since load (monotonic; for launch accounting). */
uint64_t hicom_kernel_launch_count(void);
/* Per-kernel timing for measurement (off by default): when enabled, the dominant kernels are bracketed by CUDA
 * events on their launching stream; collect() synchronises the device and writes "label\tcount\ttotal_ms" lines. */
int hicom_kernel_timing_enable(int on);
size_t hicom_kernel_timing_collect(char* buf, size_t cap);
/* Fills SM count / compute capability of the current device; fails if it is not sm_100. */
int hicom_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* SM partitioning: kernels enqueued by the calling thread after this call size their persistent grids for at most
 * `sms` SMs (rounded down to whole SM pairs; 0 = the whole device).  The compressor's HBM-bound chain (local window
 * attention) and its tensor-bound chain (global scores / pooling GEMMs) have no reference counterpart to cite: the
 * reference runs them back to back (projector.py:691,700); here they run side by side on disjoint SM sets of one GPU.
 * Returns the previous limit. */
int hicom_set_sm_limit(int sms);

/* ---- local compressor -------------------------------------------------------------------
 * hicom_grid_pool: the trilinear grid pooling `F.interpolate(..., mode='trilinear')`
 *   (projector.py:536-540).  X (B,T,H,W,d) -> Q (B, Nw, d), Nw = ceil(T/kt)*ceil(H/ks)*ceil(W/ks),
 *   align_corners=False taps, rows in (t1,h1,w1) raster order.  Needed on its own only when the
 *   pooled query goes through GEMMs before the attention (adapt_q, fine).
 */
int hicom_grid_pool(const void* X, void* Q, int B, int T, int H, int W, int d, int kt, int ks,
                    int dtype, void* stream);

/* hicom_local_attend: fused grid-pool -> instruction injection -> window gather -> softmax -> A·V
 *   (projector.py:536-558 with GuideInjector.forward_direct_and_coarse :352-372 and
 *   divide_feature/balance_divide_feature :473-522).
 *   Ksrc  (B,T,H,W,d)  keys   (frames_embed, or frames_feature when that is None, :532-533)
 *   Vsrc  (B,T,H,W,d)  values (frames_feature, :534)
 *   Psrc  (B,T,H,W,d)  pooling source (raw frames_feature, :539); may alias Vsrc
 *   qmode HICOM_Q_*; q_aux: (B,d) vector for Q_VECTOR, (B,Nw,d) rows for Q_EXPLICIT, else NULL
 *   film  (B,2d) fp32 [scale | shift] for Q_FILM_LN (the coarse_proj MLP output, :370-371)
 *   ln_w, ln_b (d) in `dtype` for Q_FILM_LN (coarse_norm, eps 1e-6)
 *   score = (q·k) * logit_scale   (1/sqrt(qk_dim) :551, or exp(logit_scale) :549; the additive
 *   logit_bias is constant over a window and cancels in the softmax).  k_l2norm != 0 divides each
 *   key by its L2 norm first (:528).  Windows follow the reference's balanced overlapping rule
 *   when a dimension is not divisible; kt/ks are the EFFECTIVE kernel (caller passes kt=1 for
 *   images / T==1, :536).  Shapes the reference cannot stack (e.g. T in {5,6,9}, kt=4) fail.
 *   out (B,Nw,d) in `dtype`.
 */
int hicom_local_attend(const void* Ksrc, const void* Vsrc, const void* Psrc, const void* q_aux,
                       const float* film, const void* ln_w, const void* ln_b, void* out, int B,
                       int T, int H, int W, int d, int kt, int ks, int qmode, float logit_scale,
                       int k_l2norm, int dtype, void* stream);

/* ---- dense layers -----------------------------------------------------------------------
 * hicom_linear: C = act(A · Wᵀ + bias) [+ R]   — nn.Linear / build_mlp stages
 *   (projector.py:180-182,226,307-312,559,646).
 *   A (M,K) lda; W (N,K) ldw (torch Linear layout); bias (N) or NULL; R (M,N) ldr residual or NULL
 *   (added after the activation); C (·,N) ldc.  Output row r is written to row
 *   (r / rows_per_group) * group_stride_rows + (r % rows_per_group) of C — this is how the local
 *   and global readouts write straight into the concatenated (Nw+Q, Dh) token block of each video
 *   (projector.py:707); pass rows_per_group = M, group_stride_rows = 0 for a plain GEMM.
 *   in_dtype covers A, W, bias, R; out_dtype covers C.  impl: HICOM_IMPL_*.
 */
int hicom_linear(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias,
                 const void* R, int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int act,
                 int in_dtype, int out_dtype, int rows_per_group, int64_t group_stride_rows,
                 int impl, void* stream);

/* Row-wise LayerNorm family over rows of length d (eps 1e-6), all tensors in `dtype`:
 *   film_layernorm : out = LN(x*(1+scale)+shift)     coarse injector on explicit rows (:369-372);
 *                    film (G,2d) fp32, row r uses film[r / rows_per_group]
 *   add_layernorm  : out = LN(a + b)                 fine injector residual (:392)
 *   mix_layernorm  : out = (1-alpha)*x + alpha*LN(y) adapters (:365,533-534,541); alpha read from device
 *   layernorm      : out = LN(x)                     producer side: the SigLIP head's layernorm in front of the MLP
 *                    that makes frames_embed (encoder.py:284; eps 1e-6 = SigLIP's layer_norm_eps)
 */
int hicom_layernorm(const void* x, const void* ln_w, const void* ln_b, void* out, int64_t rows, int d, int dtype,
                    void* stream);
int hicom_film_layernorm(const void* x, const float* film, const void* ln_w, const void* ln_b,
                         void* out, int rows, int d, int rows_per_group, int dtype, void* stream);
int hicom_add_layernorm(const void* a, const void* b, const void* ln_w, const void* ln_b, void* out,
                        int rows, int d, int dtype, void* stream);
int hicom_mix_layernorm(const void* x, const void* y, const void* ln_w, const void* ln_b,
                        const void* alpha, void* out, int64_t rows, int d, int dtype, void* stream);

/* hicom_guide_attend: multi-head attention of query rows over a SHORT key/value list — the `fine`
 *   injector's MHA over the L instruction tokens (projector.py:391 -> :193-224).  q (G,Mq,d),
 *   k,v (G,L,d) already projected; out (G,Mq,d); heads = d/128; score = q·k*scale; fp32 softmax.
 */
int hicom_guide_attend(const void* q, const void* k, const void* v, void* out, int G, int Mq, int L,
                       int d, int heads, float scale, int dtype, void* stream);

/* ---- global compressor ------------------------------------------------------------------
 * The global cross-attention (projector.py:634-646 -> MultiheadAttention.forward :166-228) is
 * evaluated in the reassociated form: with q = Wq·Qg+bq,
 *     S[n,(h,i)] = scale * q[i,h]·(Wk_h x'_n + bk_h)  =  x'_n · qfold[(h,i)]  + const(h,i)
 *     O[i,h]     = Wv_h (sum_n softmax_n(S)[n,(h,i)] x'_n) + bv_h
 * where x' = x + pos_embed.  The constant drops out of the softmax.  J = heads*Q columns.
 *
 * hicom_global_fold_query: qfold[b,h*Q+i,:] = alpha * sum_c q[b,i,h*hd+c] * Wk[h*hd+c,:]
 *   (replaces k_proj :181 together with the scale of :197).  q (B,Q,d), Wk (d,d), qfold (B,J,d).
 */
int hicom_global_fold_query(const void* q, const void* Wk, void* qfold, int B, int Q, int d,
                            int heads, float alpha, int dtype, void* stream);

/* hicom_global_attend_partial: split-softmax partials of the global attention over a block of
 *   frames (:636-640 pos-embed add, :197 scores, :213 softmax, :215 P·V).
 *   X (B,T,H,W,d); pos_t (T,d), pos_h (H,d), pos_w (W,d) fp32 per-axis sincos tables
 *   (projector.py:57-101 is separable: PE[t,h,w] = f(t)+f(h)+f(w); the caller builds the tables in
 *   float64 like the reference, rounds each once, and offsets pos_t to the shard's first global
 *   frame); qfold (B,J,d).
 *   Tokens of each video are cut into `splits` contiguous ranges; for range s and column j:
 *     m[b,s,j] = max_n S,   l[b,s,j] = sum_n exp(S-m),   o[b,s,j,:] = sum_n exp(S-m) x'_n  (fp32)
 *   workspace: at least hicom_global_attend_workspace_bytes(...) bytes, 256-byte aligned.
 */
size_t hicom_global_attend_workspace_bytes(int B, int T, int H, int W, int d, int J, int splits,
                                           int dtype, int impl);
int hicom_global_attend_partial(const void* X, const float* pos_t, const float* pos_h,
                                const float* pos_w, const void* qfold, float* m, float* l, float* o,
                                int B, int T, int H, int W, int d, int J, int splits, int dtype,
                                void* workspace, size_t workspace_bytes, int impl, void* stream);

/* hicom_global_attend_partial_keys: the clip-scale variant (use_clip_scale has 'global', projector.py:184-188):
 *   q and k are L2-normalised over all d channels BEFORE the head split and the logits are scaled by
 *   exp(logit_scale) (+ logit_bias, which cancels in the softmax).  The per-token key norm cannot be folded into the
 *   queries, so the caller supplies explicit normalised keys
 *     Kscore (B,T,H,W,d) = l2norm_rows(Wk·x' + bk)   [hicom_linear, hicom_posadd with Wk-transformed tables,
 *                                                      hicom_l2norm_rows]
 *   and qfold (B,J,d) = exp(logit_scale) * blockdiag_h(l2norm(q)); the scores are Kscore·qfoldᵀ, while the pooled
 *   operand remains x' = X + pos_embed, so the value side stays reassociated (v_proj after pooling).
 *   Same outputs and workspace contract as hicom_global_attend_partial.
 */
int hicom_global_attend_partial_keys(const void* X, const void* Kscore, const float* pos_t,
                                     const float* pos_h, const float* pos_w, const void* qfold, float* m,
                                     float* l, float* o, int B, int T, int H, int W, int d, int J, int splits,
                                     int dtype, void* workspace, size_t workspace_bytes, int impl, void* stream);

/* hicom_posadd: Y[b,t,h,w,:] = X[b,t,h,w,:] + pos_t[t] + pos_h[h] + pos_w[w]  (the separable form of the
 *   position-embedding add, projector.py:636-640; fp32 tables, X/Y in dtype, in place allowed). */
int hicom_posadd(const void* X, void* Y, const float* pos_t, const float* pos_h, const float* pos_w, int B,
                 int T, int H, int W, int d, int dtype, void* stream);

/* hicom_l2norm_rows: Y[r,:] = X[r,:] / ||X[r,:]||_2  (projector.py:184-186, :527-529); in place allowed. */
int hicom_l2norm_rows(const void* X, void* Y, long long rows, int d, int dtype, void* stream);

/* hicom_softmax_merge: combine P partials per (video, column) — across token splits and, after an
 *   all-gather, across frame shards on other GPUs:
 *     M = max_p m_p;  L = sum_p l_p e^{m_p-M};  pooled = sum_p o_p e^{m_p-M} / L
 *   m,l (B,P,J) o (B,P,J,d) fp32; pooled (B,J,d) in out_dtype.
 */
int hicom_softmax_merge(const float* m, const float* l, const float* o, int B, int P, int J, int d,
                        void* pooled, int out_dtype, void* stream);

/* hicom_softmax_reduce: same combination but WITHOUT the normalisation — reduces a rank's P token-split partials
 *   to ONE partial (m_out,l_out (B,J), o_out (B,J,d), fp32) before it is exchanged, so the NVLink message is
 *   J*(d+2) floats per video regardless of how many splits the kernels used.
 */
int hicom_softmax_reduce(const float* m, const float* l, const float* o, int B, int P, int J, int d,
                         float* m_out, float* l_out, float* o_out, void* stream);

/* hicom_softmax_merge_lse: hicom_softmax_merge that also returns lse (B,J) = M + log L, the log-sum-exp of every
 *   (video, column) over the tokens this call saw — what a frame shard sends to its peers beside its attention rows.
 */
int hicom_softmax_merge_lse(const float* m, const float* l, const float* o, int B, int P, int J, int d,
                            void* pooled, int out_dtype, float* lse, void* stream);

/* hicom_shard_combine: frame shards of one long video.  Every rank r contributes one message per video:
 *   [ attn_r (Q, d) in `dtype` | ... | lse_r (heads*Q) fp32 at byte lse_offset ], its own normalised attention output
 *   AFTER the per-head value projection (linear, so it commutes with the softmax merge; projector.py:215,223-224) and
 *   the log-sum-exp of its scores.  `msgs` holds the R gathered messages: rank r, video b at
 *   msgs + r*rank_stride + b*video_stride.   out[b,i,h*hd+c] = sum_r softmax_r(lse_r[b,h*Q+i]) * attn_r[b,i,h*hd+c].
 *   The message is Q*d 16-bit values + heads*Q floats (75 KB for 32 x 1152) instead of the J*(d+2) fp32 partial (1.3 MB).
 */
int hicom_shard_combine(const void* msgs, long long rank_stride_bytes, long long video_stride_bytes,
                        long long lse_offset_bytes, int R, int B, int Q, int d, int heads, void* out, int dtype,
                        void* stream);

/* hicom_global_value_proj: attn[b,i,h*hd+c] = sum_k Wv[h*hd+c,k] * pooled[b,h*Q+i,k] + bv[h*hd+c]
 *   (v_proj :182 applied after the pooling, and the head merge :223-224).
 */
int hicom_global_value_proj(const void* pooled, const void* Wv, const void* bv, void* attn, int B,
                            int Q, int d, int heads, int dtype, void* stream);

/* ---- backward building blocks --------------------------------------------
 * The reference trains the projector in all three stages (train.py:704-738) through PyTorch autograd; these entry
 * points are what `hicom_b200/autograd.py` composes into the backward of each forward op above.  The SigLIP tower and
 * the guide encoder are frozen (encoder.py:235,247), so no gradient flows to X, E or the instruction embedding.
 *
 * hicom_gemm: C[b1,b2] = alpha * A[b1,b2] · B[b1,b2] with arbitrary ELEMENT strides — A[m,k] at m*sAm + k*sAk,
 *   B[k,n] at k*sBk + n*sBn, C rows ldc apart, unit inner stride — and two batch levels (strides may be 0 to
 *   broadcast).  Every contraction of the backward formulas: dA = dY·W, dW = dYᵀ·A, db = 1ᵀ·dY (nn.Linear,
 *   projector.py:307-312), dP = X'·dpooledᵀ and dqfold = dSᵀ·X' (global attention, :197-215), the per-head products
 *   of the query fold and the value projection.  fp32 accumulation.  Built dtype triples (A,B,C): (f32,f32,f32),
 *   (f32,f32,bf16), (bf16,bf16,bf16), (bf16,bf16,f32), (f32,bf16,f32).
 */
int hicom_gemm(const void* A, int64_t sAm, int64_t sAk, int64_t sAb1, int64_t sAb2, const void* B, int64_t sBk,
               int64_t sBn, int64_t sBb1, int64_t sBb2, void* C, int64_t ldc, int64_t sCb1, int64_t sCb2, int M, int N,
               int K, int nb1, int nb2, float alpha, int a_dtype, int b_dtype, int c_dtype, void* stream);

/* hicom_colsum: out[n] += sum_m x[m,n] — bias gradients (db = 1ᵀ·dpre of nn.Linear, projector.py:307-312).  x (M,N) in
 *   dtype with row pitch ld (N and ld multiples of 4, x 16-byte aligned); out (N) fp32, ACCUMULATED into (zero it first). */
int hicom_colsum(const void* x, int64_t ld, float* out, int64_t M, int N, int dtype, void* stream);

/* hicom_act_backward: dx[i] = dy[i] * act'(pre[i]) for HICOM_ACT_* (exact erf GELU of projector.py:310, tanh GELU of the
 *   SigLIP head); pre in pre_dtype (fp32 pre-activations recomputed by hicom_linear), dy/dx in dtype; in place allowed. */
int hicom_act_backward(const void* pre, const void* dy, void* dx, int64_t n, int act, int pre_dtype, int dtype,
                       void* stream);

/* hicom_col_stats: per (video, token range, column) max m and sum l of exp(S - m) of fp32 scores S (B,N,J), S untouched;
 *   m, l (B, splits, J).  The attention backward derives the log-sum-exp of its OWN recomputed scores from them, so that
 *   exp(S - lse) sums to one exactly whatever precision the forward applied the position terms in. */
int hicom_col_stats(const float* S, float* m, float* l, int B, long long N, int J, int splits, void* stream);

/* hicom_softmax_backward: backward of the column softmax of the global attention (projector.py:213) in the
 *   reassociated form: with P[b,n,j] = exp(S[b,n,j] - lse[b,j]) and pooled[b,j] = sum_n P x'_n,
 *     dS[b,n,j] = P[b,n,j] * (dP[b,n,j] - delta[b,j]),  dP = x'_n·dpooled[b,j],  delta[b,j] = pooled[b,j]·dpooled[b,j].
 *   S, dP (B,N,J) fp32; lse, delta (B,J) fp32; dS (B,N,J) in out_dtype (may alias S or dP when fp32). */
int hicom_softmax_backward(const float* S, const float* dP, const float* lse, const float* delta, void* dS, int B,
                           int64_t N, int J, int out_dtype, void* stream);

/* hicom_grid_pool_backward: backward of hicom_grid_pool (the trilinear grid pooling of projector.py:539-540) —
 *   dX (B,T,H,W,d) fp32 += taps^T dQ (B,Nw,d).  ACCUMULATED with atomics: zero dX first (or pass a buffer that already
 *   holds other contributions to the gradient of frames_feature; train.py:712-715 'pure_vision_model'). */
int hicom_grid_pool_backward(const void* dQ, float* dX, int B, int T, int H, int W, int d, int kt, int ks, int dtype,
                             void* stream);

/* hicom_l2norm_rows_backward: backward of hicom_l2norm_rows (use_clip_scale, projector.py:184-188,527-529):
 *   dX = (dY - Y (Y·dY)) / |X| per row, Y = X / |X|; X, dY, dX (rows, d) in dtype. */
int hicom_l2norm_rows_backward(const void* X, const void* dY, void* dX, long long rows, int d, int dtype, void* stream);

/* hicom_local_attend_backward: gradients of hicom_local_attend's output (projector.py:546-553) with respect to
 *   dQ (B,Nw,d) in dtype   the query rows actually used by the forward (Q, same shape) — FiLM / instruction parameters;
 *   dK (B,T,H,W,d) fp32    the keys (frames_embed: stage 3 tunes the SigLIP head that produces it, train.py:717-721);
 *   dV (B,T,H,W,d) fp32    the values (trainable value adapter only).
 *   Any of the three may be NULL.  dK / dV are ACCUMULATED with atomics (overlapping balanced windows): zero them first.
 *   Same window rule, logit_scale and k_l2norm as the forward; dK is refused together with k_l2norm. */
int hicom_local_attend_backward(const void* Ksrc, const void* Vsrc, const void* Q, const void* dO, void* dQ, float* dK,
                                float* dV, int B, int T, int H, int W, int d, int kt, int ks, float logit_scale,
                                int k_l2norm, int dtype, void* stream);

/* hicom_film_layernorm_backward: backward of hicom_film_layernorm, y = LN(x*(1+scale)+shift)*w + b (coarse injector,
 *   projector.py:369-372).  x, dy (rows,d) and ln_w (d) in dtype; film (G,2d) fp32.  Outputs: dx (rows,d) in dtype or
 *   NULL; dfilm (G,2d), dw (d), dbias (d) fp32, ACCUMULATED into (zero them first). */
int hicom_film_layernorm_backward(const void* x, const float* film, const void* ln_w, const void* dy, void* dx,
                                  float* dfilm, float* dw, float* dbias, int64_t rows, int d, int rows_per_group,
                                  int dtype, void* stream);

/* hicom_mix_layernorm_backward: backward of hicom_mix_layernorm, out = (1-alpha)*x + alpha*(LN(y)*w + b) — the
 *   adaptq/adaptk/adaptv/adaptg mixes (projector.py:365,533-534,541).  x, y, dout (rows,d), ln_w, ln_b (d), alpha (1) in
 *   dtype.  Outputs: dx (rows,d) in dtype or NULL (x from a frozen tower); dy (rows,d) in dtype; dw, dbias (d) and
 *   dalpha (1) fp32, ACCUMULATED into (zero them first). */
int hicom_mix_layernorm_backward(const void* x, const void* y, const void* ln_w, const void* ln_b, const void* alpha,
                                 const void* dout, void* dx, void* dy, float* dw, float* dbias, float* dalpha,
                                 int64_t rows, int d, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HICOM_B200_H */
