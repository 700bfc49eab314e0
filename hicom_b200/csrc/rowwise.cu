// Row-wise and reduction helpers around the two hot kernels:
//   LayerNorm family (coarse FiLM on explicit rows, fine residual, adapter mixes),
//   position-embedding add, split-softmax column statistics, split-softmax merge,
//   and the short multi-head attention over instruction tokens.
#include "common.cuh"

namespace hicom {

// ------------------------------------------------------------------------------------------------
// LayerNorm family: one warp per row, lane owns chunks {lane, lane+32, ...} (d = 128*CPL).
// ------------------------------------------------------------------------------------------------
enum { LN_FILM = 0, LN_ADD = 1, LN_MIX = 2, LN_PLAIN = 3 };

struct LnParams {
  const void* a; const void* b; const float* film; const void* w; const void* bias; const void* alpha;
  void* out; long long rows; int d; int rows_per_group;
};

template <typename T, int CPL, int MODE>
__global__ void __launch_bounds__(256) rowwise_ln_kernel(const LnParams p) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int d = p.d;
  const T* a = static_cast<const T*>(p.a) + row * d;
  float u[CPL][4], keep[CPL][4];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int off = (lane + 32 * c) * 4;
    float x[4];
    Vec4<T>::load(a + off, x);
    if (MODE == LN_FILM) {
      const float* sc = p.film + (row / p.rows_per_group) * 2 * d;
      float s4[4], h4[4];
      Vec4<float>::load(sc + off, s4);
      Vec4<float>::load(sc + d + off, h4);
#pragma unroll
      for (int e = 0; e < 4; ++e) u[c][e] = fmaf(x[e], 1.f + s4[e], h4[e]);
    } else if (MODE == LN_ADD) {
      float y[4];
      Vec4<T>::load(static_cast<const T*>(p.b) + row * d + off, y);
#pragma unroll
      for (int e = 0; e < 4; ++e) u[c][e] = x[e] + y[e];
    } else if (MODE == LN_PLAIN) {
#pragma unroll
      for (int e = 0; e < 4; ++e) u[c][e] = x[e];
    } else {  // LN_MIX: normalise b (=y), keep a (=x)
      float y[4];
      Vec4<T>::load(static_cast<const T*>(p.b) + row * d + off, y);
#pragma unroll
      for (int e = 0; e < 4; ++e) { u[c][e] = y[e]; keep[c][e] = x[e]; }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) sum += u[c][e];
  }
  const float mean = warp_sum(sum) / (float)d;
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float dv = u[c][e] - mean; var = fmaf(dv, dv, var); }
  const float rstd = rsqrtf(warp_sum(var) / (float)d + kLnEps);
  float al = 0.f;
  if (MODE == LN_MIX) al = to_f32<T>(*static_cast<const T*>(p.alpha));
  T* o = static_cast<T*>(p.out) + row * d;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int off = (lane + 32 * c) * 4;
    float g4[4], b4[4], r[4];
    Vec4<T>::load(static_cast<const T*>(p.w) + off, g4);
    Vec4<T>::load(static_cast<const T*>(p.bias) + off, b4);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ln = fmaf((u[c][e] - mean) * rstd, g4[e], b4[e]);
      r[e] = (MODE == LN_MIX) ? fmaf(al, ln, (1.f - al) * keep[c][e]) : ln;
    }
    Vec4<T>::store(o + off, r);
  }
}

template <typename T, int MODE>
static int launch_ln(const LnParams& p, cudaStream_t stream) {
  if (p.rows == 0) return 0;
  const int wpb = 8;
  const long long blocks = (p.rows + wpb - 1) / wpb;
  HICOM_REQUIRE(blocks < (1ll << 31), "layernorm: too many rows");
  switch (p.d / 128) {
    case 9: rowwise_ln_kernel<T, 9, MODE><<<(unsigned)blocks, wpb * 32, 0, stream>>>(p); break;
    case 6: rowwise_ln_kernel<T, 6, MODE><<<(unsigned)blocks, wpb * 32, 0, stream>>>(p); break;
    case 8: rowwise_ln_kernel<T, 8, MODE><<<(unsigned)blocks, wpb * 32, 0, stream>>>(p); break;
    case 1: rowwise_ln_kernel<T, 1, MODE><<<(unsigned)blocks, wpb * 32, 0, stream>>>(p); break;
    default: set_error("layernorm: d=%d unsupported", p.d); return 1;
  }
  return check_launch("rowwise_ln_kernel");
}

// ------------------------------------------------------------------------------------------------
// x' = x + PE[t,h,w]  (projector.py:636-640), PE separable: pos_t[t] + pos_h[h] + pos_w[w].
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) posadd_kernel(const T* __restrict__ X, T* __restrict__ Y,
                                                     const float* __restrict__ pt, const float* __restrict__ ph,
                                                     const float* __restrict__ pw, long long rows, int T_, int H,
                                                     int W, int d) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int w = (int)(row % W), h = (int)((row / W) % H), t = (int)((row / ((long long)W * H)) % T_);
  const T* x = X + row * d;
  T* y = Y + row * d;
  for (int off = lane * 4; off < d; off += 128) {
    float v[4], a[4], b[4], c[4];
    Vec4<T>::load_stream(x + off, v);
    Vec4<float>::load(pt + (size_t)t * d + off, a);
    Vec4<float>::load(ph + (size_t)h * d + off, b);
    Vec4<float>::load(pw + (size_t)w * d + off, c);
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] += (a[e] + b[e]) + c[e];
    Vec4<T>::store(y + off, v);
  }
}

int launch_posadd(const void* X, void* Y, const float* pt, const float* ph, const float* pw, int B, int T_,
                  int H, int W, int d, int dtype, cudaStream_t stream) {
  const long long rows = (long long)B * T_ * H * W;
  if (rows == 0) return 0;
  const long long blocks = (rows + 7) / 8;
  HICOM_REQUIRE(blocks < (1ll << 31), "posadd: too many rows");
  HICOM_DISPATCH_DTYPE(dtype, E, posadd_kernel<E><<<(unsigned)blocks, 256, 0, stream>>>(
      static_cast<const E*>(X), static_cast<E*>(Y), pt, ph, pw, rows, T_, H, W, d));
  return check_launch("posadd_kernel");
}

// ------------------------------------------------------------------------------------------------
// Row-wise L2 normalisation: y = x / ||x||_2 (the clip-scale attention variant, projector.py:184-186).
// One warp per row, any d % 4 == 0; in place allowed.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) l2norm_rows_kernel(const T* __restrict__ X, T* __restrict__ Y, long long rows,
                                                          int d) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* x = X + row * d;
  T* y = Y + row * d;
  float ss = 0.f;
  for (int c = lane * 4; c < d; c += 128) {
    float v[4];
    Vec4<T>::load(x + c, v);
    ss += v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  }
  const float inv = 1.f / sqrtf(warp_sum(ss));
  for (int c = lane * 4; c < d; c += 128) {
    float v[4];
    Vec4<T>::load(x + c, v);
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] *= inv;
    Vec4<T>::store(y + c, v);
  }
}

// ------------------------------------------------------------------------------------------------
// Split-softmax column statistics.  S (B,N,J) fp32 scores -> in place P = exp(S - m[b,s,j]) with
// m the max over the split's token range, l the sum of P.   (projector.py:213, per split)
// grid (J/32, splits, B), block (32, 8).
// ------------------------------------------------------------------------------------------------
template <bool WRITE_P>
__global__ void __launch_bounds__(256) col_softmax_kernel(float* __restrict__ S, float* __restrict__ m_out,
                                                          float* __restrict__ l_out, int N, int J, int splits,
                                                          int rows_per_split) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x, ry = threadIdx.y;
  const int j = blockIdx.x * 32 + cx;
  const int s = blockIdx.y, b = blockIdx.z;
  const int r0 = s * rows_per_split;
  int r1 = r0 + rows_per_split;
  if (r1 > N) r1 = N;
  float* base = S + (size_t)b * N * J;
  float mx = -INFINITY;
  if (j < J)
    for (int r = r0 + ry; r < r1; r += 8) mx = fmaxf(mx, base[(size_t)r * J + j]);
  red[ry][cx] = mx;
  __syncthreads();
  if (ry == 0) {
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i][cx]);
    red[0][cx] = mx;
  }
  __syncthreads();
  mx = red[0][cx];
  __syncthreads();
  float sum = 0.f;
  if (j < J && mx > -INFINITY)
    for (int r = r0 + ry; r < r1; r += 8) {
      const float pv = exp2f((base[(size_t)r * J + j] - mx) * kLog2e);
      if (WRITE_P) base[(size_t)r * J + j] = pv;
      sum += pv;
    }
  red[ry][cx] = sum;
  __syncthreads();
  if (ry == 0 && j < J) {
#pragma unroll
    for (int i = 1; i < 8; ++i) sum += red[i][cx];
    m_out[((size_t)b * splits + s) * J + j] = mx;
    l_out[((size_t)b * splits + s) * J + j] = sum;
  }
}

int launch_col_softmax(float* S, float* m, float* l, int B, int N, int J, int splits, int rows_per_split,
                       cudaStream_t stream) {
  if (B == 0) return 0;
  dim3 grid((J + 31) / 32, splits, B), block(32, 8);
  HICOM_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "col_softmax: grid too large");
  col_softmax_kernel<true><<<grid, block, 0, stream>>>(S, m, l, N, J, splits, rows_per_split);
  return check_launch("col_softmax_kernel");
}

}  // namespace hicom

// Column statistics of a score tensor WITHOUT touching it: per (video, token range, column) the max and the sum of
// exp(S - max) — the attention backward takes the log-sum-exp of its own recomputed scores from them (hicom_col_stats).
extern "C" int hicom_col_stats(const float* S, float* m, float* l, int B, long long N, int J, int splits, void* stream) {
  using namespace hicom;
  HICOM_REQUIRE(S && m && l, "col_stats: null pointer");
  HICOM_REQUIRE(B >= 0 && N > 0 && J > 0 && splits > 0 && N < (1ll << 31), "col_stats: bad shape");
  if (B == 0) return 0;
  const int rows = (int)((N + splits - 1) / splits);
  dim3 grid((J + 31) / 32, splits, B), block(32, 8);
  HICOM_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "col_stats: grid too large");
  col_softmax_kernel<false><<<grid, block, 0, as_stream(stream)>>>(const_cast<float*>(S), m, l, (int)N, J, splits, rows);
  return check_launch("col_stats_kernel");
}

namespace hicom {

// ------------------------------------------------------------------------------------------------
// Split-softmax merge: one block per (video, column).
// ------------------------------------------------------------------------------------------------
// PARTIAL = false: pooled = sum_p o_p e^{m_p-M} / L  (normalised, TO storage)
// PARTIAL = true : emits ONE un-normalised partial (M, L, sum_p o_p e^{m_p-M}) in fp32 — what a rank sends over
//                  NVLink after reducing its own token splits (hicom_softmax_reduce).
template <typename TO, bool PARTIAL>
__global__ void __launch_bounds__(128) softmax_merge_kernel(const float* __restrict__ m, const float* __restrict__ l,
                                                            const float* __restrict__ o, TO* __restrict__ pooled,
                                                            float* __restrict__ m_out, float* __restrict__ l_out,
                                                            int P, int J, int d) {
  const int j = blockIdx.x, b = blockIdx.y;
  const float* mp = m + (size_t)b * P * J + j;
  const float* lp = l + (size_t)b * P * J + j;
  float M = -INFINITY;
  for (int pidx = 0; pidx < P; ++pidx) M = fmaxf(M, mp[(size_t)pidx * J]);
  float L = 0.f;
  for (int pidx = 0; pidx < P; ++pidx) {
    const float mv = mp[(size_t)pidx * J];
    if (mv > -INFINITY) L += lp[(size_t)pidx * J] * exp2f((mv - M) * kLog2e);
  }
  if (PARTIAL && threadIdx.x == 0) {
    m_out[(size_t)b * J + j] = M;
    l_out[(size_t)b * J + j] = L;
  }
  // normalising merge with m_out set: also emit the log-sum-exp of the column (what a frame shard sends to its peers)
  if (!PARTIAL && m_out != nullptr && threadIdx.x == 0) m_out[(size_t)b * J + j] = M + logf(L);
  const float invL = PARTIAL ? 1.f : 1.f / L;
  for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int pidx = 0; pidx < P; ++pidx) {
      const float mv = mp[(size_t)pidx * J];
      if (!(mv > -INFINITY)) continue;  // empty split: contributes nothing
      const float wgt = exp2f((mv - M) * kLog2e);
      float v[4];
      Vec4<float>::load(o + (((size_t)b * P + pidx) * J + j) * d + c, v);
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] = fmaf(wgt, v[e], acc[e]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] *= invL;
    Vec4<TO>::store(pooled + ((size_t)b * J + j) * d + c, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// Short multi-head attention (fine injector): one warp per (group, query row, head); head_dim 128.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) guide_attend_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                           const T* __restrict__ v, T* __restrict__ out,
                                                           long long total, int Mq, int L, int d, int heads,
                                                           float scale_log2) {
  const int lane = threadIdx.x & 31;
  const long long wi = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wi >= total) return;
  const int h = (int)(wi % heads);
  const long long row = wi / heads;  // g*Mq + i
  const long long g = row / Mq;
  const int coff = h * 128 + lane * 4;
  float qv[4];
  Vec4<T>::load(q + row * d + coff, qv);
#pragma unroll
  for (int e = 0; e < 4; ++e) qv[e] *= scale_log2;
  float m_run = -INFINITY, l_run = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n = 0; n < L; ++n) {
    float kv[4], vv[4];
    Vec4<T>::load(k + (g * L + n) * d + coff, kv);
    Vec4<T>::load(v + (g * L + n) * d + coff, vv);
    float s = qv[0] * kv[0] + qv[1] * kv[1] + qv[2] * kv[2] + qv[3] * kv[3];
    s = warp_sum(s);
    const float m_new = fmaxf(m_run, s);
    const float corr = exp2f(m_run - m_new), pv = exp2f(s - m_new);
    l_run = fmaf(l_run, corr, pv);
    m_run = m_new;
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] = fmaf(acc[e], corr, pv * vv[e]);
  }
  const float inv = 1.f / l_run;
#pragma unroll
  for (int e = 0; e < 4; ++e) acc[e] *= inv;
  Vec4<T>::store(out + row * d + coff, acc);
}

}  // namespace hicom

using namespace hicom;

extern "C" int hicom_layernorm(const void* x, const void* ln_w, const void* ln_b, void* out, int64_t rows, int d,
                               int dtype, void* stream) {
  HICOM_REQUIRE(x && ln_w && ln_b && out, "layernorm: null pointer");
  HICOM_REQUIRE(rows >= 0 && d > 0 && d % 128 == 0, "layernorm: bad shape");
  LnParams p{}; p.a = x; p.w = ln_w; p.bias = ln_b; p.out = out; p.rows = rows; p.d = d; p.rows_per_group = 1;
  HICOM_DISPATCH_DTYPE(dtype, E, return (launch_ln<E, LN_PLAIN>(p, as_stream(stream))));
}

extern "C" int hicom_film_layernorm(const void* x, const float* film, const void* ln_w, const void* ln_b,
                                    void* out, int rows, int d, int rows_per_group, int dtype, void* stream) {
  HICOM_REQUIRE(x && film && ln_w && ln_b && out, "film_layernorm: null pointer");
  HICOM_REQUIRE(rows >= 0 && d > 0 && d % 128 == 0 && rows_per_group > 0, "film_layernorm: bad shape");
  LnParams p{}; p.a = x; p.film = film; p.w = ln_w; p.bias = ln_b; p.out = out; p.rows = rows; p.d = d;
  p.rows_per_group = rows_per_group;
  HICOM_DISPATCH_DTYPE(dtype, E, return (launch_ln<E, LN_FILM>(p, as_stream(stream))));
}

extern "C" int hicom_add_layernorm(const void* a, const void* b, const void* ln_w, const void* ln_b, void* out,
                                   int rows, int d, int dtype, void* stream) {
  HICOM_REQUIRE(a && b && ln_w && ln_b && out, "add_layernorm: null pointer");
  HICOM_REQUIRE(rows >= 0 && d > 0 && d % 128 == 0, "add_layernorm: bad shape");
  LnParams p{}; p.a = a; p.b = b; p.w = ln_w; p.bias = ln_b; p.out = out; p.rows = rows; p.d = d; p.rows_per_group = 1;
  HICOM_DISPATCH_DTYPE(dtype, E, return (launch_ln<E, LN_ADD>(p, as_stream(stream))));
}

extern "C" int hicom_mix_layernorm(const void* x, const void* y, const void* ln_w, const void* ln_b,
                                   const void* alpha, void* out, int64_t rows, int d, int dtype, void* stream) {
  HICOM_REQUIRE(x && y && ln_w && ln_b && alpha && out, "mix_layernorm: null pointer");
  HICOM_REQUIRE(rows >= 0 && d > 0 && d % 128 == 0, "mix_layernorm: bad shape");
  LnParams p{}; p.a = x; p.b = y; p.w = ln_w; p.bias = ln_b; p.alpha = alpha; p.out = out; p.rows = rows; p.d = d;
  p.rows_per_group = 1;
  HICOM_DISPATCH_DTYPE(dtype, E, return (launch_ln<E, LN_MIX>(p, as_stream(stream))));
}

extern "C" int hicom_guide_attend(const void* q, const void* k, const void* v, void* out, int G, int Mq, int L,
                                  int d, int heads, float scale, int dtype, void* stream) {
  HICOM_REQUIRE(q && k && v && out, "guide_attend: null pointer");
  HICOM_REQUIRE(G >= 0 && Mq >= 0 && L > 0 && heads > 0 && d == heads * 128,
                "guide_attend: needs head_dim 128 (d=%d heads=%d)", d, heads);
  const long long total = (long long)G * Mq * heads;
  if (total == 0) return 0;
  const long long blocks = (total + 7) / 8;
  HICOM_REQUIRE(blocks < (1ll << 31), "guide_attend: too many rows");
  HICOM_DISPATCH_DTYPE(dtype, E, guide_attend_kernel<E><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      static_cast<const E*>(q), static_cast<const E*>(k), static_cast<const E*>(v), static_cast<E*>(out), total,
      Mq, L, d, heads, scale * kLog2e));
  return check_launch("guide_attend_kernel");
}

extern "C" int hicom_softmax_merge(const float* m, const float* l, const float* o, int B, int P, int J, int d,
                                   void* pooled, int out_dtype, void* stream) {
  HICOM_REQUIRE(m && l && o && pooled, "softmax_merge: null pointer");
  HICOM_REQUIRE(B >= 0 && P > 0 && J > 0 && d > 0 && d % 4 == 0, "softmax_merge: bad shape");
  if (B == 0) return 0;
  dim3 grid(J, B);
  HICOM_REQUIRE(B <= 65535, "softmax_merge: batch too large");
  HICOM_DISPATCH_DTYPE(out_dtype, E, (softmax_merge_kernel<E, false><<<grid, 128, 0, as_stream(stream)>>>(
      m, l, o, static_cast<E*>(pooled), nullptr, nullptr, P, J, d)));
  return check_launch("softmax_merge_kernel");
}

extern "C" int hicom_softmax_merge_lse(const float* m, const float* l, const float* o, int B, int P, int J, int d,
                                       void* pooled, int out_dtype, float* lse, void* stream) {
  HICOM_REQUIRE(m && l && o && pooled && lse, "softmax_merge_lse: null pointer");
  HICOM_REQUIRE(B >= 0 && P > 0 && J > 0 && d > 0 && d % 4 == 0 && B <= 65535, "softmax_merge_lse: bad shape");
  if (B == 0) return 0;
  dim3 grid(J, B);
  HICOM_DISPATCH_DTYPE(out_dtype, E, (softmax_merge_kernel<E, false><<<grid, 128, 0, as_stream(stream)>>>(
      m, l, o, static_cast<E*>(pooled), lse, nullptr, P, J, d)));
  return check_launch("softmax_merge_lse_kernel");
}

// Frame shards: rank r holds attn_r (B, Q, d) — its own normalised attention output after the value projection (linear
// per head, so it commutes with the merge) — and lse_r (B, heads*Q).  out[b,i,h*hd+c] = sum_r w_r attn_r[b,i,h*hd+c],
// w_r = softmax over r of lse_r[b, h*Q+i].
template <typename T>
__global__ void __launch_bounds__(128) shard_combine_kernel(const uint8_t* __restrict__ msgs, size_t rank_stride,
                                                            size_t video_stride, size_t lse_offset, int R, int Q, int d,
                                                            int heads, T* __restrict__ out) {
  const int i = blockIdx.x, b = blockIdx.y;
  const int hd = d / heads;
  for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
    const int h = c / hd;  // hd % 4 == 0: the four channels share a head
    float M = -INFINITY;
    for (int r = 0; r < R; ++r) {
      const float* lse = reinterpret_cast<const float*>(msgs + r * rank_stride + b * video_stride + lse_offset);
      M = fmaxf(M, lse[h * Q + i]);
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, L = 0.f;
    for (int r = 0; r < R; ++r) {
      const uint8_t* base = msgs + r * rank_stride + b * video_stride;
      const float w = exp2f((reinterpret_cast<const float*>(base + lse_offset)[h * Q + i] - M) * kLog2e);
      float v[4];
      Vec4<T>::load(reinterpret_cast<const T*>(base) + (size_t)i * d + c, v);
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] = fmaf(w, v[e], acc[e]);
      L += w;
    }
    const float inv = 1.f / L;
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] *= inv;
    Vec4<T>::store(out + ((size_t)b * Q + i) * d + c, acc);
  }
}

extern "C" int hicom_shard_combine(const void* msgs, long long rank_stride_bytes, long long video_stride_bytes,
                                   long long lse_offset_bytes, int R, int B, int Q, int d, int heads, void* out,
                                   int dtype, void* stream) {
  HICOM_REQUIRE(msgs && out, "shard_combine: null pointer");
  HICOM_REQUIRE(R > 0 && B >= 0 && Q > 0 && heads > 0 && d % heads == 0 && (d / heads) % 4 == 0 && B <= 65535,
                "shard_combine: bad shape");
  HICOM_REQUIRE(lse_offset_bytes % 4 == 0 && video_stride_bytes % 16 == 0 && rank_stride_bytes % 16 == 0 &&
                ((uintptr_t)msgs & 15) == 0, "shard_combine: message not 16-byte aligned");
  HICOM_REQUIRE(dtype != HICOM_F32 || lse_offset_bytes >= (long long)Q * d * 4, "shard_combine: lse overlaps the rows");
  if (B == 0) return 0;
  dim3 grid(Q, B);
  HICOM_DISPATCH_DTYPE(dtype, E, (shard_combine_kernel<E><<<grid, 128, 0, as_stream(stream)>>>(
      static_cast<const uint8_t*>(msgs), (size_t)rank_stride_bytes, (size_t)video_stride_bytes,
      (size_t)lse_offset_bytes, R, Q, d, heads, static_cast<E*>(out))));
  return check_launch("shard_combine_kernel");
}

extern "C" int hicom_softmax_reduce(const float* m, const float* l, const float* o, int B, int P, int J, int d,
                                    float* m_out, float* l_out, float* o_out, void* stream) {
  HICOM_REQUIRE(m && l && o && m_out && l_out && o_out, "softmax_reduce: null pointer");
  HICOM_REQUIRE(B >= 0 && P > 0 && J > 0 && d > 0 && d % 4 == 0 && B <= 65535, "softmax_reduce: bad shape");
  if (B == 0) return 0;
  dim3 grid(J, B);
  softmax_merge_kernel<float, true><<<grid, 128, 0, as_stream(stream)>>>(m, l, o, o_out, m_out, l_out, P, J, d);
  return check_launch("softmax_reduce_kernel");
}

extern "C" int hicom_posadd(const void* X, void* Y, const float* pos_t, const float* pos_h, const float* pos_w, int B,
                            int T, int H, int W, int d, int dtype, void* stream) {
  HICOM_REQUIRE(X && Y && pos_t && pos_h && pos_w, "posadd: null pointer");
  HICOM_REQUIRE(B >= 0 && T > 0 && H > 0 && W > 0 && d > 0 && d % 128 == 0, "posadd: bad shape");
  return hicom::launch_posadd(X, Y, pos_t, pos_h, pos_w, B, T, H, W, d, dtype, as_stream(stream));
}

extern "C" int hicom_l2norm_rows(const void* X, void* Y, long long rows, int d, int dtype, void* stream) {
  HICOM_REQUIRE(X && Y, "l2norm_rows: null pointer");
  HICOM_REQUIRE(rows >= 0 && d > 0 && d % 4 == 0, "l2norm_rows: bad shape");
  if (rows == 0) return 0;
  const long long blocks = (rows + 7) / 8;
  HICOM_REQUIRE(blocks < (1ll << 31), "l2norm_rows: too many rows");
  HICOM_DISPATCH_DTYPE(dtype, E, (hicom::l2norm_rows_kernel<E><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      static_cast<const E*>(X), static_cast<E*>(Y), rows, d)));
  return check_launch("l2norm_rows_kernel");
}
