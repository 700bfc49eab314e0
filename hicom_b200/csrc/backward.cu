// Backward pass building blocks (stages 1-3 of the reference TRAIN the projector, train.py:704-738).
//
// Every contraction of the backward formulas goes through hicom_gemm: the generic strided SIMT GEMM (fp32 mode, small
// problems, odd strides) or — large bf16 problems in the NT / NN / TN layouts — the persistent tcgen05 kernel of
// gemm_tc.cu (try_gemm_tc below).  The element-wise pieces around them live here.
//   hicom_gemm              C = alpha * A·B over arbitrary element strides, two batch levels
//   hicom_act_backward      dx = dy * act'(pre)                      (GELU erf / tanh, projector.py:310, encoder.py:285)
//   hicom_softmax_backward  dS = exp(S - lse) * (dP - delta)         (softmax of projector.py:213 in reassociated form)
//   hicom_local_attend_backward         d(query), d(keys), d(values) of the window attention (projector.py:546-553)
//   hicom_film_layernorm_backward       backward of LN(x*(1+scale)+shift) (projector.py:369-372)
//   hicom_mix_layernorm_backward        backward of (1-alpha)*x + alpha*LN(y), the adapter mixes (projector.py:365,533-534,541)
#include <stdlib.h>

#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

namespace hicom {

// ------------------------------------------------------------------------------------------------
// dx = dy * act'(pre)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf_grad(float x) {
  // d/dx [x Phi(x)] = Phi(x) + x phi(x)
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return fmaf(x, pdf, cdf);
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float c = 0.7978845608028654f, a = 0.044715f;
  const float t = tanhf(c * fmaf(a * x * x, x, x));
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * c * fmaf(3.0f * a * x, x, 1.0f);
}

template <typename TP, typename T>
__global__ void __launch_bounds__(256) act_backward_kernel(const TP* __restrict__ pre, const T* __restrict__ dy,
                                                           T* __restrict__ dx, long long n, int act) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float x = to_f32<TP>(pre[i]);
    const float g = act == HICOM_ACT_GELU ? gelu_erf_grad(x) : (act == HICOM_ACT_GELU_TANH ? gelu_tanh_grad(x) : 1.0f);
    dx[i] = from_f32<T>(to_f32<T>(dy[i]) * g);
  }
}

// ------------------------------------------------------------------------------------------------
// out[n] += sum_m x[m, n]  (bias gradients, db = 1ᵀ·dpre).  Block = 32 column groups of 4 x 8 row lanes; rows strided
// over gridDim.y; the 8 row lanes are folded in shared memory, blocks of one column range meet in fp32 atomics.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long ld, float* __restrict__ out,
                                                     long long M, int N) {
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cx) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (col < N) {
    for (long long r = (long long)blockIdx.y * 8 + ry; r < M; r += (long long)gridDim.y * 8) {
      float v[4];
      Vec4<T>::load(x + r * ld + col, v);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] += v[i];
    }
  }
  __shared__ float red[8][32][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) red[ry][cx][i] = acc[i];
  __syncthreads();
  if (ry == 0 && col < N) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[k][cx][i];
      atomicAdd(out + col + i, s);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dS[b,n,j] = exp(S[b,n,j] - lse[b,j]) * (dP[b,n,j] - delta[b,j])
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) softmax_backward_kernel(const float* __restrict__ S, const float* __restrict__ dP,
                                                               const float* __restrict__ lse,
                                                               const float* __restrict__ delta, T* __restrict__ dS,
                                                               long long NJ, int J, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long b = i / NJ;
    const int j = (int)(i % J);  // NJ is a multiple of J, so i % J is the column
    const float p = __expf(S[i] - lse[b * J + j]);
    dS[i] = from_f32<T>(p * (dP[i] - delta[b * J + j]));
  }
}

// ------------------------------------------------------------------------------------------------
// Window attention backward (projector.py:546-553):
//   s_k = scale * q·k_k,  p = softmax_k(s),  o = sum_k p_k v_k
//   dp_k = dO·v_k,  ds_k = p_k (dp_k - sum_k p_k dp_k)
//   dq = scale * sum_k ds_k k_k        (query rows: FiLM / instruction parameters upstream)
//   dk_k = scale * ds_k q              (keys = frames_embed: stage 3 tunes the SigLIP head that makes it, train.py:717-721)
//   dv_k = p_k dO                      (values: only with a trainable value adapter)
// One warp per window; lane L owns channels {4(L + 32c)}.  Two passes over the window members (statistics, then the
// gradients), members addressed with the reference's balanced-window starts (common.cuh).  dK / dV are fp32 and
// ACCUMULATED with atomics: balanced windows of non-divisible grids overlap by one element, so a token can belong to
// two windows per axis.  Keys/values are read with plain vector loads: this runs once per training step, next to
// GEMMs over the same tokens.
// ------------------------------------------------------------------------------------------------
template <typename T, int CPL>
__global__ void __launch_bounds__(256) local_attend_bwd_q_kernel(const T* __restrict__ Ksrc, const T* __restrict__ Vsrc,
                                                                 const T* __restrict__ Q, const T* __restrict__ dO,
                                                                 T* __restrict__ dQ, float* __restrict__ dK,
                                                                 float* __restrict__ dV, int B, int T_, int H, int W, int d,
                                                                 AxisWin at, AxisWin ah, AxisWin aw, float scale,
                                                                 int k_l2norm) {
  const int lane = threadIdx.x & 31;
  const int h1 = ah.count, w1 = aw.count;
  const long long nw = (long long)at.count * h1 * w1;
  const long long win = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (win >= (long long)B * nw) return;
  const int b = (int)(win / nw);
  const int r = (int)(win % nw);
  const int wt = r / (h1 * w1), wh = (r / w1) % h1, ww = r % w1;
  const int ts = axis_win_start(at, wt), hs = axis_win_start(ah, wh), ws = axis_win_start(aw, ww);
  const int nt = at.len, nh = ah.len, nwid = aw.len;

  float q[CPL][4], go[CPL][4];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int off = (lane + 32 * c) * 4;
    Vec4<T>::load(Q + win * d + off, q[c]);
    Vec4<T>::load(dO + win * d + off, go[c]);
  }
  const int members = nt * nh * nwid;
  // pass 1: softmax statistics and sum_k p_k dp_k
  float m = -INFINITY, l = 0.f, pd = 0.f;
  for (int e = 0; e < members; ++e) {
    const int t = ts + e / (nh * nwid), h = hs + (e / nwid) % nh, w = ws + e % nwid;
    const long long tok = (((long long)b * T_ + t) * H + h) * W + w;
    float s = 0.f, dp = 0.f, kk = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int off = (lane + 32 * c) * 4;
      float k4[4], v4[4];
      Vec4<T>::load(Ksrc + tok * d + off, k4);
      Vec4<T>::load(Vsrc + tok * d + off, v4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s = fmaf(q[c][i], k4[i], s);
        dp = fmaf(go[c][i], v4[i], dp);
        kk = fmaf(k4[i], k4[i], kk);
      }
    }
    s = warp_sum(s); dp = warp_sum(dp);
    if (k_l2norm) s *= rsqrtf(warp_sum(kk));
    s *= scale;
    const float mn = fmaxf(m, s);
    const float corr = __expf(m - mn), p = __expf(s - mn);
    l = l * corr + p;
    pd = pd * corr + p * dp;
    m = mn;
  }
  const float inv_l = 1.f / l;
  const float delta = pd * inv_l;
  // pass 2: dq = scale * sum_k p_k (dp_k - delta) k_k   (k normalised when k_l2norm)
  float acc[CPL][4];
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
  for (int e = 0; e < members; ++e) {
    const int t = ts + e / (nh * nwid), h = hs + (e / nwid) % nh, w = ws + e % nwid;
    const long long tok = (((long long)b * T_ + t) * H + h) * W + w;
    float k[CPL][4];
    float s = 0.f, dp = 0.f, kk = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int off = (lane + 32 * c) * 4;
      float v4[4];
      Vec4<T>::load(Ksrc + tok * d + off, k[c]);
      Vec4<T>::load(Vsrc + tok * d + off, v4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s = fmaf(q[c][i], k[c][i], s);
        dp = fmaf(go[c][i], v4[i], dp);
        kk = fmaf(k[c][i], k[c][i], kk);
      }
    }
    s = warp_sum(s); dp = warp_sum(dp);
    const float rn = k_l2norm ? rsqrtf(warp_sum(kk)) : 1.f;
    s *= rn * scale;
    const float p = __expf(s - m) * inv_l;
    const float ds = p * (dp - delta) * scale * rn;
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[c][i] = fmaf(ds, k[c][i], acc[c][i]);
    if (dK != nullptr) {  // (host refuses k_l2norm together with dK: the normalisation is not differentiated)
#pragma unroll
      for (int c = 0; c < CPL; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) atomicAdd(dK + tok * d + (lane + 32 * c) * 4 + i, ds * q[c][i]);
    }
    if (dV != nullptr) {
#pragma unroll
      for (int c = 0; c < CPL; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) atomicAdd(dV + tok * d + (lane + 32 * c) * 4 + i, p * go[c][i]);
    }
  }
  if (dQ != nullptr) {
#pragma unroll
    for (int c = 0; c < CPL; ++c) Vec4<T>::store(dQ + win * d + (lane + 32 * c) * 4, acc[c]);
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of y = LN(u) * w + bias, u = x * (1 + scale_g) + shift_g  (coarse injector, projector.py:369-372), one warp
// per row, group g = row / rows_per_group:
//   g_hat = dy * w;  du = rstd * (g_hat - mean(g_hat) - u_hat * mean(g_hat * u_hat))
//   dx = du * (1 + scale);  dscale_g += du * x;  dshift_g += du;  dw += dy * u_hat;  dbias += dy     (fp32 atomics)
// dx may be null (x = pooled features of a frozen tower).  dfilm (G, 2d), dw, dbias (d) are fp32 and must be zeroed
// by the caller.
// ------------------------------------------------------------------------------------------------
template <typename T, int CPL>
__global__ void __launch_bounds__(256) film_ln_backward_kernel(const T* __restrict__ x, const float* __restrict__ film,
                                                               const T* __restrict__ w, const T* __restrict__ dy,
                                                               T* __restrict__ dx, float* __restrict__ dfilm,
                                                               float* __restrict__ dw, float* __restrict__ dbias,
                                                               long long rows, int d, int rows_per_group) {
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  // a block walks a contiguous range of rows (of ONE group when rows_per_group is a multiple of the range) so that
  // the per-channel sums are accumulated in registers and flushed with one atomic per channel per warp and group
  const long long rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
  const long long r_begin = (long long)blockIdx.x * rows_per_block;
  long long r_end = r_begin + rows_per_block;
  if (r_end > rows) r_end = rows;
  float aw[CPL][4], ab[CPL][4], asc[CPL][4], ash[CPL][4];
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) aw[c][i] = ab[c][i] = asc[c][i] = ash[c][i] = 0.f;
  long long cur_group = -1;
  auto flush_film = [&]() {
    if (cur_group < 0) return;
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ch = (lane + 32 * c) * 4 + i;
        atomicAdd(dfilm + cur_group * 2 * d + ch, asc[c][i]);
        atomicAdd(dfilm + cur_group * 2 * d + d + ch, ash[c][i]);
        asc[c][i] = ash[c][i] = 0.f;
      }
  };
  for (long long row = r_begin + wib; row < r_end; row += wpb) {
    const long long g = row / rows_per_group;
    if (g != cur_group) { flush_film(); cur_group = g; }
    const float* sc = film + g * 2 * d;
    float xv[CPL][4], u[CPL][4], gy[CPL][4], s1[CPL][4];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int off = (lane + 32 * c) * 4;
      float h4[4];
      Vec4<T>::load(x + row * d + off, xv[c]);
      Vec4<T>::load(dy + row * d + off, gy[c]);
      Vec4<float>::load(sc + off, s1[c]);
      Vec4<float>::load(sc + d + off, h4);
#pragma unroll
      for (int i = 0; i < 4; ++i) { u[c][i] = fmaf(xv[c][i], 1.f + s1[c][i], h4[i]); sum += u[c][i]; }
    }
    const float mean = warp_sum(sum) / (float)d;
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float dv = u[c][i] - mean; var = fmaf(dv, dv, var); }
    const float rstd = rsqrtf(warp_sum(var) / (float)d + kLnEps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int off = (lane + 32 * c) * 4;
      float w4[4];
      Vec4<T>::load(w + off, w4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float uh = (u[c][i] - mean) * rstd;
        aw[c][i] = fmaf(gy[c][i], uh, aw[c][i]);
        ab[c][i] += gy[c][i];
        const float gh = gy[c][i] * w4[i];
        u[c][i] = uh;      // keep u_hat
        gy[c][i] = gh;     // keep g_hat
        m1 += gh;
        m2 = fmaf(gh, uh, m2);
      }
    }
    m1 = warp_sum(m1) / (float)d;
    m2 = warp_sum(m2) / (float)d;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      float o4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float du = rstd * (gy[c][i] - m1 - u[c][i] * m2);
        asc[c][i] = fmaf(du, xv[c][i], asc[c][i]);
        ash[c][i] += du;
        o4[i] = du * (1.f + s1[c][i]);
      }
      if (dx) Vec4<T>::store(dx + row * d + (lane + 32 * c) * 4, o4);
    }
  }
  flush_film();
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ch = (lane + 32 * c) * 4 + i;
      atomicAdd(dw + ch, aw[c][i]);
      atomicAdd(dbias + ch, ab[c][i]);
    }
}

// ------------------------------------------------------------------------------------------------
// Backward of out = (1 - alpha) * x + alpha * (LN(y) * w + bias)   (adapter mixes, projector.py:365,533-534,541), one warp
// per row, a block walks a contiguous range of rows:
//   dx = (1 - alpha) * dout;  g = alpha * dout;  g_hat = g * w;
//   dy = rstd * (g_hat - mean(g_hat) - y_hat * mean(g_hat * y_hat))
//   dw += g * y_hat;  dbias += g;  dalpha += sum dout * (LN(y)*w + bias - x)                  (fp32 atomics)
// dx may be null (x from a frozen tower).  dw, dbias (d) and dalpha (1) are fp32 and must be zeroed by the caller.
// ------------------------------------------------------------------------------------------------
template <typename T, int CPL>
__global__ void __launch_bounds__(256) mix_ln_backward_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                                              const T* __restrict__ w, const T* __restrict__ bias,
                                                              const T* __restrict__ alpha, const T* __restrict__ dout,
                                                              T* __restrict__ dx, T* __restrict__ dy,
                                                              float* __restrict__ dw, float* __restrict__ dbias,
                                                              float* __restrict__ dalpha, long long rows, int d) {
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const long long rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
  const long long r_begin = (long long)blockIdx.x * rows_per_block;
  long long r_end = r_begin + rows_per_block;
  if (r_end > rows) r_end = rows;
  const float al = to_f32<T>(*alpha);
  float aw[CPL][4], ab[CPL][4];
  float aal = 0.f;
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) aw[c][i] = ab[c][i] = 0.f;
  for (long long row = r_begin + wib; row < r_end; row += wpb) {
    float u[CPL][4], go[CPL][4];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int off = (lane + 32 * c) * 4;
      Vec4<T>::load(y + row * d + off, u[c]);
      Vec4<T>::load(dout + row * d + off, go[c]);
#pragma unroll
      for (int i = 0; i < 4; ++i) sum += u[c][i];
    }
    const float mean = warp_sum(sum) / (float)d;
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float dv = u[c][i] - mean; var = fmaf(dv, dv, var); }
    const float rstd = rsqrtf(warp_sum(var) / (float)d + kLnEps);
    float m1 = 0.f, m2 = 0.f;
    float gh[CPL][4];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int off = (lane + 32 * c) * 4;
      float w4[4], b4[4], x4[4], o4[4];
      Vec4<T>::load(w + off, w4);
      Vec4<T>::load(bias + off, b4);
      Vec4<T>::load(x + row * d + off, x4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float uh = (u[c][i] - mean) * rstd;
        const float ln = fmaf(uh, w4[i], b4[i]);
        aal = fmaf(go[c][i], ln - x4[i], aal);
        const float g = al * go[c][i];
        aw[c][i] = fmaf(g, uh, aw[c][i]);
        ab[c][i] += g;
        gh[c][i] = g * w4[i];
        u[c][i] = uh;
        m1 += gh[c][i];
        m2 = fmaf(gh[c][i], uh, m2);
        o4[i] = (1.f - al) * go[c][i];
      }
      if (dx) Vec4<T>::store(dx + row * d + off, o4);
    }
    m1 = warp_sum(m1) / (float)d;
    m2 = warp_sum(m2) / (float)d;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      float o4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o4[i] = rstd * (gh[c][i] - m1 - u[c][i] * m2);
      Vec4<T>::store(dy + row * d + (lane + 32 * c) * 4, o4);
    }
  }
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ch = (lane + 32 * c) * 4 + i;
      atomicAdd(dw + ch, aw[c][i]);
      atomicAdd(dbias + ch, ab[c][i]);
    }
  aal = warp_sum(aal);
  if (lane == 0) atomicAdd(dalpha, aal);
}

}  // namespace hicom

using namespace hicom;

// Large bf16 contractions of the backward go to the persistent tcgen05 kernel when their strides match one of the
// operand layouts it already serves in the forward (per batch entry, one launch each):
//   NT  A (M,K) K-major, B given as (N,K) K-major   S = x'·qfoldᵀ, dP = x'·dpooledᵀ            -> plain linear
//   NN  A (M,K) K-major, B (K,N) row-major          dA = dpre·W                                 -> MN-major B (w_is_kn)
//   TN  A stored (K,M),  B (K,N) row-major, fp32 C  dW = dpreᵀ·A, dqfold = dSᵀ·x'               -> MN-major A and B
// Returns -1 when the problem stays on the SIMT kernel.

static int try_gemm_tc(const void* A, int64_t sAm, int64_t sAk, int64_t sAb1, int64_t sAb2, const void* B, int64_t sBk,
                       int64_t sBn, int64_t sBb1, int64_t sBb2, void* C, int64_t ldc, int64_t sCb1, int64_t sCb2, int M,
                       int N, int K, int nb1, int nb2, float alpha, int a_dtype, int b_dtype, int c_dtype,
                       cudaStream_t stream) {
  if (a_dtype != HICOM_BF16 || b_dtype != HICOM_BF16) return -1;
  if (M < 2 || N < 2 || K < 16 || (long long)M * N * K < (1ll << 24)) return -1;  // small: launch-bound either way
  const long long nb = (long long)nb1 * nb2;
  if (nb > 4096) return -1;  // one launch per batch entry (per video): beyond this the batched SIMT grid is the better deal
  const bool a_kmajor = sAk == 1, a_mmajor = sAm == 1 && !a_kmajor;
  const bool b_kmajor = sBk == 1, b_nmajor = sBn == 1 && !b_kmajor;
  int mode;  // 0 NT, 1 NN, 2 TN
  long long lda, ldw;
  if (a_kmajor && b_kmajor) { mode = 0; lda = sAm; ldw = sBn; }
  else if (a_kmajor && b_nmajor) { mode = 1; lda = sAm; ldw = sBk; }
  else if (a_mmajor && b_nmajor) { mode = 2; lda = sAk; ldw = sBk; }
  else return -1;
  if (mode == 2 && (c_dtype != HICOM_F32 || alpha != 1.0f)) return -1;
  if (mode != 2 && K % 8 != 0) return -1;
  if (lda % 8 != 0 || ldw % 8 != 0) return -1;
  if (nb > 1 && ((nb1 > 1 && (sAb1 % 8 || sBb1 % 8)) || (nb2 > 1 && (sAb2 % 8 || sBb2 % 8)))) return -1;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return -1;
  const size_t csz = c_dtype == HICOM_F32 ? 4 : 2;
  // all batch entries of a TN problem in ONE launch through the kernel's batch axis (per-video dqfold = dSᵀ·x' has
  // only 15 tiles, so B launches leave most SMs idle)
  if (mode == 2 && nb1 == 1 && nb2 > 1 && sCb2 % ldc == 0 && sAb2 > 0 && sBb2 > 0) {
    TcLinearParams t{};
    t.A = A; t.W = B; t.C = C; t.bias = nullptr; t.R = nullptr; t.ldr = 0;
    t.lda = lda; t.ldw = ldw; t.ldc = ldc; t.M = M; t.N = N; t.K = K;
    t.act = HICOM_ACT_NONE; t.out_dtype = c_dtype; t.rows_per_group = 1 << 30; t.group_stride_rows = 0;
    t.alpha = 1.0f; t.w_is_kn = 1; t.a_is_km = 1;
    t.batch = nb2; t.a_batch_stride = sAb2; t.w_batch_stride = sBb2; t.c_batch_rows = sCb2 / ldc;
    return launch_tc_linear(t, stream);
  }
  // all batch entries of an NT problem in ONE launch (S[b] = x'[b]·qfold[b]ᵀ, dP[b] = x'[b]·dpooled[b]ᵀ per video):
  // the kernel's batch axis walks A, C and (when it has a batch stride) B
  if (mode == 0 && nb > 1 && (nb1 == 1 || nb2 == 1)) {
    const long long sa = nb1 == 1 ? sAb2 : sAb1, sb = nb1 == 1 ? sBb2 : sBb1, sc = nb1 == 1 ? sCb2 : sCb1;
    if (sa > 0 && sb >= 0 && sc > 0 && sc % ldc == 0) {
      TcLinearParams t{};
      t.A = A; t.W = B; t.C = C; t.bias = nullptr; t.R = nullptr; t.ldr = 0;
      t.lda = lda; t.ldw = ldw; t.ldc = ldc; t.M = M; t.N = N; t.K = K;
      t.act = HICOM_ACT_NONE; t.out_dtype = c_dtype; t.rows_per_group = 1 << 30; t.group_stride_rows = 0;
      t.alpha = alpha;
      t.batch = (int)nb; t.a_batch_stride = sa; t.w_batch_stride = sb; t.c_batch_rows = sc / ldc;
      return launch_tc_linear(t, stream);
    }
  }
  for (int b1 = 0; b1 < nb1; ++b1)
    for (int b2 = 0; b2 < nb2; ++b2) {
      TcLinearParams t{};
      t.A = static_cast<const __nv_bfloat16*>(A) + b1 * sAb1 + b2 * sAb2;
      t.W = static_cast<const __nv_bfloat16*>(B) + b1 * sBb1 + b2 * sBb2;
      t.C = static_cast<char*>(C) + (size_t)(b1 * sCb1 + b2 * sCb2) * csz;
      t.bias = nullptr; t.R = nullptr; t.ldr = 0;
      t.lda = lda; t.ldw = ldw; t.ldc = ldc; t.M = M; t.N = N; t.K = K;
      t.act = HICOM_ACT_NONE; t.out_dtype = c_dtype; t.rows_per_group = 1 << 30; t.group_stride_rows = 0;
      t.alpha = alpha;
      t.w_is_kn = mode >= 1; t.a_is_km = mode == 2;
      if (launch_tc_linear(t, stream)) return 1;
    }
  return 0;
}

// ---- skinny NN: C (M <= 8 rows, fp32) = alpha * A (M x K) · B (K x N, n contiguous) -----------------------------------
// The input gradient of a linear layer evaluated on a handful of rows (dA = dpre·W of the per-video FiLM MLPs): on the
// generic SIMT tiles it is N/64 blocks walking the whole K range (latency-bound, ~290 us at 8 x 2304 x 2304).  Here
// every block streams a K range of B with coalesced 8/16-byte loads, keeps the M x 4 partial sums of its four columns
// in registers and adds them to C with fp32 atomics (C is zeroed first).
template <typename TA, typename TB>
__global__ void __launch_bounds__(256) skinny_nn_kernel(const TA* __restrict__ A, long long sAm, long long sAk,
                                                        const TB* __restrict__ B, long long sBk, float* __restrict__ C,
                                                        long long ldc, int M, int N, int K, int kc, float alpha) {
  __shared__ float As[8][64];
  const int n = (blockIdx.x * 256 + threadIdx.x) * 4;
  const int k0 = blockIdx.y * kc, k1 = min(K, k0 + kc);
  float acc[8][4];
#pragma unroll
  for (int m = 0; m < 8; ++m)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[m][e] = 0.f;
  for (int kb = k0; kb < k1; kb += 64) {
    const int kn = min(64, k1 - kb);
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 64; i += 256) {
      const int m = i >> 6, k = i & 63;
      As[m][k] = (m < M && k < kn) ? to_f32<TA>(A[m * sAm + (long long)(kb + k) * sAk]) : 0.f;
    }
    __syncthreads();
    if (n < N) {
#pragma unroll 4
      for (int k = 0; k < kn; ++k) {
        float b[4];
        Vec4<TB>::load(B + (long long)(kb + k) * sBk + n, b);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const float a = As[m][k];
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[m][e] = fmaf(a, b[e], acc[m][e]);
        }
      }
    }
  }
  if (n < N)
    for (int m = 0; m < M; ++m)
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(C + m * ldc + n + e, alpha * acc[m][e]);
}

// -1: not this kernel's problem
static int try_skinny_nn(const void* A, int64_t sAm, int64_t sAk, const void* B, int64_t sBk, int64_t sBn, void* C,
                         int64_t ldc, int M, int N, int K, int nb, float alpha, int a_dtype, int b_dtype, int c_dtype,
                         cudaStream_t s) {
  if (nb != 1 || M < 1 || M > 8 || c_dtype != HICOM_F32 || sBn != 1 || N % 4 || sBk % 4 || K < 256) return -1;
  if ((a_dtype != HICOM_F32 && a_dtype != HICOM_BF16) || (b_dtype != HICOM_F32 && b_dtype != HICOM_BF16)) return -1;
  if (reinterpret_cast<uintptr_t>(B) & 15) return -1;
  const int gx = (N + 1023) / 1024;
  int gy = 296 / gx;                       // about two blocks per SM
  int kc = ((K + gy - 1) / gy + 7) / 8 * 8;
  if (kc < 32) kc = 32;
  gy = (K + kc - 1) / kc;
  cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, s);
  HICOM_REQUIRE(e == cudaSuccess, "gemm (skinny): memset: %s", cudaGetErrorString(e));
  dim3 grid(gx, gy);
#define HICOM_SKINNY_NN(TA, TB)                                                                                     \
  skinny_nn_kernel<TA, TB><<<grid, 256, 0, s>>>(static_cast<const TA*>(A), sAm, sAk, static_cast<const TB*>(B), sBk, \
                                                 static_cast<float*>(C), ldc, M, N, K, kc, alpha)
  if (a_dtype == HICOM_F32 && b_dtype == HICOM_BF16) HICOM_SKINNY_NN(float, __nv_bfloat16);
  else if (a_dtype == HICOM_BF16 && b_dtype == HICOM_BF16) HICOM_SKINNY_NN(__nv_bfloat16, __nv_bfloat16);
  else if (a_dtype == HICOM_F32 && b_dtype == HICOM_F32) HICOM_SKINNY_NN(float, float);
  else HICOM_SKINNY_NN(__nv_bfloat16, float);
#undef HICOM_SKINNY_NN
  return check_launch("skinny_nn_kernel");
}

extern "C" int hicom_gemm(const void* A, int64_t sAm, int64_t sAk, int64_t sAb1, int64_t sAb2, const void* B,
                          int64_t sBk, int64_t sBn, int64_t sBb1, int64_t sBb2, void* C, int64_t ldc, int64_t sCb1,
                          int64_t sCb2, int M, int N, int K, int nb1, int nb2, float alpha, int a_dtype, int b_dtype,
                          int c_dtype, void* stream) {
  HICOM_REQUIRE(A && B && C, "gemm: null pointer");
  HICOM_REQUIRE(M >= 0 && N >= 0 && K >= 0 && nb1 >= 0 && nb2 >= 0, "gemm: bad shape M=%d N=%d K=%d batch=%dx%d", M, N, K,
                nb1, nb2);
  HICOM_REQUIRE(ldc >= N, "gemm: ldc too small");
  if (M == 0 || N == 0 || nb1 * nb2 == 0) return 0;
  {
    const int rc = try_skinny_nn(A, sAm, sAk, B, sBk, sBn, C, ldc, M, N, K, nb1 * nb2, alpha, a_dtype, b_dtype, c_dtype,
                                 as_stream(stream));
    if (rc >= 0) return rc;
  }
  {
    const int rc = try_gemm_tc(A, sAm, sAk, sAb1, sAb2, B, sBk, sBn, sBb1, sBb2, C, ldc, sCb1, sCb2, M, N, K, nb1, nb2,
                               alpha, a_dtype, b_dtype, c_dtype, as_stream(stream));
    if (rc >= 0) return rc;
  }
  GemmParams g{};
  g.A = A; g.B = B; g.C = C;
  g.M = M; g.N = N; g.K = K;
  g.sAm = sAm; g.sAk = sAk; g.sAb1 = sAb1; g.sAb2 = sAb2;
  g.sBk = sBk; g.sBn = sBn; g.sBb1 = sBb1; g.sBb2 = sBb2;
  g.ldc = ldc; g.sCb1 = sCb1; g.sCb2 = sCb2;
  g.nb1 = nb1; g.nb2 = nb2;
  g.alpha = alpha; g.act = HICOM_ACT_NONE;
  g.rows_per_group = 1 << 30; g.group_stride_rows = 0;
  cudaStream_t s = as_stream(stream);
  const int key = a_dtype * 100 + b_dtype * 10 + c_dtype;
  switch (key) {
    case HICOM_F32 * 100 + HICOM_F32 * 10 + HICOM_F32: return launch_gemm_simt<float, float, float>(g, s);
    case HICOM_F32 * 100 + HICOM_F32 * 10 + HICOM_BF16: return launch_gemm_simt<float, float, __nv_bfloat16>(g, s);
    case HICOM_BF16 * 100 + HICOM_BF16 * 10 + HICOM_BF16:
      return launch_gemm_simt<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(g, s);
    case HICOM_BF16 * 100 + HICOM_BF16 * 10 + HICOM_F32: return launch_gemm_simt<__nv_bfloat16, __nv_bfloat16, float>(g, s);
    case HICOM_F32 * 100 + HICOM_BF16 * 10 + HICOM_F32: return launch_gemm_simt<float, __nv_bfloat16, float>(g, s);
    default: break;
  }
  set_error("gemm: dtype combination A=%d B=%d C=%d is not built", a_dtype, b_dtype, c_dtype);
  return 1;
}

extern "C" int hicom_colsum(const void* x, int64_t ld, float* out, int64_t M, int N, int dtype, void* stream) {
  HICOM_REQUIRE(x && out, "colsum: null pointer");
  HICOM_REQUIRE(M >= 0 && N > 0 && N % 4 == 0 && ld >= N && ld % 4 == 0, "colsum: needs N and ld to be multiples of 4");
  HICOM_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "colsum: x must be 16-byte aligned");
  if (M == 0) return 0;
  long long gy = (M + 63) / 64;  // >= 8 rows per row lane
  if (gy > 64) gy = 64;
  dim3 grid((unsigned)((N + 127) / 128), (unsigned)gy);
  cudaStream_t s = as_stream(stream);
  if (dtype == HICOM_F32) colsum_kernel<float><<<grid, 256, 0, s>>>((const float*)x, ld, out, M, N);
  else if (dtype == HICOM_BF16) colsum_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, ld, out, M, N);
  else { set_error("colsum: unknown dtype code %d", dtype); return 1; }
  return check_launch("colsum_kernel");
}

extern "C" int hicom_act_backward(const void* pre, const void* dy, void* dx, int64_t n, int act, int pre_dtype,
                                  int dtype, void* stream) {
  HICOM_REQUIRE(pre && dy && dx, "act_backward: null pointer");
  HICOM_REQUIRE(n >= 0, "act_backward: bad size");
  HICOM_REQUIRE(act == HICOM_ACT_NONE || act == HICOM_ACT_GELU || act == HICOM_ACT_GELU_TANH,
                "act_backward: bad activation %d", act);
  if (n == 0) return 0;
  const long long want = (n + 255) / 256;
  const unsigned blocks = (unsigned)(want < 148 * 16 ? want : 148 * 16);
  cudaStream_t s = as_stream(stream);
  if (pre_dtype == HICOM_F32 && dtype == HICOM_F32)
    act_backward_kernel<float, float><<<blocks, 256, 0, s>>>((const float*)pre, (const float*)dy, (float*)dx, n, act);
  else if (pre_dtype == HICOM_F32 && dtype == HICOM_BF16)
    act_backward_kernel<float, __nv_bfloat16><<<blocks, 256, 0, s>>>((const float*)pre, (const __nv_bfloat16*)dy,
                                                                     (__nv_bfloat16*)dx, n, act);
  else if (pre_dtype == HICOM_BF16 && dtype == HICOM_BF16)
    act_backward_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, s>>>(
        (const __nv_bfloat16*)pre, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, n, act);
  else {
    set_error("act_backward: dtype combination pre=%d dy=%d is not built", pre_dtype, dtype);
    return 1;
  }
  return check_launch("act_backward_kernel");
}

extern "C" int hicom_softmax_backward(const float* S, const float* dP, const float* lse, const float* delta, void* dS,
                                      int B, int64_t N, int J, int out_dtype, void* stream) {
  HICOM_REQUIRE(S && dP && lse && delta && dS, "softmax_backward: null pointer");
  HICOM_REQUIRE(B >= 0 && N >= 0 && J > 0, "softmax_backward: bad shape");
  const long long total = (long long)B * N * J;
  if (total == 0) return 0;
  const long long want = (total + 255) / 256;
  const unsigned blocks = (unsigned)(want < 148 * 16 ? want : 148 * 16);
  cudaStream_t s = as_stream(stream);
  if (out_dtype == HICOM_F32)
    softmax_backward_kernel<float><<<blocks, 256, 0, s>>>(S, dP, lse, delta, (float*)dS, N * J, J, total);
  else if (out_dtype == HICOM_BF16)
    softmax_backward_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(S, dP, lse, delta, (__nv_bfloat16*)dS, N * J, J, total);
  else {
    set_error("softmax_backward: unknown dtype code %d", out_dtype);
    return 1;
  }
  return check_launch("softmax_backward_kernel");
}

template <typename T>
static int launch_local_bwd_q(const void* K, const void* V, const void* Q, const void* dO, void* dQ, float* dK, float* dV,
                              int B, int T_, int H, int W, int d, const AxisWin& at, const AxisWin& ah, const AxisWin& aw, float scale,
                              int k_l2norm, cudaStream_t s) {
  const long long wins = (long long)B * at.count * ah.count * aw.count;
  if (wins == 0) return 0;
  const long long blocks = (wins + 7) / 8;
  HICOM_REQUIRE(blocks < (1ll << 31), "local_attend_backward: too many windows");
#define HICOM_LBQ(CPL)                                                                                              \
  local_attend_bwd_q_kernel<T, CPL><<<(unsigned)blocks, 256, 0, s>>>((const T*)K, (const T*)V, (const T*)Q,          \
                                                                     (const T*)dO, (T*)dQ, dK, dV, B, T_, H, W, d,  \
                                                                     at, ah, aw, scale, k_l2norm)
  switch (d / 128) {
    case 9: HICOM_LBQ(9); break;
    case 6: HICOM_LBQ(6); break;
    case 1: HICOM_LBQ(1); break;
    default: set_error("local_attend_backward: d=%d unsupported", d); return 1;
  }
#undef HICOM_LBQ
  return check_launch("local_attend_bwd_kernel");
}

extern "C" int hicom_local_attend_backward(const void* Ksrc, const void* Vsrc, const void* Q, const void* dO, void* dQ,
                                           float* dK, float* dV, int B, int T, int H, int W, int d, int kt, int ks,
                                           float logit_scale, int k_l2norm, int dtype, void* stream) {
  HICOM_REQUIRE(Ksrc && Vsrc && Q && dO, "local_attend_backward: null pointer");
  HICOM_REQUIRE(dQ || dK || dV, "local_attend_backward: no output requested");
  HICOM_REQUIRE(B >= 0 && T > 0 && H > 0 && W > 0 && d > 0 && d % 128 == 0 && kt > 0 && ks > 0,
                "local_attend_backward: bad shape");
  HICOM_REQUIRE(!(dK && k_l2norm), "local_attend_backward: key gradients through the L2 normalisation are not built");
  AxisWin at, ah, aw;
  HICOM_REQUIRE(make_axis_win(T, kt, &at) && make_axis_win(H, ks, &ah) && make_axis_win(W, ks, &aw),
                "local_attend_backward: the reference cannot stack these windows (T=%d H=%d W=%d, kernel %d/%d)", T, H, W,
                kt, ks);
  HICOM_DISPATCH_DTYPE(dtype, E, return (launch_local_bwd_q<E>(Ksrc, Vsrc, Q, dO, dQ, dK, dV, B, T, H, W, d, at, ah, aw,
                                                                logit_scale, k_l2norm, as_stream(stream))));
}

// ---- grid pooling backward (projector.py:539-540; gradients into frames_feature, mm_tunable_parts 'pure_vision_model') --
// dX[b, tap] += w_tap * dQ[b, window] over the <= 8 trilinear taps of upsample_trilinear3d (align_corners=False).
// One warp per window, lanes over channels; fp32 atomics (neighbouring windows share taps); dX is zeroed by the caller.
template <typename T>
__global__ void __launch_bounds__(256) grid_pool_backward_kernel(const T* __restrict__ dQ, float* __restrict__ dX,
                                                                 long long total, int T_, int H, int W, int d, int t1n,
                                                                 int h1n, int w1n) {
  const int lane = threadIdx.x & 31;
  const long long wi = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wi >= total) return;
  const long long nwin = (long long)t1n * h1n * w1n;
  const int b = (int)(wi / nwin), r = (int)(wi % nwin);
  const int t1 = r / (h1n * w1n), h1 = (r / w1n) % h1n, w1 = r % w1n;
  const Tap tt = linear_tap(t1, T_, t1n), th = linear_tap(h1, H, h1n), tw = linear_tap(w1, W, w1n);
  const int ti[2] = {tt.i0, tt.i1}; const float twt[2] = {tt.w0, tt.w1};
  const int hi[2] = {th.i0, th.i1}; const float hwt[2] = {th.w0, th.w1};
  const int wj[2] = {tw.i0, tw.i1}; const float wwt[2] = {tw.w0, tw.w1};
  const T* q = dQ + wi * d;
  float* base = dX + (size_t)b * T_ * H * W * d;
  for (int c = lane * 4; c < d; c += 128) {
    float g[4];
    Vec4<T>::load(q + c, g);
    for (int a = 0; a < 2; ++a) {
      if (twt[a] == 0.f) continue;
      for (int bb = 0; bb < 2; ++bb) {
        if (hwt[bb] == 0.f) continue;
        for (int cc = 0; cc < 2; ++cc) {
          if (wwt[cc] == 0.f) continue;
          const float wgt = twt[a] * (hwt[bb] * wwt[cc]);
          float* dst = base + ((size_t)(ti[a] * H + hi[bb]) * W + wj[cc]) * d + c;
#pragma unroll
          for (int e = 0; e < 4; ++e) atomicAdd(dst + e, wgt * g[e]);
        }
      }
    }
  }
}

extern "C" int hicom_grid_pool_backward(const void* dQ, float* dX, int B, int T, int H, int W, int d, int kt, int ks,
                                        int dtype, void* stream) {
  HICOM_REQUIRE(dQ && dX, "grid_pool_backward: null pointer");
  HICOM_REQUIRE(B >= 0 && T > 0 && H > 0 && W > 0 && d > 0 && d % 4 == 0 && kt > 0 && ks > 0, "grid_pool_backward: bad shape");
  const int t1n = ceil_div(T, kt), h1n = ceil_div(H, ks), w1n = ceil_div(W, ks);
  const long long total = (long long)B * t1n * h1n * w1n;
  if (total == 0) return 0;
  const long long blocks = (total + 7) / 8;
  HICOM_REQUIRE(blocks < (1ll << 31), "grid_pool_backward: too many windows");
  HICOM_DISPATCH_DTYPE(dtype, E, (grid_pool_backward_kernel<E><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      static_cast<const E*>(dQ), dX, total, T, H, W, d, t1n, h1n, w1n)));
  return check_launch("grid_pool_backward_kernel");
}

// ---- L2 row normalisation backward (use_clip_scale, projector.py:184-188,527-529): y = x / |x|,
//      dx = (dy - y (y·dy)) / |x|.  One warp per row.
template <typename T>
__global__ void __launch_bounds__(256) l2norm_rows_backward_kernel(const T* __restrict__ X, const T* __restrict__ dY,
                                                                   T* __restrict__ dX, long long rows, int d) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* x = X + row * d;
  const T* g = dY + row * d;
  float ss = 0.f, dot = 0.f;
  for (int c = lane * 4; c < d; c += 128) {
    float v[4], u[4];
    Vec4<T>::load(x + c, v);
    Vec4<T>::load(g + c, u);
#pragma unroll
    for (int e = 0; e < 4; ++e) { ss = fmaf(v[e], v[e], ss); dot = fmaf(v[e], u[e], dot); }
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float inv = rsqrtf(ss), k = dot / ss;
  for (int c = lane * 4; c < d; c += 128) {
    float v[4], u[4];
    Vec4<T>::load(x + c, v);
    Vec4<T>::load(g + c, u);
#pragma unroll
    for (int e = 0; e < 4; ++e) u[e] = (u[e] - v[e] * k) * inv;
    Vec4<T>::store(dX + row * d + c, u);
  }
}

extern "C" int hicom_l2norm_rows_backward(const void* X, const void* dY, void* dX, long long rows, int d, int dtype,
                                          void* stream) {
  HICOM_REQUIRE(X && dY && dX, "l2norm_rows_backward: null pointer");
  HICOM_REQUIRE(rows >= 0 && d > 0 && d % 4 == 0, "l2norm_rows_backward: bad shape");
  if (rows == 0) return 0;
  const long long blocks = (rows + 7) / 8;
  HICOM_REQUIRE(blocks < (1ll << 31), "l2norm_rows_backward: too many rows");
  HICOM_DISPATCH_DTYPE(dtype, E, (l2norm_rows_backward_kernel<E><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      static_cast<const E*>(X), static_cast<const E*>(dY), static_cast<E*>(dX), rows, d)));
  return check_launch("l2norm_rows_backward_kernel");
}

template <typename T>
static int launch_film_ln_bwd(const void* x, const float* film, const void* w, const void* dy, void* dx, float* dfilm,
                              float* dw, float* dbias, long long rows, int d, int rows_per_group, cudaStream_t s) {
  if (rows == 0) return 0;
  // blocks of 8 warps over contiguous row ranges; ranges never straddle more groups than necessary
  long long blocks = (rows + 63) / 64;
  if (blocks > 148 * 4) blocks = 148 * 4;
#define HICOM_FLB(CPL)                                                                                          \
  film_ln_backward_kernel<T, CPL><<<(unsigned)blocks, 256, 0, s>>>((const T*)x, film, (const T*)w, (const T*)dy, \
                                                                   (T*)dx, dfilm, dw, dbias, rows, d, rows_per_group)
  switch (d / 128) {
    case 9: HICOM_FLB(9); break;
    case 6: HICOM_FLB(6); break;
    case 1: HICOM_FLB(1); break;
    default: set_error("film_layernorm_backward: d=%d unsupported", d); return 1;
  }
#undef HICOM_FLB
  return check_launch("film_ln_backward_kernel");
}

extern "C" int hicom_film_layernorm_backward(const void* x, const float* film, const void* ln_w, const void* dy,
                                             void* dx, float* dfilm, float* dw, float* dbias, int64_t rows, int d,
                                             int rows_per_group, int dtype, void* stream) {
  HICOM_REQUIRE(x && film && ln_w && dy && dfilm && dw && dbias, "film_layernorm_backward: null pointer");
  HICOM_REQUIRE(rows >= 0 && d > 0 && d % 128 == 0 && rows_per_group > 0, "film_layernorm_backward: bad shape");
  HICOM_DISPATCH_DTYPE(dtype, E, return (launch_film_ln_bwd<E>(x, film, ln_w, dy, dx, dfilm, dw, dbias, rows, d,
                                                                rows_per_group, as_stream(stream))));
}

template <typename T>
static int launch_mix_ln_bwd(const void* x, const void* y, const void* w, const void* bias, const void* alpha,
                             const void* dout, void* dx, void* dy, float* dw, float* dbias, float* dalpha, long long rows,
                             int d, cudaStream_t s) {
  if (rows == 0) return 0;
  long long blocks = (rows + 63) / 64;
  if (blocks > 148 * 4) blocks = 148 * 4;
#define HICOM_MLB(CPL)                                                                                              \
  mix_ln_backward_kernel<T, CPL><<<(unsigned)blocks, 256, 0, s>>>((const T*)x, (const T*)y, (const T*)w, (const T*)bias, \
                                                                  (const T*)alpha, (const T*)dout, (T*)dx, (T*)dy, dw,  \
                                                                  dbias, dalpha, rows, d)
  switch (d / 128) {
    case 9: HICOM_MLB(9); break;
    case 6: HICOM_MLB(6); break;
    case 1: HICOM_MLB(1); break;
    default: set_error("mix_layernorm_backward: d=%d unsupported", d); return 1;
  }
#undef HICOM_MLB
  return check_launch("mix_ln_backward_kernel");
}

extern "C" int hicom_mix_layernorm_backward(const void* x, const void* y, const void* ln_w, const void* ln_b,
                                            const void* alpha, const void* dout, void* dx, void* dy, float* dw,
                                            float* dbias, float* dalpha, int64_t rows, int d, int dtype, void* stream) {
  HICOM_REQUIRE(x && y && ln_w && ln_b && alpha && dout && dy && dw && dbias && dalpha,
                "mix_layernorm_backward: null pointer");
  HICOM_REQUIRE(rows >= 0 && d > 0 && d % 128 == 0, "mix_layernorm_backward: bad shape");
  HICOM_DISPATCH_DTYPE(dtype, E, return (launch_mix_ln_bwd<E>(x, y, ln_w, ln_b, alpha, dout, dx, dy, dw, dbias, dalpha,
                                                               rows, d, as_stream(stream))));
}
