#include "gemm_simt.cuh"

namespace hicom {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4, NT = 256;

template <typename TA, typename TB, typename TC, bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int b1 = blockIdx.z / p.nb2, b2 = blockIdx.z % p.nb2;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int K = p.K;
  if (p.k_total2 > 0) {
    int rem = p.k_total2 - b2 * p.K;
    K = rem < K ? rem : K;
    if (K < 0) K = 0;
  }
  const TA* __restrict__ A = static_cast<const TA*>(p.A) + b1 * p.sAb1 + b2 * p.sAb2;
  const TB* __restrict__ B = static_cast<const TB*>(p.B) + b1 * p.sBb1 + b2 * p.sBb2;

  // element -> (row, k) maps chosen so that consecutive threads walk the contiguous axis
  int a_m[4], a_k[4], b_n[4], b_k[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = tid + NT * i;
    if (A_KCONTIG) { a_k[i] = e % BK; a_m[i] = e / BK; } else { a_m[i] = e % BM; a_k[i] = e / BM; }
    if (B_KCONTIG) { b_k[i] = e % BK; b_n[i] = e / BK; } else { b_n[i] = e % BN; b_k[i] = e / BN; }
  }
  float ra[4], rb[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + a_m[i], k = k0 + a_k[i];
      ra[i] = (m < p.M && k < K) ? to_f32<TA>(A[m * p.sAm + k * p.sAk]) : 0.f;
      const int n = n0 + b_n[i], kb = k0 + b_k[i];
      rb[i] = (n < p.N && kb < K) ? to_f32<TB>(B[kb * p.sBk + n * p.sBn]) : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[buf][a_k[i]][a_m[i]] = ra[i];
      Bs[buf][b_k[i]][b_n[i]] = rb[i];
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ktiles = (K + BK - 1) / BK;
  if (ktiles > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) gload((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < ktiles) sstore(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue ---------------------------------------------------------------------------
  const TB* bias = p.bias ? static_cast<const TB*>(p.bias) + b2 * p.sBiasb2 : nullptr;
  const TB* R = p.R ? static_cast<const TB*>(p.R) + b1 * p.sRb1 + b2 * p.sRb2 : nullptr;
  TC* C = static_cast<TC*>(p.C) + b1 * p.sCb1 + b2 * p.sCb2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const long long orow = (long long)(m / p.rows_per_group) * p.group_stride_rows + (m % p.rows_per_group);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (bias) v += to_f32<TB>(bias[n]);
      v = apply_act(v, p.act);
      if (R) v += to_f32<TB>(R[m * p.ldr + n]);
      C[orow * p.ldc + n] = from_f32<TC>(v);
    }
  }
}

template <typename TA, typename TB, typename TC>
int launch_gemm_simt(const GemmParams& p, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || p.nb1 * p.nb2 <= 0) return 0;
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.nb1 * p.nb2);
  HICOM_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm_simt: grid too large (M=%d batch=%d)", p.M, p.nb1 * p.nb2);
  const bool ak = (p.sAk == 1), bk = (p.sBk == 1);
  if (ak && bk) gemm_simt_kernel<TA, TB, TC, true, true><<<grid, NT, 0, stream>>>(p);
  else if (ak && !bk) gemm_simt_kernel<TA, TB, TC, true, false><<<grid, NT, 0, stream>>>(p);
  else if (!ak && bk) gemm_simt_kernel<TA, TB, TC, false, true><<<grid, NT, 0, stream>>>(p);
  else gemm_simt_kernel<TA, TB, TC, false, false><<<grid, NT, 0, stream>>>(p);
  return check_launch("gemm_simt_kernel");
}

template int launch_gemm_simt<float, float, float>(const GemmParams&, cudaStream_t);
template int launch_gemm_simt<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(const GemmParams&, cudaStream_t);
template int launch_gemm_simt<__nv_bfloat16, __nv_bfloat16, float>(const GemmParams&, cudaStream_t);
template int launch_gemm_simt<float, float, __nv_bfloat16>(const GemmParams&, cudaStream_t);
template int launch_gemm_simt<float, __nv_bfloat16, float>(const GemmParams&, cudaStream_t);
template int launch_gemm_simt<__half, __half, __half>(const GemmParams&, cudaStream_t);
template int launch_gemm_simt<__half, __half, float>(const GemmParams&, cudaStream_t);
template int launch_gemm_simt<float, __half, float>(const GemmParams&, cudaStream_t);

}  // namespace hicom
