// Skinny linear layer: C = act(alpha * A·Wᵀ + bias) + R for M <= 8 rows (one instruction vector per video: the FiLM
// MLPs at small batch).  A 128-row tensor-core tile would be > 75 % padding and only
// N/256 CTAs would stream the weights; here every warp owns output columns, streams W[n,:] once with 16-byte loads and
// keeps the M x K activations in shared memory, so all SMs pull weights at HBM speed.  fp32 accumulation.
#include <stdlib.h>

#include "common.cuh"

namespace hicom {

struct SkinnyParams {
  const void* A; const void* W; const void* bias; const void* R; void* C;
  long long lda, ldw, ldr, ldc;
  int M, N, K, act;
  float alpha;
  int rows_per_group; long long group_stride_rows;
};

template <typename T> struct Elems16;  // elements per 16-byte load
template <> struct Elems16<float> { static constexpr int n = 4; };
template <> struct Elems16<__nv_bfloat16> { static constexpr int n = 8; };
template <> struct Elems16<__half> { static constexpr int n = 8; };

template <typename T>
__device__ __forceinline__ void load16(const T* p, float (&v)[Elems16<T>::n]);
template <>
__device__ __forceinline__ void load16<float>(const float* p, float (&v)[4]) {
  const float4 r = *reinterpret_cast<const float4*>(p);
  v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
template <>
__device__ __forceinline__ void load16<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = __uint_as_float(w[e] << 16);
    v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}

template <>
__device__ __forceinline__ void load16<__half>(const __half* p, float (&v)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
    v[2 * e] = f.x;
    v[2 * e + 1] = f.y;
  }
}

template <typename TI, typename TO, int MMAX>
__global__ void __launch_bounds__(256) skinny_linear_kernel(const SkinnyParams p) {
  extern __shared__ __align__(16) uint8_t sk_smem[];
  TI* As = reinterpret_cast<TI*>(sk_smem);  // (M, Kp) row-major, Kp = K rounded up to 16-byte chunks
  constexpr int E = Elems16<TI>::n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int K = p.K, M = p.M;
  const TI* A = static_cast<const TI*>(p.A);
  for (int idx = threadIdx.x * E; idx < M * K; idx += blockDim.x * E) {
    const int m = idx / K, k = idx % K;  // K % E == 0, so a chunk never crosses a row
    *reinterpret_cast<uint4*>(As + (size_t)m * K + k) = *reinterpret_cast<const uint4*>(A + (size_t)m * p.lda + k);
  }
  __syncthreads();

  const TI* W = static_cast<const TI*>(p.W);
  // each warp owns two output columns at a time (twice the weight loads in flight)
  for (int n0 = (blockIdx.x * nwarps + warp) * 2; n0 < p.N; n0 += gridDim.x * nwarps * 2) {
    const bool two = n0 + 1 < p.N;
    float acc[2][MMAX];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int m = 0; m < MMAX; ++m) acc[c][m] = 0.f;
    const TI* w0 = W + (size_t)n0 * p.ldw;
    const TI* w1 = W + (size_t)(two ? n0 + 1 : n0) * p.ldw;
#pragma unroll 3
    for (int k = lane * E; k < K; k += 32 * E) {
      float wa[E], wb[E];
      load16<TI>(w0 + k, wa);
      load16<TI>(w1 + k, wb);
#pragma unroll
      for (int m = 0; m < MMAX; ++m) {
        if (m < M) {
          float a[E];
          load16<TI>(As + (size_t)m * K + k, a);
#pragma unroll
          for (int e = 0; e < E; ++e) {
            acc[0][m] = fmaf(a[e], wa[e], acc[0][m]);
            acc[1][m] = fmaf(a[e], wb[e], acc[1][m]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (c == 1 && !two) break;
      const int n = n0 + c;
      // reduce every accumulator over the warp; lane m ends up owning row m
      float mine = 0.f;
#pragma unroll
      for (int m = 0; m < MMAX; ++m) {
        const float sum = warp_sum(acc[c][m]);
        if (lane == m) mine = sum;
      }
      if (lane < M) {
        float v = mine * p.alpha;
        if (p.bias) v += to_f32<TI>(static_cast<const TI*>(p.bias)[n]);
        v = apply_act(v, p.act);
        if (p.R) v += to_f32<TI>(static_cast<const TI*>(p.R)[(size_t)lane * p.ldr + n]);
        const long long orow = (long long)(lane / p.rows_per_group) * p.group_stride_rows + (lane % p.rows_per_group);
        static_cast<TO*>(p.C)[orow * p.ldc + n] = from_f32<TO>(v);
      }
    }
  }
}

constexpr int kSkinnyMaxRows = 8;  // beyond this the 128-row tensor tile is faster (measured)

bool skinny_supported(int in_dtype, int M, int N, int K, long long lda, long long ldw, const void* A, const void* W) {
  (void)N;
  const size_t es = in_dtype == HICOM_F32 ? 4 : 2;
  const int e16 = 16 / (int)es;
  if (M < 1 || M > kSkinnyMaxRows) return false;
  if (K % e16 || lda % e16 || ldw % e16) return false;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) & 15) return false;
  return (size_t)M * K * es <= 160 * 1024;
}

template <typename TI, typename TO, int MMAX>
static int launch_skinny_m(const SkinnyParams& p, cudaStream_t stream) {
  const size_t smem = (size_t)p.M * p.K * sizeof(TI);
  auto kern = skinny_linear_kernel<TI, TO, MMAX>;
  static bool configured = false;  // one flag per kernel instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    HICOM_REQUIRE(e == cudaSuccess, "skinny_linear: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  int blocks = (p.N + 15) / 16;  // 8 warps x 2 columns per block
  if (blocks > num_sms) blocks = num_sms;
  KernelTimer timer("skinny_linear", stream);
  kern<<<blocks, 256, smem, stream>>>(p);
  return check_launch("skinny_linear_kernel");
}

template <typename TI, typename TO>
static int launch_skinny_t(const SkinnyParams& p, cudaStream_t stream) {
  if (p.M <= 2) return launch_skinny_m<TI, TO, 2>(p, stream);
  if (p.M <= 8) return launch_skinny_m<TI, TO, 8>(p, stream);
  return launch_skinny_m<TI, TO, 32>(p, stream);
}

int launch_skinny(const SkinnyParams& p, int in_dtype, int out_dtype, cudaStream_t stream) {
  if (in_dtype == HICOM_BF16 && out_dtype == HICOM_BF16) return launch_skinny_t<__nv_bfloat16, __nv_bfloat16>(p, stream);
  if (in_dtype == HICOM_BF16 && out_dtype == HICOM_F32) return launch_skinny_t<__nv_bfloat16, float>(p, stream);
  if (in_dtype == HICOM_F32 && out_dtype == HICOM_F32) return launch_skinny_t<float, float>(p, stream);
  if (in_dtype == HICOM_F16 && out_dtype == HICOM_F16) return launch_skinny_t<__half, __half>(p, stream);
  if (in_dtype == HICOM_F16 && out_dtype == HICOM_F32) return launch_skinny_t<__half, float>(p, stream);
  set_error("skinny_linear: unsupported dtype combination %d/%d", in_dtype, out_dtype);
  return 1;
}

}  // namespace hicom
