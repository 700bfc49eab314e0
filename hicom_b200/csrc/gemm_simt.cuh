// Generic strided, two-level-batched SIMT GEMM (fp32 FMA pipe, fp32 accumulation).
//
// It is (a) the fp32-mode path of every dense contraction (fp32 parity needs <= 1e-4, which the
// bf16/tf32 tensor pipe cannot give), (b) the path for skinny problems (a handful of rows) where a
// 128-row tcgen05 tile would be mostly padding, and (c) the in-tree cross-check for the tcgen05
// kernels.  C = act(alpha * A·B + bias) + R with arbitrary element strides on A and B.
#pragma once
#include "common.cuh"

namespace hicom {

struct GemmParams {
  const void* A; const void* B; const void* bias; const void* R; void* C;
  int M, N, K;
  long long sAm, sAk, sAb1, sAb2;
  long long sBk, sBn, sBb1, sBb2;
  long long ldc, sCb1, sCb2;
  long long ldr, sRb1, sRb2;
  long long sBiasb2;         // bias offset per inner batch index (elements)
  int nb1, nb2;              // batch = nb1 * nb2 (blockIdx.z = b1*nb2 + b2)
  int k_total2;              // if > 0: K_eff = min(K, k_total2 - b2*K)  (ragged split-K ranges)
  float alpha;
  int act;
  int rows_per_group; long long group_stride_rows;  // output row remap (see hicom_linear)
};

// TA: storage type of A; TB: storage type of B, bias, R; TC: storage type of C.
template <typename TA, typename TB, typename TC>
int launch_gemm_simt(const GemmParams& p, cudaStream_t stream);

}  // namespace hicom
