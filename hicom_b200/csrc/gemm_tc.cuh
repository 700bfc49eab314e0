// tcgen05 / TMEM / TMA tensor-core paths (bf16 or fp16 in, fp32 accumulate).
#pragma once
#include "common.cuh"

namespace hicom {

struct TcLinearParams {
  const void* A; const void* W; const void* bias; const void* R; void* C;
  long long lda, ldw, ldr, ldc;
  int M, N, K, act, out_dtype;
  int rows_per_group; long long group_stride_rows;
  float alpha = 1.f;        // C = act(alpha * A·op(W) + bias) + R
  int w_is_kn = 0;          // W stored (K, N) row-major instead of torch's (N, K)
  int diag_heads = 0, diag_rows = 0, diag_cols = 0;  // block-diagonal store (see gemm_tc.cu)
  int z_slices = 0, z_a_k = 0, z_b_k = 0; long long z_c_rows = 0;  // K slices on the tile z axis (one head / one range)
  int z_c_cols = 0; long long k_total = 0;                           // per-slice output column shift; true K extent
  int accumulate = 0;       // C += result (fp32 C, w_is_kn, no activation)
  int a_is_km = 0;          // A stored (K, M) row-major per batch entry (with w_is_kn): batched, K-sliced, fp32 C
  int batch = 0; long long a_batch_stride = 0, c_batch_rows = 0;
  long long w_batch_stride = 0;  // a_is_km only: != 0 -> W has a batch axis too (dqfold[b] = dS[b]ᵀ·x'[b]); 0 -> shared
  const int* guard = nullptr;  // device flag: the launch is a no-op unless *guard != 0
  // 16-bit operand formats: 0 = bf16, 1 = fp16 (tcgen05.mma.kind::f16 takes either, per operand).  bias and R follow A.
  int a_f16 = 0, w_f16 = 0;
};

bool tc_linear_supported(int in_dtype, int out_dtype, int M, int N, int K, long long lda, long long ldw,
                         long long ldc, const void* A, const void* W, const void* C);
int launch_tc_linear(const TcLinearParams& p, cudaStream_t stream);

bool tc_global_selected(int dtype, int impl, int d, int J, int T, int H, int W);
size_t tc_global_workspace_bytes(int B, int T, int H, int W, int d, int J, int splits);
int launch_tc_global(const void* X, const void* Kscore, const float* pos_t, const float* pos_h, const float* pos_w,
                     const void* qfold, float* m, float* l, float* o, int B, int T, int H, int W, int d, int J,
                     int splits, void* workspace, cudaStream_t stream, bool f16 = false);

}  // namespace hicom
