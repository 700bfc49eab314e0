// C-ABI entry points that are compositions of kernels: library info, hicom_linear, and the global
// compressor ops.  See include/hicom_b200.h for the contract of each.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

namespace hicom {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;  // kernels enqueued by this library (all threads)

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int launch_posadd(const void* X, void* Y, const float* pt, const float* ph, const float* pw, int B, int T_,
                  int H, int W, int d, int dtype, cudaStream_t stream);
int launch_col_softmax(float* S, float* m, float* l, int B, int N, int J, int splits, int rows_per_split,
                       cudaStream_t stream);

// ---- optional per-kernel timing -------------------------------------------------------------------
struct TimedLaunch { char label[96]; cudaEvent_t a, b; };
static bool g_timing = false;
static std::vector<TimedLaunch> g_timed;
bool kernel_timing_enabled() { return g_timing; }
void kernel_timing_begin(const char* label, cudaStream_t stream) {
  TimedLaunch t;
  snprintf(t.label, sizeof(t.label), "%s", label);
  cudaEventCreate(&t.a);
  cudaEventCreate(&t.b);
  cudaEventRecord(t.a, stream);
  g_timed.push_back(t);
}
void kernel_timing_end(cudaStream_t stream) {
  if (!g_timed.empty()) cudaEventRecord(g_timed.back().b, stream);
}

struct SkinnyParams {
  const void* A; const void* W; const void* bias; const void* R; void* C;
  long long lda, ldw, ldr, ldc;
  int M, N, K, act;
  float alpha;
  int rows_per_group; long long group_stride_rows;
};
bool skinny_supported(int in_dtype, int M, int N, int K, long long lda, long long ldw, const void* A, const void* W);
int launch_skinny(const SkinnyParams& p, int in_dtype, int out_dtype, cudaStream_t stream);

static GemmParams plain_gemm() {
  GemmParams g{};
  g.nb1 = g.nb2 = 1;
  g.alpha = 1.f;
  g.act = HICOM_ACT_NONE;
  g.rows_per_group = 1 << 30;
  g.group_stride_rows = 0;
  return g;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static thread_local int g_sm_limit = 0;
int sm_budget() {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  return (g_sm_limit > 0 && g_sm_limit < num_sms) ? g_sm_limit : num_sms;
}
bool sm_limited() { return g_sm_limit > 0 && sm_budget() == g_sm_limit; }

}  // namespace hicom

using namespace hicom;

extern "C" int hicom_abi_version(void) { return HICOM_ABI_VERSION; }
extern "C" const char* hicom_last_error(void) { return g_err; }
extern "C" uint64_t hicom_kernel_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int hicom_kernel_timing_enable(int on) {
  g_timing = on != 0;
  if (!g_timing) {
    for (auto& t : g_timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    g_timed.clear();
  }
  return 0;
}

// Writes "label\tcount\ttotal_ms\n" lines into buf (synchronises the device); returns bytes needed.
extern "C" size_t hicom_kernel_timing_collect(char* buf, size_t cap) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<int, double>> agg;
  for (auto& t : g_timed) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
      auto& e = agg[t.label];
      e.first += 1;
      e.second += ms;
    }
  }
  std::string out;
  for (auto& kv : agg) {
    char line[192];
    snprintf(line, sizeof(line), "%s\t%d\t%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (buf && cap > 0) {
    size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return out.size() + 1;
}

extern "C" int hicom_set_sm_limit(int sms) {
  const int old = g_sm_limit;
  g_sm_limit = sms > 0 ? (sms < 2 ? 2 : sms & ~1) : 0;
  return old;
}

extern "C" int hicom_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  HICOM_REQUIRE(e == cudaSuccess, "device_info: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  HICOM_REQUIRE(e == cudaSuccess, "device_info: %s", cudaGetErrorString(e));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  HICOM_REQUIRE(prop.major == 10, "hicom_b200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
  return 0;
}

extern "C" int hicom_linear(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias,
                            const void* R, int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int act,
                            int in_dtype, int out_dtype, int rows_per_group, int64_t group_stride_rows,
                            int impl, void* stream) {
  HICOM_REQUIRE(A && W && C, "linear: null pointer");
  HICOM_REQUIRE(M >= 0 && N > 0 && K > 0, "linear: bad shape M=%d N=%d K=%d", M, N, K);
  HICOM_REQUIRE(lda >= K && ldw >= K && ldc >= N, "linear: leading dimension too small");
  HICOM_REQUIRE(rows_per_group > 0, "linear: rows_per_group must be positive");
  HICOM_REQUIRE(act == HICOM_ACT_NONE || act == HICOM_ACT_GELU || act == HICOM_ACT_GELU_TANH, "linear: bad activation %d", act);
  if (M == 0) return 0;
  // skinny problems (M <= 32): warp-per-column kernel, all SMs stream the weights (see skinny.cu)
  if (impl == HICOM_IMPL_AUTO && !(in_dtype == HICOM_F32 && out_dtype == HICOM_BF16) &&
      skinny_supported(in_dtype, M, N, K, lda, ldw, A, W)) {
    SkinnyParams k{};
    k.A = A; k.W = W; k.bias = bias; k.R = R; k.C = C; k.lda = lda; k.ldw = ldw; k.ldr = ldr; k.ldc = ldc;
    k.M = M; k.N = N; k.K = K; k.act = act; k.alpha = 1.f;
    k.rows_per_group = rows_per_group; k.group_stride_rows = group_stride_rows;
    return launch_skinny(k, in_dtype, out_dtype, as_stream(stream));
  }
  const bool tc_ok = tc_linear_supported(in_dtype, out_dtype, M, N, K, lda, ldw, ldc, A, W, C);
  if (impl == HICOM_IMPL_TCGEN05) HICOM_REQUIRE(tc_ok, "linear: tcgen05 path does not support this problem");
  if (impl == HICOM_IMPL_TCGEN05 || (impl == HICOM_IMPL_AUTO && tc_ok)) {
    TcLinearParams t{};
    t.A = A; t.W = W; t.bias = bias; t.R = R; t.C = C;
    t.lda = lda; t.ldw = ldw; t.ldr = ldr; t.ldc = ldc; t.M = M; t.N = N; t.K = K; t.act = act;
    t.out_dtype = out_dtype; t.rows_per_group = rows_per_group; t.group_stride_rows = group_stride_rows;
    t.a_f16 = t.w_f16 = in_dtype == HICOM_F16;
    return launch_tc_linear(t, as_stream(stream));
  }
  GemmParams g = plain_gemm();
  g.A = A; g.B = W; g.bias = bias; g.R = R; g.C = C;
  g.M = M; g.N = N; g.K = K;
  g.sAm = lda; g.sAk = 1;
  g.sBk = 1; g.sBn = ldw;
  g.ldc = ldc; g.ldr = ldr;
  g.act = act; g.rows_per_group = rows_per_group; g.group_stride_rows = group_stride_rows;
  cudaStream_t s = as_stream(stream);
  if (in_dtype == HICOM_F32 && out_dtype == HICOM_F32) return launch_gemm_simt<float, float, float>(g, s);
  if (in_dtype == HICOM_BF16 && out_dtype == HICOM_BF16) return launch_gemm_simt<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(g, s);
  if (in_dtype == HICOM_BF16 && out_dtype == HICOM_F32) return launch_gemm_simt<__nv_bfloat16, __nv_bfloat16, float>(g, s);
  if (in_dtype == HICOM_F32 && out_dtype == HICOM_BF16) return launch_gemm_simt<float, float, __nv_bfloat16>(g, s);
  if (in_dtype == HICOM_F16 && out_dtype == HICOM_F16) return launch_gemm_simt<__half, __half, __half>(g, s);
  if (in_dtype == HICOM_F16 && out_dtype == HICOM_F32) return launch_gemm_simt<__half, __half, float>(g, s);
  set_error("linear: bad dtype codes %d/%d", in_dtype, out_dtype);
  return 1;
}

extern "C" int hicom_global_fold_query(const void* q, const void* Wk, void* qfold, int B, int Q, int d,
                                       int heads, float alpha, int dtype, void* stream) {
  HICOM_REQUIRE(q && Wk && qfold, "global_fold_query: null pointer");
  HICOM_REQUIRE(B >= 0 && Q > 0 && heads > 0 && d % heads == 0, "global_fold_query: bad shape");
  if (B == 0) return 0;
  const int hd = d / heads;
  if ((dtype == HICOM_BF16 || dtype == HICOM_F16) && hd % 64 == 0 && d % 8 == 0 && !((uintptr_t)q & 15) &&
      !((uintptr_t)Wk & 15)) {
    // one tcgen05 launch, blockIdx.z = head: A = q (B*Q, d) K-slice of the head, B operand = Wk (k rows, n
    // contiguous -> MN-major), rows (b,i) land at qfold row b*J + h*Q + i
    TcLinearParams t{};
    t.A = q; t.W = Wk; t.C = qfold;
    t.lda = d; t.ldw = d; t.ldc = d; t.M = B * Q; t.N = d; t.K = hd; t.act = HICOM_ACT_NONE;
    t.out_dtype = dtype; t.a_f16 = t.w_f16 = dtype == HICOM_F16;
    t.rows_per_group = Q; t.group_stride_rows = (long long)heads * Q;
    t.alpha = alpha; t.w_is_kn = 1;
    t.z_slices = heads; t.z_a_k = hd; t.z_b_k = hd; t.z_c_rows = Q;
    return launch_tc_linear(t, as_stream(stream));
  }
  GemmParams g = plain_gemm();
  g.A = q; g.B = Wk; g.C = qfold;
  g.M = Q; g.N = d; g.K = hd;
  g.sAm = d; g.sAk = 1; g.sAb1 = (long long)Q * d; g.sAb2 = hd;
  g.sBk = d; g.sBn = 1; g.sBb1 = 0; g.sBb2 = (long long)hd * d;
  g.ldc = d; g.sCb1 = (long long)heads * Q * d; g.sCb2 = (long long)Q * d;
  g.nb1 = B; g.nb2 = heads; g.alpha = alpha;
  cudaStream_t s = as_stream(stream);
  if (dtype == HICOM_F32) return launch_gemm_simt<float, float, float>(g, s);
  if (dtype == HICOM_BF16) return launch_gemm_simt<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(g, s);
  if (dtype == HICOM_F16) return launch_gemm_simt<__half, __half, __half>(g, s);
  set_error("global_fold_query: bad dtype %d", dtype);
  return 1;
}

extern "C" int hicom_global_value_proj(const void* pooled, const void* Wv, const void* bv, void* attn, int B,
                                       int Q, int d, int heads, int dtype, void* stream) {
  HICOM_REQUIRE(pooled && Wv && attn, "global_value_proj: null pointer");
  HICOM_REQUIRE(B >= 0 && Q > 0 && heads > 0 && d % heads == 0, "global_value_proj: bad shape");
  if (B == 0) return 0;
  const int hd = d / heads;
  if ((dtype == HICOM_BF16 || dtype == HICOM_F16) && hd % 32 == 0 && d % 8 == 0 && !((uintptr_t)pooled & 15) &&
      !((uintptr_t)Wv & 15)) {
    // one dense tcgen05 GEMM (B*J, d) x Wvᵀ whose epilogue keeps only the diagonal head blocks: the 9x redundant
    // flops (tens of GFLOP) are cheaper than nine launch-bound per-head GEMMs.
    TcLinearParams t{};
    t.A = pooled; t.W = Wv; t.bias = bv; t.C = attn;
    t.lda = d; t.ldw = d; t.ldc = d; t.M = B * heads * Q; t.N = d; t.K = d; t.act = HICOM_ACT_NONE;
    t.out_dtype = dtype; t.a_f16 = t.w_f16 = dtype == HICOM_F16;
    t.rows_per_group = 1 << 30; t.group_stride_rows = 0;
    t.diag_heads = heads; t.diag_rows = Q; t.diag_cols = hd;
    return launch_tc_linear(t, as_stream(stream));
  }
  GemmParams g = plain_gemm();
  g.A = pooled; g.B = Wv; g.bias = bv; g.C = attn;
  g.M = Q; g.N = hd; g.K = d;
  g.sAm = d; g.sAk = 1; g.sAb1 = (long long)heads * Q * d; g.sAb2 = (long long)Q * d;
  g.sBk = 1; g.sBn = d; g.sBb1 = 0; g.sBb2 = (long long)hd * d;
  g.ldc = d; g.sCb1 = (long long)Q * d; g.sCb2 = hd;
  g.sBiasb2 = hd;
  g.nb1 = B; g.nb2 = heads;
  cudaStream_t s = as_stream(stream);
  if (dtype == HICOM_F32) return launch_gemm_simt<float, float, float>(g, s);
  if (dtype == HICOM_BF16) return launch_gemm_simt<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(g, s);
  if (dtype == HICOM_F16) return launch_gemm_simt<__half, __half, __half>(g, s);
  set_error("global_value_proj: bad dtype %d", dtype);
  return 1;
}

// ---- global attention partials --------------------------------------------------------------------
// SIMT pipeline (fp32 mode, and the cross-check for the tensor path):
//   posadd -> S = X'·qfoldᵀ (fp32) -> per-split column softmax stats, P in place -> O = Pᵀ·X' per split.
static size_t elem_size(int dtype) { return dtype == HICOM_F32 ? 4 : 2; }

extern "C" size_t hicom_global_attend_workspace_bytes(int B, int T, int H, int W, int d, int J, int splits,
                                                      int dtype, int impl) {
  (void)splits;
  const size_t N = (size_t)T * H * W;
  if (tc_global_selected(dtype, impl, d, J, T, H, W)) return tc_global_workspace_bytes(B, T, H, W, d, J, splits);
  return align256((size_t)B * N * d * elem_size(dtype)) + align256((size_t)B * N * J * sizeof(float));
}

// Kscore != nullptr: the scores are Kscore·qfoldᵀ (no position terms: the caller's keys already contain them) while the
// pooled operand stays x' = X + pos_embed (hicom_global_attend_partial_keys, the clip-scale variant).
static int global_attend_partial_impl(const void* X, const void* Kscore, const float* pos_t, const float* pos_h,
                                      const float* pos_w, const void* qfold, float* m, float* l, float* o,
                                      int B, int T, int H, int W, int d, int J, int splits, int dtype,
                                      void* workspace, size_t workspace_bytes, int impl, void* stream) {
  HICOM_REQUIRE(X && qfold && m && l && o && workspace, "global_attend_partial: null pointer");
  HICOM_REQUIRE(pos_t && pos_h && pos_w, "global_attend_partial: position tables required");
  HICOM_REQUIRE(B >= 0 && T > 0 && H > 0 && W > 0 && d > 0 && d % 128 == 0 && J > 0 && splits > 0,
                "global_attend_partial: bad shape");
  HICOM_REQUIRE(dtype == HICOM_F32 || dtype == HICOM_BF16 || dtype == HICOM_F16, "global_attend_partial: bad dtype %d", dtype);
  HICOM_REQUIRE(workspace_bytes >= hicom_global_attend_workspace_bytes(B, T, H, W, d, J, splits, dtype, impl),
                "global_attend_partial: workspace too small");
  HICOM_REQUIRE(((uintptr_t)workspace & 255) == 0, "global_attend_partial: workspace must be 256-byte aligned");
  if (B == 0) return 0;
  cudaStream_t s = as_stream(stream);
  const int N = T * H * W;
  if (impl == HICOM_IMPL_TCGEN05)
    HICOM_REQUIRE(tc_global_selected(dtype, impl, d, J, T, H, W), "global_attend_partial: tcgen05 path needs bf16, d%%128==0");
  if (tc_global_selected(dtype, impl, d, J, T, H, W))
    return launch_tc_global(X, Kscore, pos_t, pos_h, pos_w, qfold, m, l, o, B, T, H, W, d, J, splits, workspace, s,
                            dtype == HICOM_F16);

  const int rows_per_split = (N + splits - 1) / splits;
  char* ws = static_cast<char*>(workspace);
  void* Xp = ws;
  float* S = reinterpret_cast<float*>(ws + align256((size_t)B * N * d * elem_size(dtype)));
  if (launch_posadd(X, Xp, pos_t, pos_h, pos_w, B, T, H, W, d, dtype, s)) return 1;
  {  // S[b] (N,J) = X'[b] (N,d) · qfold[b]ᵀ (d,J)
    GemmParams g = plain_gemm();
    g.A = Kscore != nullptr ? Kscore : Xp; g.B = qfold; g.C = S;
    g.M = N; g.N = J; g.K = d;
    g.sAm = d; g.sAk = 1; g.sAb1 = (long long)N * d;
    g.sBk = 1; g.sBn = d; g.sBb1 = (long long)J * d;
    g.ldc = J; g.sCb1 = (long long)N * J;
    g.nb1 = B; g.nb2 = 1;
    int rc = dtype == HICOM_F32 ? launch_gemm_simt<float, float, float>(g, s)
             : dtype == HICOM_F16 ? launch_gemm_simt<__half, __half, float>(g, s)
                                  : launch_gemm_simt<__nv_bfloat16, __nv_bfloat16, float>(g, s);
    if (rc) return rc;
  }
  if (launch_col_softmax(S, m, l, B, N, J, splits, rows_per_split, s)) return 1;
  // O[b,s] (J,d) = P[b, rows of s]ᵀ (J,n) · X'[b, rows of s] (n,d).  A is fp32 (P), B is X' in `dtype`.
  {
    GemmParams g = plain_gemm();
    g.A = S; g.B = Xp; g.C = o;
    g.M = J; g.N = d; g.K = rows_per_split; g.k_total2 = N;
    g.sAm = 1; g.sAk = J; g.sAb1 = (long long)N * J; g.sAb2 = (long long)rows_per_split * J;
    g.sBk = d; g.sBn = 1; g.sBb1 = (long long)N * d; g.sBb2 = (long long)rows_per_split * d;
    g.ldc = d; g.sCb1 = (long long)splits * J * d; g.sCb2 = (long long)J * d;
    g.nb1 = B; g.nb2 = splits;
    int rc = dtype == HICOM_F32 ? launch_gemm_simt<float, float, float>(g, s)
             : dtype == HICOM_F16 ? launch_gemm_simt<float, __half, float>(g, s)
                                  : launch_gemm_simt<float, __nv_bfloat16, float>(g, s);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int hicom_global_attend_partial(const void* X, const float* pos_t, const float* pos_h,
                                           const float* pos_w, const void* qfold, float* m, float* l, float* o,
                                           int B, int T, int H, int W, int d, int J, int splits, int dtype,
                                           void* workspace, size_t workspace_bytes, int impl, void* stream) {
  return global_attend_partial_impl(X, nullptr, pos_t, pos_h, pos_w, qfold, m, l, o, B, T, H, W, d, J, splits, dtype,
                                    workspace, workspace_bytes, impl, stream);
}

extern "C" int hicom_global_attend_partial_keys(const void* X, const void* Kscore, const float* pos_t,
                                                const float* pos_h, const float* pos_w, const void* qfold, float* m,
                                                float* l, float* o, int B, int T, int H, int W, int d, int J,
                                                int splits, int dtype, void* workspace, size_t workspace_bytes,
                                                int impl, void* stream) {
  HICOM_REQUIRE(Kscore != nullptr, "global_attend_partial_keys: null key operand");
  return global_attend_partial_impl(X, Kscore, pos_t, pos_h, pos_w, qfold, m, l, o, B, T, H, W, d, J, splits, dtype,
                                    workspace, workspace_bytes, impl, stream);
}

