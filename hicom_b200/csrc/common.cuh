// Shared helpers for the hicom_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hicom_b200.h"

namespace hicom {

// ---- error plumbing: every entry point returns 0 / non-zero and records a message -------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> 0 / 1 (message recorded)

// Optional per-kernel timing (bench.py's roofline): when enabled, launches wrapped in a KernelTimer are bracketed by
// CUDA events on the launching stream; hicom_kernel_timing_collect() synchronises and sums them per label.
bool kernel_timing_enabled();
void kernel_timing_begin(const char* label, cudaStream_t stream);
void kernel_timing_end(cudaStream_t stream);
struct KernelTimer {
  cudaStream_t s; bool on;
  KernelTimer(const char* label, cudaStream_t stream) : s(stream), on(kernel_timing_enabled()) {
    if (on) kernel_timing_begin(label, s);
  }
  ~KernelTimer() { if (on) kernel_timing_end(s); }
};

// SMs the calling thread's launches may fill (hicom_set_sm_limit; whole device when no limit is set)
int sm_budget();
bool sm_limited();  // a limit below the device's SM count is in force

#define HICOM_REQUIRE(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      ::hicom::set_error(__VA_ARGS__);  \
      return 1;                         \
    }                                   \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- dtype helpers ---------------------------------------------------------------------------
template <typename T>
struct Vec4;  // 4 consecutive elements moved with one load/store

template <>
struct Vec4<float> {
  using raw = float4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 r = *reinterpret_cast<const float4*>(p);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void load_stream(const float* p, float (&v)[4]) {
    float4 r = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <>
struct Vec4<__nv_bfloat16> {
  using raw = uint2;
  static __device__ __forceinline__ void unpack(uint2 r, float (&v)[4]) {
    // bf16 -> fp32 is a 16-bit shift
    v[0] = __uint_as_float(r.x << 16);
    v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16);
    v[3] = __uint_as_float(r.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    unpack(*reinterpret_cast<const uint2*>(p), v);
  }
  static __device__ __forceinline__ void load_stream(const __nv_bfloat16* p, float (&v)[4]) {
    unpack(__ldcs(reinterpret_cast<const uint2*>(p)), v);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

template <>
struct Vec4<__half> {
  using raw = uint2;
  static __device__ __forceinline__ void unpack(uint2 r, float (&v)[4]) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void load(const __half* p, float (&v)[4]) {
    unpack(*reinterpret_cast<const uint2*>(p), v);
  }
  static __device__ __forceinline__ void load_stream(const __half* p, float (&v)[4]) {
    unpack(__ldcs(reinterpret_cast<const uint2*>(p)), v);
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[4]) {
    __half2 a = __floats2half2_rn(v[0], v[1]);
    __half2 b = __floats2half2_rn(v[2], v[3]);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// tanh form ("gelu_pytorch_tanh", the SigLIP MLP activation): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
__device__ __forceinline__ float gelu_tanh(float x) {
  return 0.5f * x * (1.0f + tanhf(0.7978845608028654f * fmaf(0.044715f * x * x, x, x)));
}
__device__ __forceinline__ float apply_act(float v, int act) {
  return act == HICOM_ACT_GELU ? gelu_erf(v) : (act == HICOM_ACT_GELU_TANH ? gelu_tanh(v) : v);
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLnEps = 1e-6f;  // projector.py:318,403,565

// dispatch a lambda-like macro over the three storage types
#define HICOM_DISPATCH_DTYPE(dtype, T, ...)                 \
  do {                                                      \
    if ((dtype) == HICOM_F32) {                             \
      using T = float;                                      \
      __VA_ARGS__;                                          \
    } else if ((dtype) == HICOM_BF16) {                     \
      using T = __nv_bfloat16;                              \
      __VA_ARGS__;                                          \
    } else if ((dtype) == HICOM_F16) {                      \
      using T = __half;                                     \
      __VA_ARGS__;                                          \
    } else {                                                \
      ::hicom::set_error("unknown dtype code %d", (dtype)); \
      return 1;                                             \
    }                                                       \
  } while (0)

// ---- window geometry shared by host validation and kernels (projector.py:501-522) -------------
// Window i of an axis of length n with kernel k covers [start, start+len).
struct AxisWin {
  int n, k, count, keep, len;  // len = members per window (k, or n when n < k)
};

static inline __host__ __device__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Returns false when the reference's torch.stack would fail (unequal window lengths).
static inline __host__ bool make_axis_win(int n, int k, AxisWin* w) {
  w->n = n; w->k = k;
  w->count = ceil_div(n, k);
  if (n % k == 0) { w->keep = w->count; w->len = k; return true; }
  int keep = n % w->count;
  if (keep == 0) keep = w->count;
  w->keep = keep;
  // replay the reference loop and demand equal lengths
  int start = 0, len0 = -1;
  for (int i = 0; i < w->count; ++i) {
    int fresh = k - (i < keep ? 0 : 1);
    int stop = start + fresh;
    if (fresh < k) start -= 1;
    if (start < 0) return false;
    int len = (stop < n ? stop : n) - start;  // python slicing clamps the stop (n < k: one short window)
    if (len <= 0) return false;
    if (len0 < 0) len0 = len; else if (len != len0) return false;
    start = stop;
  }
  w->len = len0;
  return true;
}

// start of window i (valid only when make_axis_win succeeded)
static inline __host__ __device__ int axis_win_start(const AxisWin& w, int i) {
  if (w.n % w.k == 0) return i * w.k;
  // first `keep` windows advance by k; later ones take k-1 fresh elements and step back by one
  if (i < w.keep) return i * w.k;
  return w.keep * w.k + (i - w.keep) * (w.k - 1) - 1;
}

// PyTorch upsample_trilinear3d, align_corners=False: source taps for output index o.
struct Tap { int i0, i1; float w0, w1; };
static inline __host__ __device__ Tap linear_tap(int o, int in_size, int out_size) {
  float scale = (float)in_size / (float)out_size;
  float src = scale * ((float)o + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  int i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  int i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  float w1 = src - (float)i0;
  Tap t; t.i0 = i0; t.i1 = i1; t.w1 = w1; t.w0 = 1.f - w1;
  return t;
}

}  // namespace hicom
