// Fused local compressor: grid-pool -> instruction injection -> window softmax -> A·V.
//
// Replaces, in one pass over HBM, the reference's
//   K/V identity mixes       projector.py:533-534  (6 elementwise passes, eliminated)
//   trilinear grid pooling   projector.py:539-540
//   coarse/direct injection  projector.py:352-372  (the MLP itself runs once per video elsewhere)
//   window gather x3         projector.py:473-522,544-546  (never materialised)
//   bmm / softmax / bmm      projector.py:549-553
//
// HBM-bound: per window it must read kt*ks*ks key rows and as many value rows of d elements.
// One warp owns one window; lane L owns the 4-element chunks {L, L+32, ...} of the channel axis
// (d = 128*CPL), so every row read is a fully coalesced 256 B (bf16) / 512 B (fp32) request per
// chunk, and all per-channel state (query, accumulator, FiLM vectors) lives in registers.
// Softmax is the online form so arbitrary window sizes (e.g. local412: 576 members) work.
#include <stdlib.h>

#include "common.cuh"

namespace hicom {

struct LocalParams {
  const void* K;
  const void* V;
  const void* P;
  const void* q_aux;
  const float* film;
  const void* ln_w;
  const void* ln_b;
  void* out;
  int B, T, H, W, d;
  AxisWin wt, wh, ww;
  int qmode;
  float scale_log2;  // logit scale * log2(e)
  int k_l2norm;
  int same_kv;    // K and V alias: load each row once
  int pool_only;  // write the pooled query and stop (hicom_grid_pool)
};

template <typename T, int CPL>
__global__ void __launch_bounds__(256) local_attend_kernel(const LocalParams p) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long nwin_video = (long long)p.wt.count * p.wh.count * p.ww.count;
  const long long total = nwin_video * p.B;
  const int d = p.d;
  const T* __restrict__ Kp = static_cast<const T*>(p.K);
  const T* __restrict__ Vp = static_cast<const T*>(p.V);
  const T* __restrict__ Pp = static_cast<const T*>(p.P);
  T* __restrict__ outp = static_cast<T*>(p.out);

  for (long long wi = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); wi < total;
       wi += (long long)gridDim.x * warps_per_block) {
    const int b = (int)(wi / nwin_video);
    const int r = (int)(wi % nwin_video);
    const int t1 = r / (p.wh.count * p.ww.count);
    const int h1 = (r / p.ww.count) % p.wh.count;
    const int w1 = r % p.ww.count;
    const size_t video_off = (size_t)b * p.T * p.H * p.W * d;

    float q[CPL][4];

    // ---- query ------------------------------------------------------------------------------
    if (p.qmode == HICOM_Q_POOLED || p.qmode == HICOM_Q_FILM_LN) {
      // trilinear taps, nested exactly like upsample_trilinear3d: t(h(w))
      const Tap tt = linear_tap(t1, p.T, p.wt.count);
      const Tap th = linear_tap(h1, p.H, p.wh.count);
      const Tap tw = linear_tap(w1, p.W, p.ww.count);
#pragma unroll
      for (int c = 0; c < CPL; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) q[c][e] = 0.f;
      const int ti[2] = {tt.i0, tt.i1}; const float tw_[2] = {tt.w0, tt.w1};
      const int hi[2] = {th.i0, th.i1}; const float hw_[2] = {th.w0, th.w1};
      const int wi_[2] = {tw.i0, tw.i1}; const float ww_[2] = {tw.w0, tw.w1};
      for (int a = 0; a < 2; ++a) {
        if (tw_[a] == 0.f) continue;
        for (int bb = 0; bb < 2; ++bb) {
          if (hw_[bb] == 0.f) continue;
          for (int cc = 0; cc < 2; ++cc) {
            if (ww_[cc] == 0.f) continue;
            const float wgt = tw_[a] * (hw_[bb] * ww_[cc]);
            const T* row = Pp + video_off + ((size_t)(ti[a] * p.H + hi[bb]) * p.W + wi_[cc]) * d;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
              float x[4];
              Vec4<T>::load(row + (lane + 32 * c) * 4, x);
#pragma unroll
              for (int e = 0; e < 4; ++e) q[c][e] = fmaf(wgt, x[e], q[c][e]);
            }
          }
        }
      }
      if (p.pool_only) {
        T* orow = outp + (size_t)wi * d;
#pragma unroll
        for (int c = 0; c < CPL; ++c) Vec4<T>::store(orow + (lane + 32 * c) * 4, q[c]);
        continue;
      }
      if (p.qmode == HICOM_Q_FILM_LN) {
        const float* sc = p.film + (size_t)b * 2 * d;
        const float* sh = sc + d;
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          float s4[4], h4[4];
          Vec4<float>::load(sc + (lane + 32 * c) * 4, s4);
          Vec4<float>::load(sh + (lane + 32 * c) * 4, h4);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            q[c][e] = fmaf(q[c][e], 1.f + s4[e], h4[e]);
            sum += q[c][e];
          }
        }
        const float mean = warp_sum(sum) / (float)d;
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float dv = q[c][e] - mean;
            var = fmaf(dv, dv, var);
          }
        const float rstd = rsqrtf(warp_sum(var) / (float)d + kLnEps);
        const T* gw = static_cast<const T*>(p.ln_w);
        const T* gb = static_cast<const T*>(p.ln_b);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          float g4[4], b4[4];
          Vec4<T>::load(gw + (lane + 32 * c) * 4, g4);
          Vec4<T>::load(gb + (lane + 32 * c) * 4, b4);
#pragma unroll
          for (int e = 0; e < 4; ++e) q[c][e] = fmaf((q[c][e] - mean) * rstd, g4[e], b4[e]);
        }
      }
    } else {
      const T* qrow = static_cast<const T*>(p.q_aux) +
                      (p.qmode == HICOM_Q_VECTOR ? (size_t)b * d : (size_t)wi * d);
#pragma unroll
      for (int c = 0; c < CPL; ++c) Vec4<T>::load(qrow + (lane + 32 * c) * 4, q[c]);
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) q[c][e] *= p.scale_log2;

    // ---- window members: online softmax over (t2,h2,w2) ----------------------------------------
    const int t0 = axis_win_start(p.wt, t1), h0 = axis_win_start(p.wh, h1), w0 = axis_win_start(p.ww, w1);
    const int Lt = p.wt.len, Lh = p.wh.len, Lw = p.ww.len;
    const int members = Lt * Lh * Lw;
    float acc[CPL][4];
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[c][e] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    for (int mi = 0; mi < members; mi += 2) {
      const bool two = (mi + 1) < members;
      size_t off[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int mj = two ? mi + j : mi;
        const int t2 = mj / (Lh * Lw), h2 = (mj / Lw) % Lh, w2 = mj % Lw;
        off[j] = video_off + ((size_t)((t0 + t2) * p.H + h0 + h2) * p.W + w0 + w2) * d;
      }
      float kv[2][CPL][4];
      float dot[2] = {0.f, 0.f}, kk[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < CPL; ++c) Vec4<T>::load_stream(Kp + off[j] + (lane + 32 * c) * 4, kv[j][c]);
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < CPL; ++c)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            dot[j] = fmaf(q[c][e], kv[j][c][e], dot[j]);
            if (p.k_l2norm) kk[j] = fmaf(kv[j][c][e], kv[j][c][e], kk[j]);
          }
      if (!p.same_kv) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int c = 0; c < CPL; ++c) Vec4<T>::load_stream(Vp + off[j] + (lane + 32 * c) * 4, kv[j][c]);
      }
      float s0 = warp_sum(dot[0]), s1 = warp_sum(dot[1]);
      if (p.k_l2norm) {
        s0 *= rsqrtf(warp_sum(kk[0]));
        s1 *= rsqrtf(warp_sum(kk[1]));
      }
      if (!two) s1 = -INFINITY;
      const float m_new = fmaxf(m_run, fmaxf(s0, s1));
      const float corr = exp2f(m_run - m_new);
      const float p0 = exp2f(s0 - m_new), p1 = exp2f(s1 - m_new);
      l_run = fmaf(l_run, corr, p0 + p1);
      m_run = m_new;
#pragma unroll
      for (int c = 0; c < CPL; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          acc[c][e] = fmaf(acc[c][e], corr, fmaf(p0, kv[0][c][e], p1 * kv[1][c][e]));
    }
    const float inv = 1.f / l_run;
    T* orow = outp + (size_t)wi * d;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[c][e] *= inv;
      Vec4<T>::store(orow + (lane + 32 * c) * 4, acc[c]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// v2: TMA-staged variant.  Each warp owns a ring of RM (key row, value row) slots in shared memory; one lane issues
// `cp.async.bulk` (1-D TMA, SASS UBLKCP) copies of whole rows that complete on a per-slot mbarrier, so RM members
// (RM * 2 rows * d elements) are in flight per warp without holding them in registers — twice the bytes in flight and
// twice the resident warps of the register-staged kernel.  Same math, same accumulation order.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t la_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void la_mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(la_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void la_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(la_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void la_bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(la_smem_u32(dst)), "l"(src), "r"(bytes), "r"(la_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void la_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(la_smem_u32(bar)), "r"(parity) : "memory");
  }
}

// WPB = 8: two CTAs per SM (bf16).  WPB = 16: ONE CTA per SM holding the same 16 rings (all of the SM's shared memory),
// launched as clusters of two so that a grid limited to n SMs (hicom_set_sm_limit) occupies n/2 whole SM pairs and the
// tensor-core kernels running beside it keep the remaining pairs for their cta_group::2 tiles.
template <typename T, int CPL, int RM, int WPB>
__global__ void __launch_bounds__(WPB * 32, (sizeof(T) == 2 && WPB == 8) ? 2 : 1) local_attend_v2_kernel(const LocalParams p) {
  extern __shared__ __align__(128) uint8_t la_smem[];
  constexpr int ROWB = CPL * 128 * (int)sizeof(T);
  constexpr int SLOTB = 2 * ROWB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  uint8_t* ring = la_smem + (size_t)warp * RM * SLOTB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(la_smem + (size_t)warps_per_block * RM * SLOTB) + warp * RM;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < RM; ++i) la_mbar_init(&bars[i]);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const long long nwin_video = (long long)p.wt.count * p.wh.count * p.ww.count;
  const long long total = nwin_video * p.B;
  const int d = p.d;
  const T* __restrict__ Kp = static_cast<const T*>(p.K);
  const T* __restrict__ Vp = static_cast<const T*>(p.V);
  const T* __restrict__ Pp = static_cast<const T*>(p.P);
  T* __restrict__ outp = static_cast<T*>(p.out);
  const uint32_t tx_bytes = p.same_kv ? ROWB : SLOTB;
  uint32_t produced = 0, consumed = 0;  // members issued / consumed by this warp since kernel start (slot + phase)

  for (long long wi = (long long)blockIdx.x * warps_per_block + warp; wi < total;
       wi += (long long)gridDim.x * warps_per_block) {
    const int b = (int)(wi / nwin_video);
    const int r = (int)(wi % nwin_video);
    const int t1 = r / (p.wh.count * p.ww.count);
    const int h1 = (r / p.ww.count) % p.wh.count;
    const int w1 = r % p.ww.count;
    const size_t video_off = (size_t)b * p.T * p.H * p.W * d;
    const int t0 = axis_win_start(p.wt, t1), h0 = axis_win_start(p.wh, h1), w0 = axis_win_start(p.ww, w1);
    const int Lt = p.wt.len, Lh = p.wh.len, Lw = p.ww.len;
    const int members = Lt * Lh * Lw;

    auto issue = [&](int mj) {  // lane 0 only
      const int t2 = mj / (Lh * Lw), h2 = (mj / Lw) % Lh, w2 = mj % Lw;
      const size_t off = video_off + ((size_t)((t0 + t2) * p.H + h0 + h2) * p.W + w0 + w2) * d;
      const uint32_t slot = produced % RM;
      uint8_t* dst = ring + (size_t)slot * SLOTB;
      la_expect_tx(&bars[slot], tx_bytes);
      la_bulk_copy(dst, Kp + off, ROWB, &bars[slot]);
      if (!p.same_kv) la_bulk_copy(dst + ROWB, Vp + off, ROWB, &bars[slot]);
      ++produced;
    };
    // start the window's first RM members before computing the query (their latency hides the query's)
    if (lane == 0) {
      for (int k = 0; k < RM && k < members; ++k) issue(k);
    }

    float q[CPL][4];
    // ---- query (identical to v1) -------------------------------------------------------------
    if (p.qmode == HICOM_Q_POOLED || p.qmode == HICOM_Q_FILM_LN) {
      const Tap tt = linear_tap(t1, p.T, p.wt.count);
      const Tap th = linear_tap(h1, p.H, p.wh.count);
      const Tap tw = linear_tap(w1, p.W, p.ww.count);
#pragma unroll
      for (int c = 0; c < CPL; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) q[c][e] = 0.f;
      const int ti[2] = {tt.i0, tt.i1}; const float tw_[2] = {tt.w0, tt.w1};
      const int hi[2] = {th.i0, th.i1}; const float hw_[2] = {th.w0, th.w1};
      const int wi_[2] = {tw.i0, tw.i1}; const float ww_[2] = {tw.w0, tw.w1};
      for (int a = 0; a < 2; ++a) {
        if (tw_[a] == 0.f) continue;
        for (int bb = 0; bb < 2; ++bb) {
          if (hw_[bb] == 0.f) continue;
          for (int cc = 0; cc < 2; ++cc) {
            if (ww_[cc] == 0.f) continue;
            const float wgt = tw_[a] * (hw_[bb] * ww_[cc]);
            const T* row = Pp + video_off + ((size_t)(ti[a] * p.H + hi[bb]) * p.W + wi_[cc]) * d;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
              float x[4];
              Vec4<T>::load(row + (lane + 32 * c) * 4, x);
#pragma unroll
              for (int e = 0; e < 4; ++e) q[c][e] = fmaf(wgt, x[e], q[c][e]);
            }
          }
        }
      }
      if (p.qmode == HICOM_Q_FILM_LN) {
        const float* sc = p.film + (size_t)b * 2 * d;
        const float* sh = sc + d;
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          float s4[4], h4[4];
          Vec4<float>::load(sc + (lane + 32 * c) * 4, s4);
          Vec4<float>::load(sh + (lane + 32 * c) * 4, h4);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            q[c][e] = fmaf(q[c][e], 1.f + s4[e], h4[e]);
            sum += q[c][e];
          }
        }
        const float mean = warp_sum(sum) / (float)d;
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float dv = q[c][e] - mean;
            var = fmaf(dv, dv, var);
          }
        const float rstd = rsqrtf(warp_sum(var) / (float)d + kLnEps);
        const T* gw = static_cast<const T*>(p.ln_w);
        const T* gb = static_cast<const T*>(p.ln_b);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          float g4[4], b4[4];
          Vec4<T>::load(gw + (lane + 32 * c) * 4, g4);
          Vec4<T>::load(gb + (lane + 32 * c) * 4, b4);
#pragma unroll
          for (int e = 0; e < 4; ++e) q[c][e] = fmaf((q[c][e] - mean) * rstd, g4[e], b4[e]);
        }
      }
    } else {
      const T* qrow = static_cast<const T*>(p.q_aux) +
                      (p.qmode == HICOM_Q_VECTOR ? (size_t)b * d : (size_t)wi * d);
#pragma unroll
      for (int c = 0; c < CPL; ++c) Vec4<T>::load(qrow + (lane + 32 * c) * 4, q[c]);
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) q[c][e] *= p.scale_log2;

    // ---- members from the smem ring: online softmax ----------------------------------------------
    float acc[CPL][4];
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[c][e] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int mi = 0; mi < members; ++mi) {
      const uint32_t slot = consumed % RM;
      la_wait(&bars[slot], (consumed / RM) & 1);
      const T* krow = reinterpret_cast<const T*>(ring + (size_t)slot * SLOTB);
      const T* vrow = p.same_kv ? krow : krow + CPL * 128;
      float dot = 0.f, kk = 0.f;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        float x[4];
        Vec4<T>::load(krow + (lane + 32 * c) * 4, x);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          dot = fmaf(q[c][e], x[e], dot);
          if (p.k_l2norm) kk = fmaf(x[e], x[e], kk);
        }
      }
      float s = warp_sum(dot);
      if (p.k_l2norm) s *= rsqrtf(warp_sum(kk));
      const float m_new = fmaxf(m_run, s);
      const float corr = exp2f(m_run - m_new);
      const float pw = exp2f(s - m_new);
      l_run = fmaf(l_run, corr, pw);
      m_run = m_new;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        float x[4];
        Vec4<T>::load(vrow + (lane + 32 * c) * 4, x);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[c][e] = fmaf(acc[c][e], corr, pw * x[e]);
      }
      ++consumed;
      __syncwarp();  // every lane is done with this slot before it is refilled
      if (lane == 0 && mi + RM < members) issue(mi + RM);
    }
    const float inv = 1.f / l_run;
    T* orow = outp + (size_t)wi * d;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[c][e] *= inv;
      Vec4<T>::store(orow + (lane + 32 * c) * 4, acc[c]);
    }
  }
}

template <typename T, int CPL, int WPB>
static int launch_local_v2(const LocalParams& p, cudaStream_t stream, bool pairs) {
  constexpr int RM = sizeof(T) == 2 ? 3 : 2;  // members in flight per warp
  constexpr size_t smem = (size_t)WPB * RM * 2 * CPL * 128 * sizeof(T) + WPB * RM * 8;
  static bool configured = false;
  auto kern = local_attend_v2_kernel<T, CPL, RM, WPB>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    HICOM_REQUIRE(e == cudaSuccess, "local_attend: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const long long total = (long long)p.wt.count * p.wh.count * p.ww.count * p.B;
  const int num_sms = sm_budget();
  long long blocks = (total + WPB - 1) / WPB;
  const long long resident = (long long)num_sms * ((sizeof(T) == 2 && WPB == 8) ? 2 : 1);
  if (blocks > resident) blocks = resident;  // persistent warps: the ring stays primed across windows
  KernelTimer timer("local_attend_v2", stream);
  if (pairs) {
    blocks = (blocks + 1) & ~1ll;  // whole clusters; a surplus CTA finds no window and exits
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(WPB * 32); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    HICOM_REQUIRE(e == cudaSuccess, "cudaLaunchKernelEx(local_attend_v2): %s", cudaGetErrorString(e));
  } else {
    kern<<<(unsigned)blocks, WPB * 32, smem, stream>>>(p);
  }
  return check_launch("local_attend_v2_kernel");
}

template <typename T>
static int launch_local(const LocalParams& p, cudaStream_t stream) {
  const long long total = (long long)p.wt.count * p.wh.count * p.ww.count * p.B;
  if (total == 0) return 0;
  const int warps_per_block = 8;
  long long blocks = (total + warps_per_block - 1) / warps_per_block;
  if (blocks > (1 << 20)) blocks = 1 << 20;
  const int cpl = p.d / 128;
  const bool aligned = (((uintptr_t)p.K | (uintptr_t)p.V) & 15) == 0;
  if (!p.pool_only && aligned) {
    // under an SM limit (the tensor-bound chain runs beside this kernel): one 16-warp CTA per SM, whole SM pairs
    if (sizeof(T) == 2 && sm_limited() && cpl == 9) return launch_local_v2<T, 9, 16>(p, stream, true);
    switch (cpl) {
      case 9: return launch_local_v2<T, 9, 8>(p, stream, false);
      case 6: return launch_local_v2<T, 6, 8>(p, stream, false);
      case 8: return launch_local_v2<T, 8, 8>(p, stream, false);
      default: break;
    }
  }
  switch (cpl) {
    case 9: local_attend_kernel<T, 9><<<(unsigned)blocks, warps_per_block * 32, 0, stream>>>(p); break;
    case 6: local_attend_kernel<T, 6><<<(unsigned)blocks, warps_per_block * 32, 0, stream>>>(p); break;
    case 8: local_attend_kernel<T, 8><<<(unsigned)blocks, warps_per_block * 32, 0, stream>>>(p); break;
    case 1: local_attend_kernel<T, 1><<<(unsigned)blocks, warps_per_block * 32, 0, stream>>>(p); break;
    default:
      set_error("local_attend: d=%d unsupported (need d in {128,768,1024,1152})", p.d);
      return 1;
  }
  return check_launch("local_attend_kernel");
}

static int fill_geometry(LocalParams& p, int T, int H, int W, int kt, int ks) {
  HICOM_REQUIRE(kt >= 1 && ks >= 1, "local_attend: kernel sizes must be >= 1 (kt=%d ks=%d)", kt, ks);
  HICOM_REQUIRE(make_axis_win(T, kt, &p.wt),
                "local_attend: T=%d with temporal kernel %d gives unequal windows (the reference's "
                "torch.stack raises here, projector.py:520)", T, kt);
  HICOM_REQUIRE(make_axis_win(H, ks, &p.wh), "local_attend: H=%d with kernel %d gives unequal windows", H, ks);
  HICOM_REQUIRE(make_axis_win(W, ks, &p.ww), "local_attend: W=%d with kernel %d gives unequal windows", W, ks);
  return 0;
}

}  // namespace hicom

using namespace hicom;

extern "C" int hicom_grid_pool(const void* X, void* Q, int B, int T, int H, int W, int d, int kt, int ks,
                               int dtype, void* stream) {
  HICOM_REQUIRE(X && Q, "grid_pool: null pointer");
  HICOM_REQUIRE(B >= 0 && T > 0 && H > 0 && W > 0 && d > 0 && d % 128 == 0, "grid_pool: bad shape");
  LocalParams p{};
  p.K = p.V = p.P = X; p.out = Q;
  p.B = B; p.T = T; p.H = H; p.W = W; p.d = d;
  p.qmode = HICOM_Q_POOLED; p.pool_only = 1; p.scale_log2 = 1.f;
  // pooling only needs the window COUNTS (ceil(n/k)); geometry validity is irrelevant here
  p.wt.n = T; p.wt.k = kt; p.wt.count = ceil_div(T, kt); p.wt.keep = p.wt.count; p.wt.len = 1;
  p.wh.n = H; p.wh.k = ks; p.wh.count = ceil_div(H, ks); p.wh.keep = p.wh.count; p.wh.len = 1;
  p.ww.n = W; p.ww.k = ks; p.ww.count = ceil_div(W, ks); p.ww.keep = p.ww.count; p.ww.len = 1;
  HICOM_DISPATCH_DTYPE(dtype, T_, return launch_local<T_>(p, as_stream(stream)));
}

extern "C" int hicom_local_attend(const void* Ksrc, const void* Vsrc, const void* Psrc, const void* q_aux,
                                  const float* film, const void* ln_w, const void* ln_b, void* out, int B,
                                  int T, int H, int W, int d, int kt, int ks, int qmode, float logit_scale,
                                  int k_l2norm, int dtype, void* stream) {
  HICOM_REQUIRE(Ksrc && Vsrc && out, "local_attend: null pointer");
  HICOM_REQUIRE(B >= 0 && T > 0 && H > 0 && W > 0 && d > 0 && d % 128 == 0, "local_attend: bad shape");
  HICOM_REQUIRE(qmode >= HICOM_Q_POOLED && qmode <= HICOM_Q_EXPLICIT, "local_attend: bad qmode %d", qmode);
  if (qmode == HICOM_Q_POOLED || qmode == HICOM_Q_FILM_LN) HICOM_REQUIRE(Psrc, "local_attend: Psrc required");
  if (qmode == HICOM_Q_FILM_LN) HICOM_REQUIRE(film && ln_w && ln_b, "local_attend: film/ln required");
  if (qmode == HICOM_Q_VECTOR || qmode == HICOM_Q_EXPLICIT) HICOM_REQUIRE(q_aux, "local_attend: q_aux required");
  LocalParams p{};
  p.K = Ksrc; p.V = Vsrc; p.P = Psrc; p.q_aux = q_aux; p.film = film; p.ln_w = ln_w; p.ln_b = ln_b;
  p.out = out; p.B = B; p.T = T; p.H = H; p.W = W; p.d = d; p.qmode = qmode;
  p.scale_log2 = logit_scale * kLog2e; p.k_l2norm = k_l2norm; p.same_kv = (Ksrc == Vsrc);
  if (fill_geometry(p, T, H, W, kt, ks)) return 1;
  HICOM_DISPATCH_DTYPE(dtype, T_, return launch_local<T_>(p, as_stream(stream)));
}
