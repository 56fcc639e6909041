// ByteNet residual blocks (reference modules/block.py:86-173) and the frame-at-a-time LinearConv1d
// (modules/linear_conv_ops.py:39-68) on the generic NCL path.
//
// A block is  seq + [LN, ReLU, 1x1, LN, ReLU, (MU(k,d), MU(1) | causal conv, LN, ReLU), 1x1](seq).  Here no normalised
// tensor is ever written in inference: `ln_stats_kernel` leaves (mean, 1/(std+eps)) per frame, the contraction that
// consumes the LayerNorm applies gamma * (x - mean) * r + beta and the ReLU as it loads its operand
// (WNB200_PRE_LNRELU, taps_simt.cu), the MultiplicativeUnit's gate is the contraction's epilogue (WNB200_EPI_MU) and
// the residual add rides on the last contraction's store.  This file holds the statistics kernel, the stand-alone
// LayerNorm+ReLU (training keeps h for the MU gate's backward), the backward of LayerNorm+ReLU (dx and the gamma /
// beta gradients in one pass over a [C x 128 frames] tile) and the one-frame GEMV of LinearConv1d.linear with its
// ring-buffer history for incremental decoding.
#include <string.h>

#include "common.cuh"

namespace wnb {

// ------------------------------------------------------------------------------------------ LayerNorm statistics
// CTA = [C channels x 128 frames]: 32 frame lanes x 4 frames, 8 channel groups.  Two passes over the tile (the second
// finds it in L1/L2): mean, then the UNBIASED variance about that mean (layernorm.py:26-27 -- a one-pass sum of squares
// would lose the fp32 parity when |mean| >> std).
constexpr int LN_FR = 128, LN_CG = 8;

template <typename T>
__device__ __forceinline__ void load4(const T* row, int t, int Tn, bool vec, float v[4]) {
  if (vec && t + 3 < Tn) {
    if constexpr (sizeof(T) == 4) {
      const float4 q = *reinterpret_cast<const float4*>(row + t);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
      const uint2 q = *reinterpret_cast<const uint2*>(row + t);
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&q.x);
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&q.y);
      v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (t + i < Tn) ? to_f32<T>(row[t + i]) : 0.f;
  }
}

template <typename T>
__device__ __forceinline__ void store4(T* row, int t, int Tn, bool vec, const float v[4]) {
  if (vec && t + 3 < Tn) {
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(row + t) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
      uint2 q;
      q.x = *reinterpret_cast<const uint32_t*>(&a);
      q.y = *reinterpret_cast<const uint32_t*>(&b);
      *reinterpret_cast<uint2*>(row + t) = q;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (t + i < Tn) row[t + i] = from_f32<T>(v[i]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) ln_stats_kernel(int C, int Tn, const T* x, float eps, float* stats) {
  __shared__ float red[LN_CG][LN_FR];
  __shared__ float mean_s[LN_FR];
  const int b = blockIdx.y, t0 = blockIdx.x * LN_FR;
  const int lane = threadIdx.x & 31, cg = threadIdx.x >> 5;
  const int t = t0 + lane * 4;
  const bool vec = (Tn % 4 == 0);
  const T* xb = x + (long long)b * C * Tn;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = cg; c < C; c += LN_CG) {
    float v[4];
    load4<T>(xb + (long long)c * Tn, t, Tn, vec, v);
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] += v[i];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) red[cg][lane * 4 + i] = s[i];
  __syncthreads();
  if (threadIdx.x < LN_FR) {
    float a = 0.f;
#pragma unroll
    for (int g = 0; g < LN_CG; ++g) a += red[g][threadIdx.x];
    mean_s[threadIdx.x] = a / (float)C;
  }
  __syncthreads();
  float m[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = mean_s[lane * 4 + i]; s[i] = 0.f; }
  for (int c = cg; c < C; c += LN_CG) {
    float v[4];
    load4<T>(xb + (long long)c * Tn, t, Tn, vec, v);
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] += (v[i] - m[i]) * (v[i] - m[i]);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) red[cg][lane * 4 + i] = s[i];
  __syncthreads();
  if (threadIdx.x < LN_FR && t0 + threadIdx.x < Tn) {
    float var = 0.f;
#pragma unroll
    for (int g = 0; g < LN_CG; ++g) var += red[g][threadIdx.x];
    const float r = 1.f / (sqrtf(var / (float)(C - 1)) + eps);      // unbiased std, eps on the std (layernorm.py:27)
    *reinterpret_cast<float2*>(stats + ((long long)b * Tn + t0 + threadIdx.x) * 2) = make_float2(mean_s[threadIdx.x], r);
  }
}

// ------------------------------------------------------------------------------------------ LayerNorm + ReLU, stand-alone
template <typename T>
__global__ void __launch_bounds__(256) ln_relu_fwd_kernel(int C, int Tn, const T* x, const float* stats,
                                                          const float* gamma, const float* beta, T* y) {
  const int b = blockIdx.y, t0 = blockIdx.x * LN_FR;
  const int lane = threadIdx.x & 31, cg = threadIdx.x >> 5;
  const int t = t0 + lane * 4;
  if (t >= Tn) return;
  const bool vec = (Tn % 4 == 0);
  float m[4], r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 st = (t + i < Tn) ? *reinterpret_cast<const float2*>(stats + ((long long)b * Tn + t + i) * 2)
                                   : make_float2(0.f, 0.f);
    m[i] = st.x; r[i] = st.y;
  }
  const long long base = (long long)b * C * Tn;
  for (int c = cg; c < C; c += LN_CG) {
    float v[4];
    load4<T>(x + base + (long long)c * Tn, t, Tn, vec, v);
    const float g = gamma[c], bb = beta[c];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = fmaxf(g * (v[i] - m[i]) * r[i] + bb, 0.f);
    store4<T>(y + base + (long long)c * Tn, t, Tn, vec, v);
  }
}

// ------------------------------------------------------------------------------------------ backward of ReLU(LayerNorm(x))
// y = relu(gamma * xc * r + beta), xc = x - mean, r = 1 / (s + eps), s = sqrt(sum xc^2 / (C - 1)).  With
// g_c = dy_c * [y_c > 0] * gamma_c:   dx_c = g_c * r - r * sum(g) / C - r^2 * sum(g * xc) / ((C - 1) * s) * xc_c
// and  dgamma_c += sum_t dy_c [y_c > 0] xc_c r,  dbeta_c += sum_t dy_c [y_c > 0]  (one atomicAdd per channel and CTA).
template <typename T>
__global__ void __launch_bounds__(256) ln_relu_bwd_kernel(int C, int Tn, const T* x, const float* stats,
                                                          const float* gamma, const float* beta, float eps, const T* dy,
                                                          T* dx, float* dgamma, float* dbeta, int relu) {
  __shared__ float red1[LN_CG][LN_FR];
  __shared__ float red2[LN_CG][LN_FR];
  __shared__ float kk[LN_FR], mu[LN_FR];
  const int b = blockIdx.y, t0 = blockIdx.x * LN_FR;
  const int lane = threadIdx.x & 31, cg = threadIdx.x >> 5;
  const int t = t0 + lane * 4;
  const bool vec = (Tn % 4 == 0);
  const long long base = (long long)b * C * Tn;
  float m[4], r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 st = (t + i < Tn) ? *reinterpret_cast<const float2*>(stats + ((long long)b * Tn + t + i) * 2)
                                   : make_float2(0.f, 0.f);
    m[i] = st.x; r[i] = st.y;
  }
  float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = cg; c < C; c += LN_CG) {
    float v[4], d[4];
    load4<T>(x + base + (long long)c * Tn, t, Tn, vec, v);
    load4<T>(dy + base + (long long)c * Tn, t, Tn, vec, d);
    const float g = gamma[c], bb = beta[c];
    float dg = 0.f, db = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float xc = v[i] - m[i], xh = xc * r[i];
      const float dd = (t + i < Tn && (!relu || g * xh + bb > 0.f)) ? d[i] : 0.f;
      a1[i] += dd * g;
      a2[i] += dd * g * xc;
      dg += dd * xh;
      db += dd;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dg += __shfl_xor_sync(0xffffffffu, dg, o);
      db += __shfl_xor_sync(0xffffffffu, db, o);
    }
    if (lane == 0) {
      if (dgamma) atomicAdd(&dgamma[c], dg);
      if (dbeta) atomicAdd(&dbeta[c], db);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { red1[cg][lane * 4 + i] = a1[i]; red2[cg][lane * 4 + i] = a2[i]; }
  __syncthreads();
  if (threadIdx.x < LN_FR) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int g = 0; g < LN_CG; ++g) { s1 += red1[g][threadIdx.x]; s2 += red2[g][threadIdx.x]; }
    float rr = 0.f;
    if (t0 + threadIdx.x < Tn) rr = stats[((long long)b * Tn + t0 + threadIdx.x) * 2 + 1];
    const float sd = rr > 0.f ? 1.f / rr - eps : 0.f;
    kk[threadIdx.x] = (sd > 0.f) ? (-rr * rr * s2 / ((float)(C - 1) * sd)) : 0.f;
    mu[threadIdx.x] = rr * s1 / (float)C;
  }
  __syncthreads();
  if (t >= Tn || dx == nullptr) return;
  float k4[4], mu4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { k4[i] = kk[lane * 4 + i]; mu4[i] = mu[lane * 4 + i]; }
  for (int c = cg; c < C; c += LN_CG) {
    float v[4], d[4], o[4];
    load4<T>(x + base + (long long)c * Tn, t, Tn, vec, v);
    load4<T>(dy + base + (long long)c * Tn, t, Tn, vec, d);
    const float g = gamma[c], bb = beta[c];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float xc = v[i] - m[i];
      const float dd = (!relu || g * xc * r[i] + bb > 0.f) ? d[i] : 0.f;
      o[i] = dd * g * r[i] + k4[i] * xc - mu4[i];
    }
    store4<T>(dx + base + (long long)c * Tn, t, Tn, vec, o);
  }
}

// ------------------------------------------------------------------------------------------ LinearConv1d, one frame
// y[n, m] = bias[m] + sum_c sum_j W[m, c, j] * in_j[n, c]     (linear_conv_ops.py:57-63: F.linear on the gathered taps)
// in_j is tap j's frame: element (n, c) at tap[j] + n * sn + c * sc.  The gathered operand (N x Cin*k, a few KB) is
// staged in shared memory once per CTA; a warp owns output rows and streams W's rows with coalesced loads -- every
// weight is read once per 8 batch items: the launch is weight-bandwidth (in practice launch-latency) bound.
constexpr int LS_MAXK = 32, LS_NB = 8;

struct LinParams {
  const void* tap[LS_MAXK];
  long long sn, sc;
  int N, Cin, Cout, k, rows_per_cta;
  const void* w;
  const float* bias;
  void* y;
  // ring update (incremental decoding): the new frame [N, Cin] is copied into this slot by CTA 0; null: none
  void* ring_slot;
  const void* new_frame;
};

template <typename T>
__global__ void __launch_bounds__(256) linear_frame_kernel(const LinParams p) {
  extern __shared__ __align__(16) float xin[];   // [nb][K], K = Cin * k, W's column order (c major, tap minor)
  const int K = p.Cin * p.k;
  const int n0 = blockIdx.y * LS_NB;
  const int nb = min(LS_NB, p.N - n0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // gather: one (batch item, tap) pair per warp pass, lanes along the channels (contiguous in the history ring)
  for (int pair = warp; pair < nb * p.k; pair += 8) {
    const int n = pair / p.k, j = pair - n * p.k;
    const T* src = reinterpret_cast<const T*>(p.tap[j]) + (long long)(n0 + n) * p.sn;
    float* dst = xin + n * K + j;
    for (int c = lane; c < p.Cin; c += 32) dst[c * p.k] = to_f32<T>(src[(long long)c * p.sc]);
  }
  if (p.ring_slot && blockIdx.x == 0 && blockIdx.y == 0) {
    // the slot written here is not among the taps of this step (ring length = receptive field)
    T* dst = reinterpret_cast<T*>(p.ring_slot);
    const T* src = reinterpret_cast<const T*>(p.new_frame);
    for (int i = threadIdx.x; i < p.N * p.Cin; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const T* w = reinterpret_cast<const T*>(p.w);
  T* y = reinterpret_cast<T*>(p.y);
  const int m_begin = blockIdx.x * p.rows_per_cta;
  const int m_end = min(p.Cout, m_begin + p.rows_per_cta);
  const bool vec = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0;
  for (int m = m_begin + warp; m < m_end; m += 8) {
    float acc[LS_NB];
#pragma unroll
    for (int n = 0; n < LS_NB; ++n) acc[n] = 0.f;
    const T* wr = w + (long long)m * K;
    if (vec) {
      // 4 weights per lane and step (16-byte loads when fp32), four steps in flight
#pragma unroll 4
      for (int e = lane * 4; e < K; e += 128) {
        float wv[4];
        if constexpr (sizeof(T) == 4) {
          const float4 q = *reinterpret_cast<const float4*>(wr + e);
          wv[0] = q.x; wv[1] = q.y; wv[2] = q.z; wv[3] = q.w;
        } else {
          const uint2 q = *reinterpret_cast<const uint2*>(wr + e);
          const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&q.x);
          const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&q.y);
          wv[0] = __low2float(a); wv[1] = __high2float(a); wv[2] = __low2float(b); wv[3] = __high2float(b);
        }
#pragma unroll
        for (int n = 0; n < LS_NB; ++n) {
          if (n < nb) {
            const float4 xv = *reinterpret_cast<const float4*>(xin + n * K + e);
            acc[n] = fmaf(wv[0], xv.x, fmaf(wv[1], xv.y, fmaf(wv[2], xv.z, fmaf(wv[3], xv.w, acc[n]))));
          }
        }
      }
    } else {
      for (int e = lane; e < K; e += 32) {
        const float wv = to_f32<T>(wr[e]);
#pragma unroll
        for (int n = 0; n < LS_NB; ++n)
          if (n < nb) acc[n] = fmaf(wv, xin[n * K + e], acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < LS_NB; ++n) {
      float a = acc[n];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0 && n < nb) y[(long long)(n0 + n) * p.Cout + m] = from_f32<T>(a + (p.bias ? p.bias[m] : 0.f));
    }
  }
}

template <typename T>
static int launch_linear(LinParams& p, cudaStream_t st) {
  const int K = p.Cin * p.k;
  const size_t smem = (size_t)LS_NB * K * sizeof(float);
  WNB_CHECK_ARG(smem <= 200 * 1024, "linear: Cin * k = %d does not fit the shared staging (<= 6400)", K);
  if (smem > 48 * 1024) WNB_SET_SMEM_ATTR((int)smem, linear_frame_kernel<T>);
  // W spread over every SM: a CTA takes ceil(Cout / 148) rows (one per warp, up to 8 at a time)
  int ctas = p.Cout < 148 ? p.Cout : 148;
  p.rows_per_cta = ceil_div(p.Cout, ctas);
  dim3 grid(ceil_div(p.Cout, p.rows_per_cta), ceil_div(p.N, LS_NB));
  linear_frame_kernel<T><<<grid, 256, smem, st>>>(p);
  WNB_LAUNCH_OK();
  return 0;
}


// ------------------------------------------------------------------------------------------ tensor-core form (NLC rows)
// On the tcgen05 path (bf16 inference, bytenet_tc.py) activations are NLC: a frame's channels are contiguous, LayerNorm
// is a row reduction.  One warp per frame row, the row in registers (C <= 1024: 4 x 16-byte vectors per lane), two
// passes (mean, then the unbiased variance about it), ReLU, one 16-byte store per vector.
constexpr int LR_MAXV = 4;

// VPR = 16-byte vectors per lane and row (C <= 256 * VPR); a warp works on 4 / VPR rows at once so that every lane
// always has 4 loads in flight (a single 512-byte row per warp left the kernel at 0.39 of the copy peak; 8 loads per lane
// cost the occupancy they were meant to use: 110-128 registers, 0.53).
template <int VPR>
__global__ void __launch_bounds__(256) lnrelu_rows_kernel(long long rows, int C, const __nv_bfloat16* __restrict__ x,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float eps, __nv_bfloat16* __restrict__ y) {
  constexpr int RPW = LR_MAXV / VPR;             // rows per warp pass
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 3;                       // 16-byte vectors per row
  // (one pass per warp, many short CTAs: a grid-stride loop over the row groups measured 0.52 of the copy peak against
  // 0.72 -- the next group's loads do not start before this group's stores have been issued)
  const long long row0 = (blockIdx.x * 8ll + (threadIdx.x >> 5)) * RPW;
  if (row0 >= rows) return;
  uint4 raw[LR_MAXV];
  float s[RPW], var[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    s[r] = 0.f;
#pragma unroll
    for (int j = 0; j < VPR; ++j) {
      const int v = lane + 32 * j;
      const int i = r * VPR + j;
      raw[i] = make_uint4(0, 0, 0, 0);
      if (v < nvec && row0 + r < rows) raw[i] = reinterpret_cast<const uint4*>(x + (row0 + r) * C)[v];
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
#pragma unroll
    for (int j = 0; j < VPR; ++j) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[r * VPR + j]);
#pragma unroll
      for (int q = 0; q < 4; ++q) s[r] += __low2float(h[q]) + __high2float(h[q]);      // vectors past the row hold zeros
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
    s[r] /= (float)C;
    var[r] = 0.f;
#pragma unroll
    for (int j = 0; j < VPR; ++j) {
      if (lane + 32 * j < nvec) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[r * VPR + j]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float a = __low2float(h[q]) - s[r], b = __high2float(h[q]) - s[r];
          var[r] += a * a + b * b;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var[r] += __shfl_xor_sync(0xffffffffu, var[r], o);
    var[r] = 1.f / (sqrtf(var[r] / (float)(C - 1)) + eps);          // unbiased std, eps on the std (layernorm.py:27)
  }
#pragma unroll
  for (int j = 0; j < VPR; ++j) {
    const int v = lane + 32 * j;
    if (v >= nvec) continue;
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8), g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8), b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      if (row0 + r >= rows) continue;
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[r * VPR + j]);
      uint4 o;
      __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a = fmaxf(gg[2 * q] * (__low2float(h[q]) - s[r]) * var[r] + bb[2 * q], 0.f);
        const float b = fmaxf(gg[2 * q + 1] * (__high2float(h[q]) - s[r]) * var[r] + bb[2 * q + 1], 0.f);
        oh[q] = __floats2bfloat162_rn(a, b);
      }
      reinterpret_cast<uint4*>(y + (row0 + r) * C)[v] = o;
    }
  }
}

__device__ __forceinline__ float bn_tanh_approx(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float bn_sigmoid_approx(float v) { return fmaf(bn_tanh_approx(0.5f * v), 0.5f, 0.5f); }

// MultiplicativeUnit gate on NLC rows: pre-activations of unit u at pre[u] + row * pitch + c (four tensors, or four
// column blocks of one), h and out [rows, C].  8 channels per thread, 16-byte accesses.
struct MuRows {
  const __nv_bfloat16* pre[4];
  long long pitch;
};

__global__ void __launch_bounds__(256) mu_gate_rows_kernel(long long rows, int C, const MuRows m,
                                                           const __nv_bfloat16* __restrict__ h,
                                                           __nv_bfloat16* __restrict__ out) {
  const int nvec = C >> 3;
  const long long total = rows * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / nvec;
    const int c = (int)(i - row * nvec) * 8;
    uint4 p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) p[u] = *reinterpret_cast<const uint4*>(m.pre[u] + row * m.pitch + c);
    const uint4 hv = *reinterpret_cast<const uint4*>(h + row * C + c);
    uint4 o;
    __nv_bfloat16* oe = reinterpret_cast<__nv_bfloat16*>(&o);
    const __nv_bfloat16* he = reinterpret_cast<const __nv_bfloat16*>(&hv);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      // bf16 in / out: the MUFU approximations (5e-4 absolute) are below the output's rounding (4e-3 relative)
      const float g1 = bn_sigmoid_approx(__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&p[0])[q]));
      const float g2 = bn_sigmoid_approx(__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&p[1])[q]));
      const float g3 = bn_sigmoid_approx(__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&p[2])[q]));
      const float u = bn_tanh_approx(__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&p[3])[q]));
      oe[q] = __float2bfloat16_rn(g1 * bn_tanh_approx(g2 * __bfloat162float(he[q]) + g3 * u));
    }
    *reinterpret_cast<uint4*>(out + row * C + c) = o;
  }
}

// out[b, p * Cp + c, t] = residual[b, p * Cp + c, t] + part_p[b, t, c]: the block's last 1x1 leaves the tensor-core
// path as up to 4 NLC column blocks (N <= 256 per contraction); this is the layout change back to NCL with the
// `seq +` of block.py:119,166 riding on it.  64 x 64 tiles through shared memory, 16-byte accesses on both sides.
struct PartsArgs {
  const __nv_bfloat16* part[4];
  int nparts, Cp;
};

__global__ void __launch_bounds__(256) nlc_parts_to_ncl_add_kernel(int C, int Tn, const PartsArgs a,
                                                                   const __nv_bfloat16* __restrict__ residual,
                                                                   __nv_bfloat16* __restrict__ out) {
  __shared__ uint32_t tile[64][33];               // [frame][channel pair], odd pitch: conflict-free writes, 2-way reads
  const int b = blockIdx.z, c0 = blockIdx.y * 64, t0 = blockIdx.x * 64;
  const int part = c0 / a.Cp, cp0 = c0 - part * a.Cp;
  const __nv_bfloat16* src = a.part[part] + ((long long)b * Tn) * a.Cp + cp0;
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {       // a frame's 64 channels = 8 vectors of 16 bytes
    const int f = i >> 3, v = i & 7;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (t0 + f < Tn) q = *reinterpret_cast<const uint4*>(src + (long long)(t0 + f) * a.Cp + v * 8);
    tile[f][v * 4 + 0] = q.x; tile[f][v * 4 + 1] = q.y; tile[f][v * 4 + 2] = q.z; tile[f][v * 4 + 3] = q.w;
  }
  __syncthreads();
  const bool vec = (Tn & 7) == 0;
  {
    const int cp = threadIdx.x >> 3, fv = (threadIdx.x & 7) * 8;     // channel pair, first of 8 frames
    if (t0 + fv < Tn) {
      uint32_t w[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) w[q] = tile[fv + q][cp];
#pragma unroll
      for (int hch = 0; hch < 2; ++hch) {
        const long long o = ((long long)b * C + c0 + 2 * cp + hch) * Tn + t0 + fv;
        float v8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const __nv_bfloat162 pr = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
          v8[q] = hch ? __high2float(pr) : __low2float(pr);
        }
        if (vec) {
          uint4 rq = make_uint4(0, 0, 0, 0);
          if (residual) rq = *reinterpret_cast<const uint4*>(residual + o);
          const __nv_bfloat162* re = reinterpret_cast<const __nv_bfloat162*>(&rq);
          uint4 ov;
          __nv_bfloat162* oe = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            oe[q] = residual ? __floats2bfloat162_rn(__low2float(re[q]) + v8[2 * q], __high2float(re[q]) + v8[2 * q + 1])
                             : __floats2bfloat162_rn(v8[2 * q], v8[2 * q + 1]);    // pure layout change (keeps -0)
          *reinterpret_cast<uint4*>(out + o) = ov;
        } else {
          for (int q = 0; q < 8 && t0 + fv + q < Tn; ++q)
            out[o + q] = __float2bfloat16_rn(residual ? __bfloat162float(residual[o + q]) + v8[q] : v8[q]);
        }
      }
    }
  }
}


// AvgPool1d(P) fused with NCL -> NLC (classifier.py:53,102 feeding the tensor-core stack), bf16 rows that start 16-byte
// aligned: a thread pools 8 output frames of TWO adjacent channels from P 16-byte vectors per row (eight lanes cover
// 128 P contiguous bytes of a row), the (c, c+1) words cross the [64][33] tile and leave as 16-byte vectors of 8 channels.
template <int P>
__global__ void __launch_bounds__(256) avgpool_ncl_to_nlc_v3_kernel(int C, int Tn, int To, bool f16,
                                                                    const __nv_bfloat16* __restrict__ x,
                                                                    __nv_bfloat16* __restrict__ y) {
  __shared__ uint32_t tile[64][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 64, t0 = blockIdx.x * 64;       // t0: pooled frames
  const int cp = threadIdx.x >> 3, fv = (threadIdx.x & 7) * 8;
  const float inv = 1.f / (float)P;
  float pooled[2][8];
#pragma unroll
  for (int hch = 0; hch < 2; ++hch) {
    const __nv_bfloat16* row = x + ((long long)b * C + c0 + 2 * cp + hch) * Tn + (long long)(t0 + fv) * P;
    float v[8 * P];
    if ((long long)(t0 + fv + 8) * P <= Tn) {
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(row) + j);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[8 * j + 2 * k] = __uint_as_float(w[k] << 16);
          v[8 * j + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8 * P; ++i) v[i] = ((long long)(t0 + fv) * P + i < Tn) ? __bfloat162float(row[i]) : 0.f;
    }
#pragma unroll
    for (int f = 0; f < 8; ++f) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < P; ++k) a += v[f * P + k];
      pooled[hch][f] = a * inv;
    }
  }
#pragma unroll
  for (int f = 0; f < 8; ++f) {
    uint32_t w;
    if (f16) {
      const __half2 h2 = __floats2half2_rn(pooled[0][f], pooled[1][f]);
      w = *reinterpret_cast<const uint32_t*>(&h2);
    } else {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(pooled[0][f], pooled[1][f]);
      w = *reinterpret_cast<const uint32_t*>(&h2);
    }
    tile[fv + f][cp] = w;
  }
  __syncthreads();
#pragma unroll
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {
    const int f = i >> 3, vv = i & 7;
    if (t0 + f < To) {
      const uint4 o = make_uint4(tile[f][4 * vv], tile[f][4 * vv + 1], tile[f][4 * vv + 2], tile[f][4 * vv + 3]);
      *reinterpret_cast<uint4*>(y + ((long long)b * To + t0 + f) * C + c0 + 8 * vv) = o;
    }
  }
}

// host side of the two fast paths that other translation units route to (dense_tc.cu, chain_tc.cu)
int avgpool_ncl_to_nlc_v3_launch(int B, int C, int T_, int pool, bool f16, const void* x, void* y, cudaStream_t st) {
  const int To = T_ / pool;
  dim3 grid(ceil_div(To, 64), C / 64, B);
  const __nv_bfloat16* xs = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* ys = reinterpret_cast<__nv_bfloat16*>(y);
  switch (pool) {
    case 1: avgpool_ncl_to_nlc_v3_kernel<1><<<grid, 256, 0, st>>>(C, T_, To, f16, xs, ys); break;
    case 2: avgpool_ncl_to_nlc_v3_kernel<2><<<grid, 256, 0, st>>>(C, T_, To, f16, xs, ys); break;
    case 3: avgpool_ncl_to_nlc_v3_kernel<3><<<grid, 256, 0, st>>>(C, T_, To, f16, xs, ys); break;
    case 4: avgpool_ncl_to_nlc_v3_kernel<4><<<grid, 256, 0, st>>>(C, T_, To, f16, xs, ys); break;
    default: return -1;
  }
  return 0;
}

int nlc_to_ncl_bf16_fast_launch(int B, int C, int T_, const void* x, void* y, cudaStream_t st) {
  PartsArgs a;
  a.nparts = 1; a.Cp = C;
  a.part[0] = reinterpret_cast<const __nv_bfloat16*>(x);
  a.part[1] = a.part[2] = a.part[3] = nullptr;
  dim3 grid(ceil_div(T_, 64), C / 64, B);
  nlc_parts_to_ncl_add_kernel<<<grid, 256, 0, st>>>(C, T_, a, nullptr, reinterpret_cast<__nv_bfloat16*>(y));
  return 0;
}

}  // namespace wnb

using namespace wnb;
using bf16 = __nv_bfloat16;

#define BN_DISPATCH(dtype, NAME, ...)                                        \
  do {                                                                       \
    if ((dtype) == WNB200_F32) { using T = float; __VA_ARGS__; }             \
    else if ((dtype) == WNB200_BF16) { using T = bf16; __VA_ARGS__; }        \
    else { set_error(NAME ": bad dtype %d", (int)(dtype)); return 1; }       \
  } while (0)

extern "C" int wnb200_ln_stats(int dtype, int B, int C, int T_, const void* x, float eps, float* stats, void* stream) {
  if (B == 0 || T_ == 0) return 0;
  WNB_CHECK_ARG(x && stats && C >= 2, "ln_stats: bad args (C = %d)", C);
  WNB_CHECK_ARG(B <= 65535, "ln_stats: batch %d > 65535", B);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(T_, LN_FR), B);
  BN_DISPATCH(dtype, "ln_stats", (ln_stats_kernel<T><<<grid, 256, 0, st>>>(C, T_, (const T*)x, eps, stats)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_ln_relu_fwd(int dtype, int B, int C, int T_, const void* x, const float* stats, const float* gamma,
                                  const float* beta, void* y, void* stream) {
  if (B == 0 || T_ == 0) return 0;
  WNB_CHECK_ARG(x && stats && gamma && beta && y, "ln_relu_fwd: null pointer");
  WNB_CHECK_ARG(B <= 65535, "ln_relu_fwd: batch %d > 65535", B);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(T_, LN_FR), B);
  BN_DISPATCH(dtype, "ln_relu_fwd",
              (ln_relu_fwd_kernel<T><<<grid, 256, 0, st>>>(C, T_, (const T*)x, stats, gamma, beta, (T*)y)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_ln_relu_bwd(int dtype, int B, int C, int T_, const void* x, const float* stats, const float* gamma,
                                  const float* beta, float eps, const void* dy, void* dx, float* dgamma, float* dbeta,
                                  void* stream) {
  if (B == 0 || T_ == 0) return 0;
  WNB_CHECK_ARG(x && stats && gamma && beta && dy, "ln_relu_bwd: null pointer");
  WNB_CHECK_ARG(B <= 65535, "ln_relu_bwd: batch %d > 65535", B);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(T_, LN_FR), B);
  BN_DISPATCH(dtype, "ln_relu_bwd",
              (ln_relu_bwd_kernel<T><<<grid, 256, 0, st>>>(C, T_, (const T*)x, stats, gamma, beta, eps, (const T*)dy,
                                                          (T*)dx, dgamma, dbeta, 1)));
  WNB_LAUNCH_OK();
  return 0;
}

static int linear_common(const char* what, int dtype, int N, int Cin, int Cout, int k) {
  WNB_CHECK_ARG(dtype == WNB200_F32 || dtype == WNB200_BF16, "%s: bad dtype %d", what, dtype);
  WNB_CHECK_ARG(N >= 0 && Cin >= 1 && Cout >= 1, "%s: bad shape N=%d Cin=%d Cout=%d", what, N, Cin, Cout);
  WNB_CHECK_ARG(k >= 1 && k <= LS_MAXK, "%s: kernel width %d outside [1, %d]", what, k, LS_MAXK);
  WNB_CHECK_ARG(N <= 65535 * LS_NB, "%s: batch too large", what);
  return 0;
}

extern "C" int wnb200_linear_frame(int dtype, int N, int Cin, int Cout, int k, int dilation, const void* w,
                                   const float* bias, const void* frame, void* y, void* stream) {
  if (linear_common("linear_frame", dtype, N, Cin, Cout, k)) return 1;
  if (N == 0) return 0;
  WNB_CHECK_ARG(w && frame && y && dilation >= 1, "linear_frame: bad args");
  const int rf = k + (dilation - 1) * (k - 1);
  const size_t es = dtype == WNB200_F32 ? 4 : 2;
  LinParams p;
  memset(&p, 0, sizeof(p));
  for (int j = 0; j < k; ++j) p.tap[j] = (const char*)frame + (size_t)j * dilation * es;   // get_ker_ixs: j * d
  p.sn = (long long)Cin * rf; p.sc = rf;
  p.N = N; p.Cin = Cin; p.Cout = Cout; p.k = k; p.w = w; p.bias = bias; p.y = y;
  cudaStream_t st = (cudaStream_t)stream;
  BN_DISPATCH(dtype, "linear_frame", return launch_linear<T>(p, st));
}

extern "C" int wnb200_linear_step(int dtype, int N, int Cin, int Cout, int k, int dilation, int64_t step, const void* w,
                                  const float* bias, const void* x, void* hist, void* y, void* stream) {
  if (linear_common("linear_step", dtype, N, Cin, Cout, k)) return 1;
  if (N == 0) return 0;
  WNB_CHECK_ARG(w && x && y && dilation >= 1 && step >= 0, "linear_step: bad args");
  WNB_CHECK_ARG(hist || k == 1, "linear_step: a kernel wider than 1 needs its history ring");
  const long long R = (long long)(k - 1) * dilation + 1;        // ring length = receptive field
  const size_t es = dtype == WNB200_F32 ? 4 : 2;
  const size_t slot_bytes = (size_t)N * Cin * es;
  LinParams p;
  memset(&p, 0, sizeof(p));
  for (int j = 0; j < k; ++j) {
    const long long back = (long long)(k - 1 - j) * dilation;   // tap j reads frame step - back (conv_ops.py:39-44)
    if (back == 0) p.tap[j] = x;
    else {
      // frames before the first one are the ring's initial zeros: slot of frame (step - back), for negative frame
      // numbers the slot has not been written yet
      const long long slot = ((step - back) % R + R) % R;
      p.tap[j] = (const char*)hist + (size_t)slot * slot_bytes;
    }
  }
  p.sn = Cin; p.sc = 1;
  p.N = N; p.Cin = Cin; p.Cout = Cout; p.k = k; p.w = w; p.bias = bias; p.y = y;
  if (k > 1) {
    p.ring_slot = (char*)hist + (size_t)(step % R) * slot_bytes;
    p.new_frame = x;
  }
  cudaStream_t st = (cudaStream_t)stream;
  BN_DISPATCH(dtype, "linear_step", return launch_linear<T>(p, st));
}

extern "C" int wnb200_lnrelu_rows(int64_t rows, int C, const void* x, const float* gamma, const float* beta, float eps,
                                  void* y, void* stream) {
  if (rows == 0) return 0;
  WNB_CHECK_ARG(x && gamma && beta && y, "lnrelu_rows: null pointer");
  WNB_CHECK_ARG(C >= 8 && C % 8 == 0 && C <= 256 * LR_MAXV, "lnrelu_rows: C=%d must be a multiple of 8 <= %d", C, 256 * LR_MAXV);
  WNB_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                  reinterpret_cast<uintptr_t>(beta)) & 15) == 0, "lnrelu_rows: pointers must be 16-byte aligned");
  const int vpr = C <= 256 ? 1 : (C <= 512 ? 2 : 4);
  const long long ctas = (rows + 8 * (LR_MAXV / vpr) - 1) / (8 * (LR_MAXV / vpr));
  WNB_CHECK_ARG(ctas < (1ll << 31), "lnrelu_rows: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (vpr == 1) lnrelu_rows_kernel<1><<<(unsigned)ctas, 256, 0, st>>>(rows, C, (const bf16*)x, gamma, beta, eps, (bf16*)y);
  else if (vpr == 2) lnrelu_rows_kernel<2><<<(unsigned)ctas, 256, 0, st>>>(rows, C, (const bf16*)x, gamma, beta, eps, (bf16*)y);
  else lnrelu_rows_kernel<4><<<(unsigned)ctas, 256, 0, st>>>(rows, C, (const bf16*)x, gamma, beta, eps, (bf16*)y);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_mu_gate_rows(int64_t rows, int C, const void* pre0, const void* pre1, const void* pre2,
                                   const void* pre3, int64_t pre_pitch, const void* h, void* out, void* stream) {
  if (rows == 0) return 0;
  WNB_CHECK_ARG(pre0 && pre1 && pre2 && pre3 && h && out, "mu_gate_rows: null pointer");
  WNB_CHECK_ARG(C >= 8 && C % 8 == 0 && pre_pitch >= C && pre_pitch % 8 == 0, "mu_gate_rows: C=%d pitch=%lld", C,
                (long long)pre_pitch);
  MuRows m;
  m.pre[0] = (const bf16*)pre0; m.pre[1] = (const bf16*)pre1; m.pre[2] = (const bf16*)pre2; m.pre[3] = (const bf16*)pre3;
  m.pitch = pre_pitch;
  const long long total = rows * (C / 8);
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  mu_gate_rows_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(rows, C, m, (const bf16*)h, (bf16*)out);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_nlc_parts_to_ncl_add(int B, int C, int T_, int nparts, int Cp, const void* const* parts,
                                           const void* residual, void* out, void* stream) {
  if (B == 0 || T_ == 0) return 0;
  WNB_CHECK_ARG(parts && out, "nlc_parts_to_ncl_add: null pointer");      // residual == NULL: layout change only
  WNB_CHECK_ARG(nparts >= 1 && nparts <= 4 && Cp % 64 == 0 && nparts * Cp == C, "nlc_parts_to_ncl_add: C=%d = %d x %d?", C,
                nparts, Cp);
  WNB_CHECK_ARG(B <= 65535, "nlc_parts_to_ncl_add: batch too large");
  PartsArgs a;
  a.nparts = nparts; a.Cp = Cp;
  for (int i = 0; i < 4; ++i) a.part[i] = i < nparts ? (const bf16*)parts[i] : nullptr;
  for (int i = 0; i < nparts; ++i) WNB_CHECK_ARG(parts[i] != nullptr, "nlc_parts_to_ncl_add: part %d is null", i);
  dim3 grid(ceil_div(T_, 64), C / 64, B);
  nlc_parts_to_ncl_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(C, T_, a, (const bf16*)residual, (bf16*)out);
  WNB_LAUNCH_OK();
  return 0;
}
