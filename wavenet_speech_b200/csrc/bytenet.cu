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
