// CTC loss on the device (reference: warpctc_pytorch.CTCLoss as called by legacy_code/train.py:42-46,
// legacy_code/run_raw_ctc.py:59-62, Loss.py:50-53): activations are PRE-softmax, label 0 is the blank, the loss is
// summed over the batch, the gradient is taken with respect to the activations.
//
//   ctc_logprobs_kernel   lp[b, t, c] = log_softmax_c(act[b, c, t])                      (fully parallel)
//   ctc_alpha_beta_kernel one CTA per (read, direction): the forward (alpha) or backward (beta) recursion over
//                         the 2*len+1 extended label states.  The whole state vector of the previous frame sits in
//                         shared memory (double buffered), each thread owns states tid, tid+blockDim, ...; one
//                         __syncthreads per frame.  Both trellises are written to HBM ([B, T, S] fp32 each) and
//                         nll[b] = -log p(labels | act) comes from the last alpha frame.  Every CTC_RENORM frames the
//                         frame maximum is taken out of the state vector and added to a double-precision offset
//                         (stored per frame): log-domain values stay O(50) instead of O(T), which keeps the fp32
//                         rounding of long reads out of the gradient (an fp32 log-domain CTC without it -- warp-ctc,
//                         torch -- is ~1e-2 off an fp64 evaluation at T = 1500).
//   ctc_grad_kernel       d act[b, c, t] = scale * ( softmax[b, c, t] - sum_{s: lab(s) = c} w_t(s) ),
//                         w_t(s) = exp(alpha_t(s) + beta_t(s) - lp_t(lab(s)) + offsets_t + nll)  -- the w_t(s) of one frame lie in
//                         [0, 1] and sum to 1, so they are added in the linear domain; one warp per (read, frame).
//
// Only the band of states that are reachable from the start AND can still reach the end is computed, stored and
// read: at frame t these are the pairs j in [len - (Ta - t), t] (both trellises share it), about half of the
// (frame, state) plane when the label sequence is about as long as the read.
//
// The activation tensor is addressed by strides, so both the reference's (T, B, C) layout and the classifier's
// native (B, C, T) output are read in place.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace wnb {

// All log-domain quantities are in base-2 units (ex2 / lg2 are the native MUFU ops).  CTC_NEG stands for -inf: it
// absorbs every finite addend, so the recursion needs no special cases.
constexpr float CTC_NEG = -1e30f;
constexpr float CTC_LOG2E = 1.4426950408889634f;
constexpr double CTC_LN2 = 0.6931471805599453;
constexpr int CTC_MAX_L = 64;       // classes (incl. blank)
constexpr int CTC_RENORM = 16;      // frames between renormalisations of the state vector

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log2(2^a + 2^b): one ex2 + one lg2
__device__ __forceinline__ float lse2(float a, float b) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  return hi + lg2f(1.f + ex2f(lo - hi));
}
// log2(2^a + 2^b + 2^c): two ex2 + one lg2 (the largest term is 2^0)
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  const float m = fmaxf(hi, c), mid = fminf(hi, c);
  return m + lg2f(1.f + ex2f(lo - m) + ex2f(mid - m));
}

template <typename T>
__global__ void ctc_logprobs_kernel(int B, int L, int Tn, const T* act, long long sb, long long sc, long long st,
                                    float* lp) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * Tn) return;
  const long long b = i / Tn, t = i - b * Tn;
  const T* p = act + b * sb + t * st;
  float m = -FLT_MAX;
  for (int c = 0; c < L; ++c) m = fmaxf(m, to_f32<T>(p[c * sc]));
  float s = 0.f;
  for (int c = 0; c < L; ++c) s += expf(to_f32<T>(p[c * sc]) - m);
  const float lz = m + logf(s);
  float* o = lp + i * L;
  for (int c = 0; c < L; ++c) o[c] = (to_f32<T>(p[c * sc]) - lz) * CTC_LOG2E;
}

// grid (B, 2): blockIdx.y = 0 alpha, 1 beta.  Each thread owns NP (blank, label) state pairs j = tid + i * blockDim:
// states 2j and 2j+1 (the last pair is the final blank alone).  A pair costs 5 MUFU ops per frame.
// Dynamic smem: 2 * (Smax + 7) floats (state vectors with guards) + 2 * CTC_MAX_L (log-probs) + 32 (maxima).
template <int NP>
__global__ void __launch_bounds__(1024)
ctc_alpha_beta_kernel(int L, int Tn, int Smax, const float* __restrict__ lp, const int* __restrict__ labels,
                      const long long* __restrict__ lab_off, const int* __restrict__ act_len, float* alpha,
                      float* beta, double* coff, double* nll_d, float* nll) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, dir = blockIdx.y;
  const int len = (int)(lab_off[b + 1] - lab_off[b]);
  const int S = 2 * len + 1;
  const int Ta = act_len ? min(act_len[b], Tn) : Tn;
  const int* lab = labels + lab_off[b];
  const int W = Smax + 7;                     // even: both buffers keep their (blank, label) pairs 8-byte aligned
  float* buf0 = sm + 2;                       // states -2 .. Smax+4 addressable; guards hold CTC_NEG
  float* buf1 = sm + W + 2;
  float* lpb = sm + 2 * W;                    // [2][CTC_MAX_L]
  float* red = lpb + 2 * CTC_MAX_L;           // [32] per-warp maxima
  double* co = coff + ((long long)dir * gridDim.x + b) * Tn;
  const int SP = Smax + 1;                    // even row pitch: (blank, label) pairs are stored as float2
  float* out = (dir == 0 ? alpha : beta) + (long long)b * Tn * SP;
  const float* lpr = lp + (long long)b * Tn * L;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (Ta <= 0) {
    if (tid == 0 && dir == 0) {
      nll[b] = (len == 0) ? 0.f : INFINITY;
      nll_d[b] = nll[b];
    }
    return;
  }
  int cls[NP];                                // class of the pair's label state
  bool skip[NP], live[NP], haslab[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int j = tid + i * nt;
    live[i] = j <= len;
    haslab[i] = j < len;
    cls[i] = haslab[i] ? min(max(lab[j], 0), L - 1) : 0;      // labels outside [0, L) are clamped: never an out-of-range read
    skip[i] = false;
    if (haslab[i]) {
      if (dir == 0) skip[i] = j >= 1 && lab[j - 1] != cls[i];            // alpha: 2j+1 <- 2j-1
      else skip[i] = j + 1 < len && lab[j + 1] != cls[i];                 // beta:  2j+1 <- 2j+3
    }
  }
  for (int s = tid - 2; s < W - 2; s += nt) { buf0[s] = CTC_NEG; buf1[s] = CTC_NEG; }
  const int tfirst = dir == 0 ? 0 : Ta - 1, tstep = dir == 0 ? 1 : -1;
  if (tid < L) lpb[tid] = lpr[(long long)tfirst * L + tid];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NP; ++i) {               // first frame: alpha starts in states 0, 1; beta in S-1, S-2
    const int j = tid + i * nt;
    if (live[i]) {
      const int s0 = 2 * j, s1 = 2 * j + 1;
      const float v0 = (dir == 0 ? s0 < 2 : s0 >= S - 2) ? lpb[0] : CTC_NEG;
      const float v1 = (haslab[i] && (dir == 0 ? s1 < 2 : s1 >= S - 2)) ? lpb[cls[i]] : CTC_NEG;
      *reinterpret_cast<float2*>(buf0 + s0) = make_float2(v0, v1);      // v1 = NEG lands on the guard after the last state
      *reinterpret_cast<float2*>(out + (long long)tfirst * SP + s0) = make_float2(v0, v1);
    }
  }
  if (tid == 0) co[tfirst] = 0.0;
  // log-probs of the NEXT frame are fetched one iteration ahead (an L2 round trip per frame otherwise sits in front
  // of the barrier: the load was issued and consumed in the same iteration)
  float nxt = 0.f;
  if (tid < L && Ta > 1) nxt = lpr[(long long)(tfirst + tstep) * L + tid];
  double C = 0.0;                             // every thread tracks the same offset
  for (int k = 1; k < Ta; ++k) {
    const int t = tfirst + k * tstep;
    const float* prev = (k & 1) ? buf0 : buf1;
    float* cur = (k & 1) ? buf1 : buf0;
    float* lpc = lpb + (k & 1) * CTC_MAX_L;
    if (tid < L) {
      lpc[tid] = nxt;
      if (k + 1 < Ta) nxt = lpr[(long long)(t + tstep) * L + tid];
    }
    __syncthreads();
    // band of useful pairs at this frame and at the previous one (see the header)
    const int jlo = max(0, len - (Ta - t)), jhi = min(len, t);
    const int tp = t - tstep;
    const int plo = max(0, len - (Ta - tp)), phi = min(len, tp);
    float m = 0.f;
    if ((k % CTC_RENORM) == 0) {              // take the frame maximum out of the state vector
      float lm = CTC_NEG;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int j = tid + i * nt;
        if (j >= plo && j <= phi) lm = fmaxf(lm, fmaxf(prev[2 * j], prev[2 * j + 1]));   // guard = NEG
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lm = fmaxf(lm, __shfl_xor_sync(0xffffffffu, lm, o));
      if ((tid & 31) == 0) red[tid >> 5] = lm;
      __syncthreads();
      m = CTC_NEG;
      for (int wq = 0; wq < (nt >> 5); ++wq) m = fmaxf(m, red[wq]);
      if (m < -1e29f) m = 0.f;
      C += (double)m;
    }
    if (tid == 0) co[t] = C;
    float* orow = out + (long long)t * SP;
    const float lpblank = lpc[0] - m;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int j = tid + i * nt;
      if (j >= jlo && j <= jhi) {
        const int s0 = 2 * j;
        float r0, r1;
        if (dir == 0) {
          const float p0 = prev[s0 - 1];
          const float2 p12 = *reinterpret_cast<const float2*>(prev + s0);
          r0 = lse2(p12.x, p0) + lpblank;
          r1 = lse3(p12.y, p12.x, skip[i] ? p0 : CTC_NEG) + (lpc[cls[i]] - m);
        } else {
          const float2 q01 = *reinterpret_cast<const float2*>(prev + s0), q23 = *reinterpret_cast<const float2*>(prev + s0 + 2);
          r0 = lse2(q01.x, q01.y) + lpblank;
          r1 = lse3(q01.y, q23.x, skip[i] ? q23.y : CTC_NEG) + (lpc[cls[i]] - m);
        }
        if (!haslab[i]) r1 = CTC_NEG;
        const float2 rr = make_float2(r0, r1);
        *reinterpret_cast<float2*>(cur + s0) = rr;
        *reinterpret_cast<float2*>(orow + s0) = rr;
      }
    }
  }
  __syncthreads();
  if (dir == 0 && tid == 0) {
    const float* last = ((Ta - 1) & 1) ? buf1 : buf0;
    const float ll = lse2(last[S - 1], last[S - 2]);            // S = 1: last[-1] is a guard
    const double n = ll < -1e29f ? (double)INFINITY : -((double)ll + C) * CTC_LN2;
    nll_d[b] = n;
    nll[b] = (float)n;
  }
}

// one warp per (read, frame)
template <typename T>
__global__ void __launch_bounds__(256)
ctc_grad_kernel(int B, int L, int Tn, int Smax, const float* __restrict__ lp, const int* __restrict__ labels,
                const long long* __restrict__ lab_off, const int* __restrict__ act_len,
                const float* __restrict__ alpha, const float* __restrict__ beta, const double* __restrict__ coff,
                const double* __restrict__ nll_d, const float* __restrict__ gscale, T* grad, long long sb,
                long long sc, long long st) {
  __shared__ float acc[8][CTC_MAX_L];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w = blockIdx.x * 8ll + warp;
  if (w >= (long long)B * Tn) return;
  const int b = (int)(w / Tn), t = (int)(w - (long long)b * Tn);
  const int len = (int)(lab_off[b + 1] - lab_off[b]);
  const int Ta = act_len ? min(act_len[b], Tn) : Tn;
  T* g = grad + b * sb + t * st;
  const double nd = nll_d[b];
  const float scale = gscale ? gscale[0] : 1.f;
  if (t >= Ta || !(nd < (double)INFINITY)) {                      // padding frames / infeasible alignment: zero gradient
    for (int c = lane; c < L; c += 32) g[c * sc] = from_f32<T>(0.f);
    return;
  }
  // offsets of the two trellises at this frame + nll (base-2 units), combined in double: O(T) magnitudes cancel here
  const float n = (float)(coff[(long long)b * Tn + t] + coff[((long long)B + b) * Tn + t] + nd / CTC_LN2);
  for (int c = lane; c < L; c += 32) acc[warp][c] = 0.f;
  __syncwarp();
  const int SP = Smax + 1;
  const float2* a = reinterpret_cast<const float2*>(alpha + ((long long)b * Tn + t) * SP);
  const float2* be = reinterpret_cast<const float2*>(beta + ((long long)b * Tn + t) * SP);
  const float* lpt = lp + ((long long)b * Tn + t) * L;
  const int* lab = labels + lab_off[b];
  // one float2 per (blank, label) pair, band of useful pairs only; blanks are one class (plain sum), labels go
  // through shared-memory atomics per class
  float blank = 0.f;
  const float lp0 = lpt[0];
  const int jlo = max(0, len - (Ta - t)), jhi = min(len, t);
  if (L <= 8) {
    // few classes (nucleotides): per-lane register accumulators, one warp reduction per class at the end
    // (shared-memory atomics from 32 lanes onto 4 addresses serialise)
    float occ[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = jlo + lane; j <= jhi; j += 32) {
      const float2 av = a[j], bv = be[j];
      blank += ex2f(av.x + bv.x - lp0 + n);
      if (j < len) {
        const int c = min(max(lab[j], 0), L - 1);
        const float wl = ex2f(av.y + bv.y - lpt[c] + n);
#pragma unroll
        for (int q = 1; q < 8; ++q) occ[q] += (c == q) ? wl : 0.f;
      }
    }
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      float v = occ[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && q < L) acc[warp][q] = v;
    }
  } else {
    for (int j = jlo + lane; j <= jhi; j += 32) {
      const float2 av = a[j], bv = be[j];
      blank += ex2f(av.x + bv.x - lp0 + n);
      if (j < len) {
        const int c = min(max(lab[j], 0), L - 1);
        atomicAdd(&acc[warp][c], ex2f(av.y + bv.y - lpt[c] + n));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) blank += __shfl_xor_sync(0xffffffffu, blank, o);
  __syncwarp();
  for (int c = lane; c < L; c += 32) {
    const float occ = acc[warp][c] + (c == 0 ? blank : 0.f);
    g[c * sc] = from_f32<T>(scale * (exp2f(lpt[c]) - occ));
  }
}

}  // namespace wnb

using namespace wnb;

// workspace: lp [B,T,L] f32 (padded to even) | alpha [B,T,S+1] f32 | beta [B,T,S+1] f32 | offsets [2,B,T] f64 | nll [B] f64
static inline size_t ctc_lp_count(int B, int L, int T_) { return ((size_t)B * T_ * L + 1) & ~(size_t)1; }
static inline size_t ctc_f32_count(int B, int L, int T_, int Smax) {
  return ctc_lp_count(B, L, T_) + 2 * (size_t)B * T_ * (Smax + 1);      // even row pitch Smax + 1
}

extern "C" size_t wnb200_ctc_workspace_bytes(int B, int L, int T_, int max_label_len) {
  const int S = 2 * max_label_len + 1;
  return sizeof(float) * ctc_f32_count(B, L, T_, S) + sizeof(double) * (2 * (size_t)B * T_ + B);
}

extern "C" int wnb200_ctc_fwd(int dtype, int B, int L, int T_, int max_label_len, const void* act, int64_t sb,
                              int64_t sc, int64_t st, const int32_t* labels, const int64_t* label_offsets,
                              const int32_t* act_lengths, float* workspace, float* nll, void* stream) {
  WNB_CHECK_ARG(L >= 1 && L <= CTC_MAX_L, "ctc_fwd: %d classes (1..%d supported)", L, CTC_MAX_L);
  WNB_CHECK_ARG(max_label_len >= 0 && max_label_len + 1 <= 8 * 1024, "ctc_fwd: label length %d too long",
                max_label_len);
  if (B == 0) return 0;
  WNB_CHECK_ARG(act && label_offsets && workspace && nll && T_ >= 0, "ctc_fwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int Smax = 2 * max_label_len + 1;
  float* lp = workspace;
  float* alpha = lp + ctc_lp_count(B, L, T_);
  float* beta = alpha + (size_t)B * T_ * (Smax + 1);
  double* coff = reinterpret_cast<double*>(workspace + ctc_f32_count(B, L, T_, Smax));
  double* nll_d = coff + 2 * (size_t)B * T_;
  WNB_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "ctc_fwd: workspace must be 8-byte aligned");
  const long long cols = (long long)B * T_;
  if (cols > 0) {
    const unsigned grid = (unsigned)ceil_div64(cols, 256);
    if (dtype == WNB200_F32)
      ctc_logprobs_kernel<float><<<grid, 256, 0, s>>>(B, L, T_, (const float*)act, sb, sc, st, lp);
    else if (dtype == WNB200_BF16)
      ctc_logprobs_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(B, L, T_, (const __nv_bfloat16*)act, sb, sc, st, lp);
    else
      WNB_CHECK_ARG(false, "ctc_fwd: bad dtype %d", dtype);
    WNB_LAUNCH_OK();
  }
  const int pairs = max_label_len + 1;
  int threads = ((pairs + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 64) threads = 64;
  const size_t smem = sizeof(float) * (2 * ((size_t)Smax + 7) + 2 * CTC_MAX_L + 32);
  const int np = (pairs + threads - 1) / threads;
#define CTC_LAUNCH(NP)                                                                                              \
  do {                                                                                                              \
    WNB_CUDA_OK(cudaFuncSetAttribute(ctc_alpha_beta_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                     (int)smem));                                                                   \
    ctc_alpha_beta_kernel<NP><<<dim3(B, 2), threads, smem, s>>>(L, T_, Smax, lp, labels,                            \
                                                               (const long long*)label_offsets, act_lengths, alpha, \
                                                               beta, coff, nll_d, nll);                             \
  } while (0)
  if (np <= 1) CTC_LAUNCH(1);
  else if (np <= 2) CTC_LAUNCH(2);
  else if (np <= 3) CTC_LAUNCH(3);
  else if (np <= 4) CTC_LAUNCH(4);
  else if (np <= 6) CTC_LAUNCH(6);
  else CTC_LAUNCH(8);
#undef CTC_LAUNCH
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_ctc_bwd(int dtype, int B, int L, int T_, int max_label_len, const int32_t* labels,
                              const int64_t* label_offsets, const int32_t* act_lengths, const float* workspace,
                              const float* nll, const float* gscale, void* grad, int64_t sb, int64_t sc, int64_t st,
                              void* stream) {
  WNB_CHECK_ARG(L >= 1 && L <= CTC_MAX_L, "ctc_bwd: %d classes (1..%d supported)", L, CTC_MAX_L);
  if (B == 0 || T_ == 0) return 0;
  WNB_CHECK_ARG(label_offsets && workspace && nll && grad, "ctc_bwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int Smax = 2 * max_label_len + 1;
  const float* lp = workspace;
  const float* alpha = lp + ctc_lp_count(B, L, T_);
  const float* beta = alpha + (size_t)B * T_ * (Smax + 1);
  const double* coff = reinterpret_cast<const double*>(workspace + ctc_f32_count(B, L, T_, Smax));
  const double* nll_d = coff + 2 * (size_t)B * T_;
  (void)nll;
  const unsigned grid = (unsigned)ceil_div64((long long)B * T_, 8);
  if (dtype == WNB200_F32)
    ctc_grad_kernel<float><<<grid, 256, 0, s>>>(B, L, T_, Smax, lp, labels, (const long long*)label_offsets,
                                                act_lengths, alpha, beta, coff, nll_d, gscale, (float*)grad, sb, sc, st);
  else if (dtype == WNB200_BF16)
    ctc_grad_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(B, L, T_, Smax, lp, labels, (const long long*)label_offsets,
                                                        act_lengths, alpha, beta, coff, nll_d, gscale, (__nv_bfloat16*)grad,
                                                        sb, sc, st);
  else
    WNB_CHECK_ARG(false, "ctc_bwd: bad dtype %d", dtype);
  WNB_LAUNCH_OK();
  return 0;
}
