// On-device synthetic Gaussian pore-model signal: the generator the reference runs on the CPU with
// numpy / scipy.ndimage.generic_filter (utils/raw_signal_generator.py:77-118,189-203; bases ~ U{1..4} as
// utils/gaussian_kmer_model.py:281) and its mu-law quantisation + one-hot encoding (utils/pore_model.py:58-62,78-96).
//
//   siggen_raw_kernel   one CTA per read: bases -> 5-mer ids (sum (nt-1) * [256,64,16,4,1] over a sliding window)
//                       -> samples per k-mer = max(1, int(Gamma(2.461964, 1/587.2858) * 800)) (Marsaglia-Tsang)
//                       -> block prefix sum -> sample t belongs to the first k-mer whose cumulative count exceeds t
//                       (binary search in shared memory) -> pA sample = mean[k] + stdv[k] * z_t, z_t ~ N(0,1).
//                       Every random draw (bases, counts, z) can be written out, so the deterministic part is checked
//                       exactly against the numpy restatement on the same draws.
//   siggen_onehot_kernel  per read: (x - mean) / (max - min), sign(x) log1p(mu |x|) / log1p(mu), digitize on
//                       linspace(-1, 1, levels) (double precision, as numpy), one-hot [B, levels, T].
// Random numbers: counter-based (a 64-bit mix of seed, stream, read, index, attempt) -- no state, reproducible for any
// launch geometry.
#include <math.h>

#include "common.cuh"

namespace wnb {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// uniform in (0, 1) from (seed, stream, read, index, attempt)
__device__ __forceinline__ float urand(unsigned long long seed, unsigned stream, unsigned b, unsigned i, unsigned k) {
  unsigned long long h = mix64(seed + 0x9E3779B97F4A7C15ull * (stream + 1));
  h = mix64(h ^ ((unsigned long long)b << 32 | i));
  h = mix64(h + k);
  return ((float)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float nrand(unsigned long long seed, unsigned stream, unsigned b, unsigned i, unsigned k) {
  const float u1 = urand(seed, stream, b, i, 2 * k), u2 = urand(seed, stream, b, i, 2 * k + 1);
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

constexpr float SG_SHAPE = 2.461964f, SG_RATE = 587.2858f, SG_FS = 800.f;

__global__ void __launch_bounds__(1024)
siggen_raw_kernel(int T, int nbases, unsigned long long seed, const float* __restrict__ means,
                  const float* __restrict__ stdvs, int* bases, int* reps, int* n_used, float* zout, float* sig) {
  extern __shared__ int cum[];                  // [nk] inclusive prefix sums of the per-k-mer sample counts
  __shared__ int wsum[32];
  __shared__ int carry;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int nk = nbases - 4;
  int* bs = bases + (long long)b * nbases;
  int* rp = reps + (long long)b * nk;
  for (int i = tid; i < nbases; i += blockDim.x)
    bs[i] = 1 + (int)(urand(seed, 0, b, i, 0) * 4.f);
  if (tid == 0) carry = 0;
  __syncthreads();
  // sample counts (Marsaglia & Tsang 2000, shape >= 1), then their prefix sum chunk by chunk
  const float d = SG_SHAPE - 1.f / 3.f, c = rsqrtf(9.f * d);
  for (int i0 = 0; i0 < nk; i0 += blockDim.x) {
    const int i = i0 + tid;
    int r = 0;
    if (i < nk) {
      float g = d;
      for (unsigned k = 0; k < 64; ++k) {
        const float x = nrand(seed, 1, b, i, k), u = urand(seed, 2, b, i, k);
        float v = 1.f + c * x;
        if (v <= 0.f) continue;
        v = v * v * v;
        if (logf(u) < 0.5f * x * x + d - d * v + d * logf(v)) { g = d * v; break; }
      }
      r = max(1, (int)(g / SG_RATE * SG_FS));
      rp[i] = r;
    }
    int incl = r;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int v = lane < nw ? wsum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      wsum[lane] = v;
    }
    __syncthreads();
    if (i < nk) cum[i] = carry + (warp ? wsum[warp - 1] : 0) + incl;
    __syncthreads();
    if (tid == 0) carry += wsum[nw - 1];
    __syncthreads();
  }
  const int total = carry;
  if (tid == 0) {
    int nu = -1;                                // not enough bases for T samples: the host retries with more
    if (total >= T) {
      int lo = 0, hi = nk - 1;                  // first i with cum[i] >= T  (numpy.searchsorted(cumsum, T, 'left'))
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cum[mid] >= T) hi = mid; else lo = mid + 1;
      }
      nu = lo + 1;
    }
    n_used[b] = nu;
  }
  if (total < T) return;
  for (int t = tid; t < T; t += blockDim.x) {
    int lo = 0, hi = nk - 1;                    // first i with cum[i] > t
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cum[mid] > t) hi = mid; else lo = mid + 1;
    }
    const int i = lo;
    const int k = (bs[i] - 1) * 256 + (bs[i + 1] - 1) * 64 + (bs[i + 2] - 1) * 16 + (bs[i + 3] - 1) * 4 + (bs[i + 4] - 1);
    const float z = nrand(seed, 3, b, t, 0);
    if (zout) zout[(long long)b * T + t] = z;
    sig[(long long)b * T + t] = means[k] + stdvs[k] * z;
  }
}

template <typename TO>
__global__ void __launch_bounds__(1024)
siggen_onehot_kernel(int T, int levels, const float* __restrict__ sig, TO* onehot, long long* lev_out) {
  __shared__ double rs[32];
  __shared__ float rmx[32], rmn[32];
  __shared__ double s_mean;
  __shared__ float s_max, s_min;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const float* x = sig + (long long)b * T;
  double sum = 0.0;
  float mx = -INFINITY, mn = INFINITY;
  for (int t = tid; t < T; t += blockDim.x) {
    const float v = x[t];
    sum += (double)v;
    mx = fmaxf(mx, v);
    mn = fminf(mn, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if (lane == 0) { rs[warp] = sum; rmx[warp] = mx; rmn[warp] = mn; }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    float a = -INFINITY, c = INFINITY;
    for (int w = 0; w < nw; ++w) { s += rs[w]; a = fmaxf(a, rmx[w]); c = fminf(c, rmn[w]); }
    s_mean = s / (double)T; s_max = a; s_min = c;
  }
  __syncthreads();
  const double mean = s_mean, range = (double)s_max - (double)s_min, mu = (double)levels;
  const double step = 2.0 / (double)(levels - 1), l1p = log1p(mu);
  TO* oh = onehot + (long long)b * levels * T;
  for (int t = tid; t < T; t += blockDim.x) {
    const double n = ((double)x[t] - mean) / range;
    const double m = (n > 0.0 ? 1.0 : (n < 0.0 ? -1.0 : 0.0)) * log1p(mu * fabs(n)) / l1p;
    // numpy.digitize(m, linspace(-1, 1, levels)): number of edges <= m, clipped into [0, levels-1]
    int lev = (int)floor((m + 1.0) / step) + 1;
    while (lev > 0 && -1.0 + step * (double)(lev - 1) > m) --lev;          // guard the floating-point edge cases
    while (lev < levels && -1.0 + step * (double)lev <= m) ++lev;
    lev = min(max(lev, 0), levels - 1);
    if (lev_out) lev_out[(long long)b * T + t] = lev;
    for (int c = 0; c < levels; ++c) oh[(long long)c * T + t] = from_f32<TO>(c == lev ? 1.f : 0.f);
  }
}

}  // namespace wnb

using namespace wnb;

extern "C" int wnb200_siggen_raw(int B, int T_, int nbases, uint64_t seed, const float* means, const float* stdvs,
                                 int32_t* bases, int32_t* reps, int32_t* n_used, float* z, float* sig, void* stream) {
  WNB_CHECK_ARG(nbases >= 5 && T_ >= 1, "siggen_raw: need at least 5 bases and one sample");
  if (B == 0) return 0;
  WNB_CHECK_ARG(means && stdvs && bases && reps && n_used && sig, "siggen_raw: null pointer");
  const size_t smem = sizeof(int) * (size_t)(nbases - 4);
  WNB_CHECK_ARG(smem <= 200 * 1024, "siggen_raw: %d bases per read do not fit in shared memory", nbases);
  if (smem > 48 * 1024)
    WNB_CUDA_OK(cudaFuncSetAttribute(siggen_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  siggen_raw_kernel<<<B, 1024, smem, (cudaStream_t)stream>>>(T_, nbases, (unsigned long long)seed, means, stdvs, bases,
                                                            reps, n_used, z, sig);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_siggen_onehot(int dtype, int B, int T_, int levels, const float* sig, void* onehot,
                                    int64_t* lev_out, void* stream) {
  WNB_CHECK_ARG(levels >= 2 && T_ >= 1, "siggen_onehot: bad sizes");
  if (B == 0) return 0;
  WNB_CHECK_ARG(sig && onehot, "siggen_onehot: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == WNB200_F32)
    siggen_onehot_kernel<float><<<B, 1024, 0, st>>>(T_, levels, sig, (float*)onehot, (long long*)lev_out);
  else if (dtype == WNB200_BF16)
    siggen_onehot_kernel<__nv_bfloat16><<<B, 1024, 0, st>>>(T_, levels, sig, (__nv_bfloat16*)onehot,
                                                           (long long*)lev_out);
  else
    WNB_CHECK_ARG(false, "siggen_onehot: bad dtype %d", dtype);
  WNB_LAUNCH_OK();
  return 0;
}
