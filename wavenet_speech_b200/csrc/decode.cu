// Greedy CTC decoding on the device (reference: modules/sequence_decoders.py:9-41 `argmax_decode` = per-frame argmax
// without repeat collapse; the notebooks collapse repeats and drop blanks by hand on the host,
// ipynbs/Size 1 Pore Model Check.ipynb cell 24; Decoder.py:20-35 turns the labels into strings).
//
// One CTA per read walks the frames in chunks of blockDim: per-frame argmax over the classes (ties -> lowest index,
// as torch.max), keep frame t iff label != blank and label != label of frame t-1, ordered compaction with a block
// prefix scan -> the decoded labels packed at the front of out[b, :] and their count.  The activation tensor is
// addressed by strides, so (B, C, T) network outputs and (B, T, C) / (T, B, C) views are read in place.
#include <float.h>

#include "common.cuh"

namespace wnb {

template <typename T>
__device__ __forceinline__ int frame_argmax(const T* p, long long sc, int L) {
  float best = to_f32<T>(p[0]);
  int arg = 0;
  for (int c = 1; c < L; ++c) {
    const float v = to_f32<T>(p[c * sc]);
    if (v > best) { best = v; arg = c; }
  }
  return arg;
}

template <typename T>
__global__ void __launch_bounds__(1024)
greedy_decode_kernel(int L, int Tn, const T* act, long long sb, long long sc, long long st, const int* act_len,
                     int blank, int* out, int* out_len) {
  __shared__ int wsum[32];
  __shared__ int carry_label, base;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int Ta = act_len ? min(act_len[b], Tn) : Tn;
  const T* pa = act + b * sb;
  int* po = out + (long long)b * Tn;
  if (tid == 0) { carry_label = -1; base = 0; }
  __syncthreads();
  for (int t0 = 0; t0 < Ta; t0 += blockDim.x) {
    const int t = t0 + tid;
    int lab = -1, prev = -1;
    if (t < Ta) {
      lab = frame_argmax(pa + t * st, sc, L);
      // previous frame's label: recomputed for lanes > 0 is wasteful -> take it from the neighbour lane / carry
    }
    prev = __shfl_up_sync(0xffffffffu, lab, 1);
    __shared__ int wlast[32];
    if (lane == 31) wlast[warp] = lab;
    __syncthreads();
    if (lane == 0) prev = warp == 0 ? carry_label : wlast[warp - 1];
    const int keep = (t < Ta && lab != blank && lab != prev) ? 1 : 0;
    // block exclusive scan of keep
    int incl = keep;
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
      wsum[lane] = v;                           // inclusive sums of the warp totals
    }
    __syncthreads();
    const int pos = base + (warp ? wsum[warp - 1] : 0) + incl - keep;
    if (keep) po[pos] = lab;
    __syncthreads();
    if (tid == blockDim.x - 1) {
      base += wsum[nw - 1];
      carry_label = lab;                        // chunk is full here (t0 + blockDim <= Ta) or the loop ends
    }
    __syncthreads();
  }
  if (tid == 0) out_len[b] = base;
}

template <typename T>
__global__ void frame_argmax_kernel(int B, int L, int Tn, const T* act, long long sb, long long sc, long long st,
                                    long long* out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * Tn) return;
  const long long b = i / Tn, t = i - b * Tn;
  out[i] = frame_argmax(act + b * sb + t * st, sc, L);
}

}  // namespace wnb

using namespace wnb;

extern "C" int wnb200_ctc_greedy_decode(int dtype, int B, int L, int T_, const void* act, int64_t sb, int64_t sc,
                                        int64_t st, const int32_t* act_lengths, int blank, int32_t* out_labels,
                                        int32_t* out_lengths, void* stream) {
  WNB_CHECK_ARG(L >= 1, "ctc_greedy_decode: no classes");
  if (B == 0) return 0;
  WNB_CHECK_ARG(act && out_labels && out_lengths, "ctc_greedy_decode: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  int threads = ((T_ + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 32) threads = 32;
  if (dtype == WNB200_F32)
    greedy_decode_kernel<float><<<B, threads, 0, s>>>(L, T_, (const float*)act, sb, sc, st, act_lengths, blank,
                                                      out_labels, out_lengths);
  else if (dtype == WNB200_BF16)
    greedy_decode_kernel<__nv_bfloat16><<<B, threads, 0, s>>>(L, T_, (const __nv_bfloat16*)act, sb, sc, st,
                                                              act_lengths, blank, out_labels, out_lengths);
  else
    WNB_CHECK_ARG(false, "ctc_greedy_decode: bad dtype %d", dtype);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_frame_argmax(int dtype, int B, int L, int T_, const void* act, int64_t sb, int64_t sc,
                                   int64_t st, int64_t* out, void* stream) {
  WNB_CHECK_ARG(L >= 1, "frame_argmax: no classes");
  const long long n = (long long)B * T_;
  if (n == 0) return 0;
  WNB_CHECK_ARG(act && out, "frame_argmax: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)ceil_div64(n, 256);
  if (dtype == WNB200_F32)
    frame_argmax_kernel<float><<<grid, 256, 0, s>>>(B, L, T_, (const float*)act, sb, sc, st, (long long*)out);
  else if (dtype == WNB200_BF16)
    frame_argmax_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(B, L, T_, (const __nv_bfloat16*)act, sb, sc, st,
                                                           (long long*)out);
  else
    WNB_CHECK_ARG(false, "frame_argmax: bad dtype %d", dtype);
  WNB_LAUNCH_OK();
  return 0;
}
