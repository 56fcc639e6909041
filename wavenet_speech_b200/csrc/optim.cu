// Adam step of the train step (reference legacy_code/train.py:55 `opt.step()` with torch.optim.Adam, train.py:112-114)
// over ALL parameters in one launch.  torch's multi-tensor Adam takes ~1.5 ms for the ~450 tensors of the WaveNet-CTC
// pair (19 M parameters: 0.1 ms of HBM traffic); here the tensors are cut into 16 K-element chunks, a device table
// maps a CTA to (tensor, offset) and every CTA streams its chunk with 16-byte accesses.
#include "common.cuh"

namespace wnb {

constexpr int AD_CHUNK = 16384;

struct AdamHyper {
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2_rsqrt;   // bc1 = 1 - beta1^t, bc2_rsqrt = 1 / sqrt(1 - beta2^t)
};

template <typename T>
__device__ __forceinline__ void adam_elem(T& p, const T& g_in, float& m, float& v, const AdamHyper& h) {
  float pf = to_f32<T>(p);
  float g = to_f32<T>(g_in);
  if (h.weight_decay != 0.f) g = fmaf(h.weight_decay, pf, g);
  m = h.beta1 * m + (1.f - h.beta1) * g;            // exp_avg.lerp_(grad, 1 - beta1) up to rounding
  v = h.beta2 * v + (1.f - h.beta2) * g * g;
  const float denom = sqrtf(v) * h.bc2_rsqrt + h.eps;
  pf -= (h.lr / h.bc1) * (m / denom);
  p = from_f32<T>(pf);
}

// items: per tensor (param, grad, exp_avg, exp_avg_sq, numel, is_bf16); chunks: (item, first element) per CTA
__global__ void __launch_bounds__(256) adam_multi_kernel(const wnb200_adam_item_t* __restrict__ items,
                                                         const int2* __restrict__ chunks, const AdamHyper h) {
  const int2 ck = chunks[blockIdx.x];
  const wnb200_adam_item_t it = items[ck.x];
  const long long e0 = (long long)ck.y;
  const long long n = it.numel - e0 < AD_CHUNK ? it.numel - e0 : AD_CHUNK;
  float* m = it.exp_avg + e0;
  float* v = it.exp_avg_sq + e0;
  if (it.is_bf16 == 0) {
    float* p = reinterpret_cast<float*>(it.param) + e0;
    const float* g = reinterpret_cast<const float*>(it.grad) + e0;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec) {
      const long long n4 = n >> 2;
      for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        adam_elem<float>(pp.x, gg.x, mm.x, vv.x, h);
        adam_elem<float>(pp.y, gg.y, mm.y, vv.y, h);
        adam_elem<float>(pp.z, gg.z, mm.z, vv.z, h);
        adam_elem<float>(pp.w, gg.w, mm.w, vv.w, h);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
      }
      for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) adam_elem<float>(p[i], g[i], m[i], v[i], h);
    } else {
      for (long long i = threadIdx.x; i < n; i += blockDim.x) adam_elem<float>(p[i], g[i], m[i], v[i], h);
    }
  } else {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(it.param) + e0;
    const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(it.grad) + e0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) adam_elem<__nv_bfloat16>(p[i], g[i], m[i], v[i], h);
  }
}

}  // namespace wnb

using namespace wnb;

extern "C" int wnb200_adam_chunk_elems(void) { return AD_CHUNK; }

extern "C" int wnb200_adam_step(int nchunks, const wnb200_adam_item_t* items, const int32_t* chunks, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int64_t step, void* stream) {
  if (nchunks == 0) return 0;
  WNB_CHECK_ARG(nchunks > 0 && items && chunks, "adam_step: bad arguments");
  WNB_CHECK_ARG(step >= 1, "adam_step: step counts from 1");
  AdamHyper h;
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay;
  h.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  h.bc2_rsqrt = (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)step)));
  adam_multi_kernel<<<nchunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      items, reinterpret_cast<const int2*>(chunks), h);
  WNB_LAUNCH_OK();
  return 0;
}
