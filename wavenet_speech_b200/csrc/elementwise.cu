// Bandwidth-bound kernels of the hot path (NCL layout): gate backward, LeakyReLU backward, mean-pool, position
// mixing, column sums; plus the per-column fallbacks (`*_col`: one thread per (batch, frame) column) of the tiled
// channel-reduction kernels in column_ops.cu, used only when a channel count is too large for a shared-memory tile.
// No transposing copies (reshape_in / reshape_out, reference conv_ops.py:91-101) are ever materialised.
#include "common.cuh"

namespace wnb {

// ---------------------------------------------------------------- gate backward (block.py:185)
template <typename T>
__global__ void gate_bwd_kernel(int B, int C, int Tn, const T* dact, const T* th, const T* sg, T* dab) {
  const long long n = (long long)B * C * Tn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / ((long long)C * Tn);
    const long long r = i - b * (long long)C * Tn;
    const float g = to_f32<T>(dact[i]), t = to_f32<T>(th[i]), s = to_f32<T>(sg[i]);
    const long long o = b * 2ll * C * Tn + r;
    dab[o] = from_f32<T>(g * s * (1.f - t * t));
    dab[o + (long long)C * Tn] = from_f32<T>(g * t * s * (1.f - s));
  }
}

// stand-alone GatedActivationUnit.forward (block.py:184-185); inside ResidualBlock the gate is a GEMM epilogue
template <typename T>
__global__ void gate_fwd_kernel(long long n, const T* a, const T* b, T* out, T* th, T* sg) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float t = tanhf(to_f32<T>(a[i])), s = sigmoid_precise(to_f32<T>(b[i]));
    out[i] = from_f32<T>(t * s);
    if (th) th[i] = from_f32<T>(t);
    if (sg) sg[i] = from_f32<T>(s);
  }
}

// LayerNorm parameter gradients: dgamma[c] += sum_{b,t} dy * (x - mean) * r ; dbeta[c] += sum_{b,t} dy
template <typename T>
__global__ void __launch_bounds__(256) layernorm_bwd_params_kernel(int C, int Tn, const T* x, const float* stats,
                                                                   const T* dy, float* dgamma, float* dbeta) {
  const int c = blockIdx.x, b = blockIdx.y;
  const long long base = ((long long)b * C + c) * Tn;
  float sg_ = 0.f, sb_ = 0.f;
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
    const float g = to_f32<T>(dy[base + t]);
    const float mean = stats[((long long)b * Tn + t) * 2], r = stats[((long long)b * Tn + t) * 2 + 1];
    sg_ += g * (to_f32<T>(x[base + t]) - mean) * r;
    sb_ += g;
  }
  __shared__ float red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sg_ += __shfl_xor_sync(0xffffffffu, sg_, o);
    sb_ += __shfl_xor_sync(0xffffffffu, sb_, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sg_; red[1][threadIdx.x >> 5] = sb_; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bsum = 0.f;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; bsum += red[1][i]; }
    atomicAdd(&dgamma[c], a);
    atomicAdd(&dbeta[c], bsum);
  }
}

template <typename T>
__global__ void leaky_bwd_kernel(long long n, const T* dy, const T* ref, T* dx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float g = to_f32<T>(dy[i]);
    dx[i] = from_f32<T>(to_f32<T>(ref[i]) > 0.f ? g : 0.01f * g);
  }
}

// 16 bytes per thread per tensor (the scalar kernel above handles the tail / unaligned buffers)
template <typename T>
__global__ void leaky_bwd_vec_kernel(long long nv, const uint4* dy, const uint4* ref, uint4* dx) {
  constexpr int V = 16 / sizeof(T);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    const uint4 g4 = __ldg(dy + i), r4 = __ldg(ref + i);
    uint4 o4;
    const T* g = reinterpret_cast<const T*>(&g4);
    const T* r = reinterpret_cast<const T*>(&r4);
    T* o = reinterpret_cast<T*>(&o4);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float gv = to_f32<T>(g[k]);
      o[k] = from_f32<T>(to_f32<T>(r[k]) > 0.f ? gv : 0.01f * gv);
    }
    dx[i] = o4;
  }
}

// ---------------------------------------------------------------- channel softmax
template <typename T>
__global__ void softmax_fwd_kernel(int B, int C, int Tn, const T* x, T* y, int log_mode) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= (long long)B * Tn) return;
  const long long b = col / Tn, t = col - b * Tn;
  const T* xp = x + b * (long long)C * Tn + t;
  T* yp = y + b * (long long)C * Tn + t;
  float m = -INFINITY, s = 0.f;
  for (int c = 0; c < C; ++c) {
    const float v = to_f32<T>(xp[(long long)c * Tn]);
    if (v > m) { s = s * expf(m - v); m = v; }
    s += expf(v - m);
  }
  const float lse = m + logf(s);
  const float inv = 1.f / s;
  for (int c = 0; c < C; ++c) {
    const float v = to_f32<T>(xp[(long long)c * Tn]);
    yp[(long long)c * Tn] = from_f32<T>(log_mode ? (v - lse) : expf(v - m) * inv);
  }
}

template <typename T>
__global__ void softmax_bwd_kernel(int B, int C, int Tn, const T* y, const T* dy, T* dx, int log_mode) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= (long long)B * Tn) return;
  const long long b = col / Tn, t = col - b * Tn;
  const long long base = b * (long long)C * Tn + t;
  float dot = 0.f;
  for (int c = 0; c < C; ++c) {
    const float g = to_f32<T>(dy[base + (long long)c * Tn]);
    dot += log_mode ? g : g * to_f32<T>(y[base + (long long)c * Tn]);
  }
  for (int c = 0; c < C; ++c) {
    const float g = to_f32<T>(dy[base + (long long)c * Tn]);
    const float yy = to_f32<T>(y[base + (long long)c * Tn]);
    dx[base + (long long)c * Tn] = from_f32<T>(log_mode ? (g - expf(yy) * dot) : yy * (g - dot));
  }
}

// ---------------------------------------------------------------- fused log-softmax + NLL
template <typename T>
__global__ void xent_fwd_kernel(int B, int C, int Tn, const T* x, const long long* target, float* loss_bt,
                                float* lse_out) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= (long long)B * Tn) return;
  const long long b = col / Tn, t = col - b * Tn;
  const T* xp = x + b * (long long)C * Tn + t;
  float m = -INFINITY, s = 0.f;
  for (int c = 0; c < C; ++c) {
    const float v = to_f32<T>(xp[(long long)c * Tn]);
    if (v > m) { s = s * expf(m - v); m = v; }
    s += expf(v - m);
  }
  const float lse = m + logf(s);
  long long tg = target[col];
  tg = tg < 0 ? 0 : (tg >= C ? C - 1 : tg);
  loss_bt[col] = lse - to_f32<T>(xp[tg * Tn]);
  lse_out[col] = lse;
}

template <typename T>
__global__ void xent_bwd_kernel(int B, int C, int Tn, const T* x, const long long* target, const float* lse,
                                const float* gscale, T* dx) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= (long long)B * Tn) return;
  const long long b = col / Tn, t = col - b * Tn;
  const long long base = b * (long long)C * Tn + t;
  const float g = *gscale, l = lse[col];
  const long long tg = target[col];
  for (int c = 0; c < C; ++c) {
    const float pr = expf(to_f32<T>(x[base + (long long)c * Tn]) - l);
    dx[base + (long long)c * Tn] = from_f32<T>((pr - (c == tg ? 1.f : 0.f)) * g);
  }
}

__global__ void sum_stage_kernel(long long n, const float* x, float* partial) {
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    s += x[i];
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
}

// ---------------------------------------------------------------- mean pool (classifier.py:53,102)
template <typename T>
__global__ void avgpool_fwd_kernel(long long rows, int Tn, int To, int pool, const T* x, T* y) {
  const long long n = rows * To;
  const float inv = 1.f / (float)pool;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / To;
    const int to = (int)(i - r * To);
    const T* xp = x + r * Tn + (long long)to * pool;
    float s = 0.f;
    for (int j = 0; j < pool; ++j) s += to_f32<T>(xp[j]);
    y[i] = from_f32<T>(s * inv);
  }
}

// input rows 16-byte aligned: a thread pools V outputs (one 16-byte vector when the pooled rows are aligned too, element
// stores otherwise -- T / pool is rarely a multiple of 8) from P 16-byte vectors of input
template <typename T, int P>
__global__ void __launch_bounds__(256) avgpool_fwd_vec_kernel(long long rows, int Tn, int To, bool out_vec,
                                                              const T* __restrict__ x, T* __restrict__ y) {
  constexpr int V = 16 / (int)sizeof(T);
  const int nv = (To + V - 1) / V;
  const long long n = rows * nv;
  const float inv = 1.f / (float)P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nv;
    const int j = (int)(i - r * nv);
    const T* xr = x + r * Tn + (long long)j * V * P;
    float v[V * P];
    if ((j + 1) * V * P <= Tn) {
#pragma unroll
      for (int q = 0; q < P; ++q) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr) + q);
        const T* e = reinterpret_cast<const T*>(&u);
#pragma unroll
        for (int k = 0; k < V; ++k) v[q * V + k] = to_f32<T>(e[k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < V * P; ++k) v[k] = (j * V * P + k < Tn) ? to_f32<T>(xr[k]) : 0.f;
    }
    uint4 o;
    T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int f = 0; f < V; ++f) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < P; ++k) s += v[f * P + k];
      oe[f] = from_f32<T>(s * inv);
    }
    T* yr = y + r * To + (long long)j * V;
    if (out_vec && (j + 1) * V <= To) {
      *reinterpret_cast<uint4*>(yr) = o;
    } else {
#pragma unroll
      for (int f = 0; f < V; ++f)
        if (j * V + f < To) yr[f] = oe[f];
    }
  }
}

template <typename T>
__global__ void avgpool_bwd_kernel(long long rows, int Tn, int To, int pool, const T* dy, T* dx) {
  const long long n = rows * Tn;
  const float inv = 1.f / (float)pool;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Tn;
    const int t = (int)(i - r * Tn);
    const int to = t / pool;
    dx[i] = from_f32<T>(to < To ? to_f32<T>(dy[r * To + to]) * inv : 0.f);
  }
}

// ---------------------------------------------------------------- LayerNorm (layernorm.py:25-28)
template <typename T>
__global__ void layernorm_fwd_kernel(int B, int C, int Tn, const T* x, const float* gamma, const float* beta,
                                     float eps, T* y, float* stats) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= (long long)B * Tn) return;
  const long long b = col / Tn, t = col - b * Tn;
  const long long base = b * (long long)C * Tn + t;
  float mean = 0.f;
  for (int c = 0; c < C; ++c) mean += to_f32<T>(x[base + (long long)c * Tn]);
  mean /= (float)C;
  float var = 0.f;
  for (int c = 0; c < C; ++c) {
    const float d = to_f32<T>(x[base + (long long)c * Tn]) - mean;
    var += d * d;
  }
  const float sd = sqrtf(var / (float)(C - 1));   // unbiased, as torch.std
  const float r = 1.f / (sd + eps);
  for (int c = 0; c < C; ++c) {
    const float d = to_f32<T>(x[base + (long long)c * Tn]) - mean;
    y[base + (long long)c * Tn] = from_f32<T>(gamma[c] * d * r + beta[c]);
  }
  if (stats) { stats[col * 2] = mean; stats[col * 2 + 1] = r; }
}

template <typename T>
__global__ void layernorm_bwd_kernel(int B, int C, int Tn, const T* x, const float* gamma, const float* stats,
                                     float eps, const T* dy, T* dx) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= (long long)B * Tn) return;
  const long long b = col / Tn, t = col - b * Tn;
  const long long base = b * (long long)C * Tn + t;
  const float mean = stats[col * 2], r = stats[col * 2 + 1];
  const float sd = 1.f / r - eps;
  // y = gamma * xc * r + beta,  r = 1/(sd+eps),  sd = sqrt(sum xc^2 / (C-1))
  float s1 = 0.f, s2 = 0.f;   // sum(g), sum(g * xc)
  for (int c = 0; c < C; ++c) {
    const float g = to_f32<T>(dy[base + (long long)c * Tn]) * gamma[c];
    const float xc = to_f32<T>(x[base + (long long)c * Tn]) - mean;
    s1 += g; s2 += g * xc;
  }
  // d sd = -r^2 * s2 ; d xc_i (via sd) = dsd * xc_i / ((C-1) sd)
  const float k = (sd > 0.f) ? (-r * r * s2 / ((float)(C - 1) * sd)) : 0.f;
  // dx_i = dxc_i - mean_j(dxc_j); sum_j xc_j = 0 so mean_j(dxc_j) = r * s1 / C
  const float mu = r * s1 / (float)C;
  for (int c = 0; c < C; ++c) {
    const float g = to_f32<T>(dy[base + (long long)c * Tn]) * gamma[c];
    const float xc = to_f32<T>(x[base + (long long)c * Tn]) - mean;
    dx[base + (long long)c * Tn] = from_f32<T>(g * r + k * xc - mu);
  }
}

// ---------------------------------------------------------------- positions (raw_ctcnet.py:131-135)
template <typename T>
__global__ void positions_add_kernel(int B, int F, int Tn, int t0, const float* w, const float* bias, T* out) {
  const long long n = (long long)B * F * Tn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % Tn);
    const int f = (int)((i / Tn) % F);
    float v = w[f] * (float)(t + t0) + bias[f];
    v = fminf(1.f, fmaxf(-1.f, v));
    out[i] = from_f32<T>(to_f32<T>(out[i]) + v);
  }
}

// Gradients of the position layer's two parameters (autograd of raw_ctcnet.py:131-135; the reference trains RawCTCNet
// with positions=True in pretrain_tnt.py:121-124): with u = w[f] * (t + t0) + bias[f],
//   dw[f] += sum_{b,t} g[b,f,t] * (t + t0) * [|u| < 1],   db[f] += sum_{b,t} g[b,f,t] * [|u| < 1]     (hardtanh's slope)
// grid (F, time chunks): coalesced along t, one pair of atomics per block.
template <typename T>
__global__ void __launch_bounds__(256) positions_bwd_kernel(int B, int F, int Tn, int t0, int tchunk, const float* w,
                                                            const float* bias, const T* g, float* dw, float* db) {
  const int f = blockIdx.x;
  const int tb = blockIdx.y * tchunk, te = min(Tn, tb + tchunk);
  const float wf = w[f], bf = bias[f];
  float sw = 0.f, sb = 0.f;
  for (int b = 0; b < B; ++b) {
    const T* gp = g + ((long long)b * F + f) * Tn;
    for (int t = tb + threadIdx.x; t < te; t += blockDim.x) {
      const float tt = (float)(t + t0);
      const float u = wf * tt + bf;
      if (u > -1.f && u < 1.f) {
        const float gv = to_f32<T>(gp[t]);
        sw += gv * tt;
        sb += gv;
      }
    }
  }
  __shared__ float red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sw; red[1][threadIdx.x >> 5] = sb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; c += red[1][i]; }
    atomicAdd(dw + f, a);
    atomicAdd(db + f, c);
  }
}

template <typename T>
__global__ void argmax_kernel(int B, int C, int Tn, const T* x, long long* out) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= (long long)B * Tn) return;
  const long long b = col / Tn, t = col - b * Tn;
  const T* xp = x + b * (long long)C * Tn + t;
  float m = to_f32<T>(xp[0]);
  int arg = 0;
  for (int c = 1; c < C; ++c) {
    const float v = to_f32<T>(xp[(long long)c * Tn]);
    if (v > m) { m = v; arg = c; }
  }
  out[col] = arg;
}

// gate backward on NLC bf16: 8 channels (16 B) per thread
// A thread owns one 8-channel group for the whole launch (C/8 threads per frame side by side, 256/(C/8) frames per
// block pass, grid-stride over frames), so the bias gradients -- the column sums of dab -- ride along in registers:
// dbias[0:C] += sum_rows da, dbias[C:2C] += sum_rows ds (fp32, one atomicAdd per channel per block; optional).
__global__ void __launch_bounds__(256)
gate_bwd_nlc_kernel(long long rows, int C, const uint4* dact, const uint4* th, const uint4* sg, uint4* dab,
                    float* dbias, bool th_is_gate) {
  const int g = C / 8, rpb = 256 / g;
  const int cg = threadIdx.x % g, rl = threadIdx.x / g;
  float sa[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, sb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < rpb) {
    for (long long r = blockIdx.x * (long long)rpb + rl; r < rows; r += (long long)gridDim.x * rpb) {
      const long long i = r * g + cg;
      const uint4 d4 = __ldg(dact + i), t4 = __ldg(th + i), s4 = __ldg(sg + i);
      const __nv_bfloat162* dp = reinterpret_cast<const __nv_bfloat162*>(&d4);
      const __nv_bfloat162* tp = reinterpret_cast<const __nv_bfloat162*>(&t4);
      const __nv_bfloat162* sp = reinterpret_cast<const __nv_bfloat162*>(&s4);
      uint4 oa, ob;
      __nv_bfloat162* ap = reinterpret_cast<__nv_bfloat162*>(&oa);
      __nv_bfloat162* bp = reinterpret_cast<__nv_bfloat162*>(&ob);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 d = __bfloat1622float2(dp[k]), s_ = __bfloat1622float2(sp[k]);
        float2 t = __bfloat1622float2(tp[k]);
        if (th_is_gate) {            // the forward kept gate = tanh * sigmoid and sigmoid: tanh = gate / sigmoid (0 / 0 -> 0)
          t.x = s_.x > 0.f ? fminf(1.f, fmaxf(-1.f, __fdividef(t.x, s_.x))) : 0.f;
          t.y = s_.y > 0.f ? fminf(1.f, fmaxf(-1.f, __fdividef(t.y, s_.y))) : 0.f;
        }
        const float a0 = d.x * s_.x * (1.f - t.x * t.x), a1 = d.y * s_.y * (1.f - t.y * t.y);
        const float b0 = d.x * t.x * s_.x * (1.f - s_.x), b1 = d.y * t.y * s_.y * (1.f - s_.y);
        ap[k] = __floats2bfloat162_rn(a0, a1);
        bp[k] = __floats2bfloat162_rn(b0, b1);
        sa[2 * k] += a0; sa[2 * k + 1] += a1;
        sb[2 * k] += b0; sb[2 * k + 1] += b1;
      }
      dab[r * (2 * g) + cg] = oa;
      dab[r * (2 * g) + g + cg] = ob;
    }
  }
  if (dbias == nullptr) return;
  __shared__ float red[256][17];
#pragma unroll
  for (int k = 0; k < 8; ++k) { red[threadIdx.x][k] = sa[k]; red[threadIdx.x][8 + k] = sb[k]; }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    const int half = c >= C, cc = c - half * C;
    const int gq = cc >> 3, k = (cc & 7) + 8 * half;
    float tsum = 0.f;
    for (int q = 0; q < rpb; ++q) tsum += red[q * g + gq][k];
    atomicAdd(&dbias[c], tsum);
  }
}

// column sums of an NLC bf16 tensor [rows][C], C % 8 == 0: a thread owns 8 channels (one 16-byte load per row), the
// C/8 threads of a row sit side by side (a warp reads whole 512-byte rows), 256 / (C/8) rows per block pass,
// grid-stride over rows; partials meet in shared memory, one atomicAdd per channel per block.
__global__ void __launch_bounds__(256) colsum_nlc_kernel(long long rows, int C, const uint4* x, float* out) {
  const int g = C >> 3;                       // 16-byte groups per row
  const int rpb = 256 / g;                    // rows per block pass (g <= 256)
  const int cg = threadIdx.x % g, rl = threadIdx.x / g;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < rpb) {
    for (long long r = blockIdx.x * (long long)rpb + rl; r < rows; r += (long long)gridDim.x * rpb) {
      const uint4 v = __ldg(x + r * g + cg);
      const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(p[k]);
        acc[2 * k] += f.x;
        acc[2 * k + 1] += f.y;
      }
    }
  }
  __shared__ float red[256][9];
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x][k] = acc[k];
  __syncthreads();
  // thread -> one channel: sum over the rpb row lanes
  for (int c = threadIdx.x; c < C; c += 256) {
    const int gq = c >> 3, k = c & 7;
    float t = 0.f;
    for (int q = 0; q < rpb; ++q) t += red[q * g + gq][k];
    atomicAdd(&out[c], t);
  }
}

// Backward of the RawCTCNet featuriser's first layer (raw_ctcnet.py:57-59: Conv1d(1, F, fk, padding=fk-1) + LeakyReLU)
// in ONE pass over the NLC bf16 tensors: with dpre = df * LeakyReLU'(f),
//   dw[f, j] += sum_{b,t'} dpre[b,t',f] * seq[b, t'+j-(fk-1)]      db[f] += sum dpre[b,t',f]
//   dseq[b, t] += sum_j sum_f w[f, j] * dpre[b, t+(fk-1)-j, f]      (fp32, atomics; fk values per frame)
// Thread = 8 channels of a frame (one 16-byte load per tensor), the F/8 threads of a frame sit side by side (a power
// of two <= 32: the per-frame dot products over F are warp-shuffle reductions), a block walks FB_FRAMES frames of one
// read with its fk x 8 weight-gradient accumulators in registers.
constexpr int FB_FRAMES = 512;
template <typename TS, int FK>
__global__ void __launch_bounds__(256)
featurize_bwd_nlc_kernel(int Tn, int F, const uint4* df, const uint4* fact, const TS* seq, const float* w, float* dw,
                         float* db, float* dseq) {
  __shared__ float xs[FB_FRAMES + FK];
  __shared__ float red[256][FK * 8 + 9];
  const int b = blockIdx.y, t0 = blockIdx.x * FB_FRAMES, To = Tn + FK - 1;
  const int groups = F / 8, rows = 256 / groups;
  const int gq = threadIdx.x % groups, rl = threadIdx.x / groups, f0 = gq * 8;
  for (int i = threadIdx.x; i < FB_FRAMES + FK - 1; i += 256) {
    const int t = t0 + i - (FK - 1);
    xs[i] = (t >= 0 && t < Tn) ? to_f32<TS>(seq[(long long)b * Tn + t]) : 0.f;
  }
  float wr[FK * 8], acc[FK * 8], accb[8];
#pragma unroll
  for (int j = 0; j < FK; ++j)
#pragma unroll
    for (int k = 0; k < 8; ++k) { wr[j * 8 + k] = w[(f0 + k) * FK + j]; acc[j * 8 + k] = 0.f; }
#pragma unroll
  for (int k = 0; k < 8; ++k) accb[k] = 0.f;
  __syncthreads();
  for (int tl = rl; tl < FB_FRAMES; tl += rows) {
    const int t = t0 + tl;
    const bool live = t < To;                  // no early exit: every lane takes part in the shuffles below
    const long long o = ((long long)b * To + (live ? t : 0)) * groups + gq;
    uint4 d4 = make_uint4(0, 0, 0, 0), f4 = d4;
    if (live) { d4 = __ldg(df + o); f4 = __ldg(fact + o); }
    const __nv_bfloat162* dp = reinterpret_cast<const __nv_bfloat162*>(&d4);
    const __nv_bfloat162* fp = reinterpret_cast<const __nv_bfloat162*>(&f4);
    float d[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 dv = __bfloat1622float2(dp[k]), fv = __bfloat1622float2(fp[k]);
      d[2 * k] = fv.x > 0.f ? dv.x : 0.01f * dv.x;
      d[2 * k + 1] = fv.y > 0.f ? dv.y : 0.01f * dv.y;
    }
    float pj[FK];
#pragma unroll
    for (int j = 0; j < FK; ++j) {
      const float xv = xs[tl + j];
      float p = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[j * 8 + k] = fmaf(d[k], xv, acc[j * 8 + k]);
        p = fmaf(wr[j * 8 + k], d[k], p);
      }
      pj[j] = p;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) accb[k] += d[k];
    if (dseq) {
#pragma unroll
      for (int j = 0; j < FK; ++j) {
        float p = pj[j];
        for (int o2 = groups >> 1; o2 > 0; o2 >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o2);
        const int ts = t + j - (FK - 1);
        if (live && gq == 0 && ts >= 0 && ts < Tn) atomicAdd(&dseq[(long long)b * Tn + ts], p);
      }
    }
  }
  // block reduction over the row lanes, then one atomic per (channel, tap) per block
#pragma unroll
  for (int i = 0; i < FK * 8; ++i) red[threadIdx.x][i] = acc[i];
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x][FK * 8 + k] = accb[k];
  __syncthreads();
  for (int i = threadIdx.x; i < F * (FK + 1); i += 256) {
    const int c = i / (FK + 1), j = i - c * (FK + 1);          // j == FK: bias
    const int g2 = c >> 3, k = c & 7;
    float tsum = 0.f;
    for (int q = 0; q < rows; ++q) tsum += red[q * groups + g2][j * 8 + k];
    if (j < FK) atomicAdd(&dw[c * FK + j], tsum);
    else atomicAdd(&db[c], tsum);
  }
}

// MultiplicativeUnit gate (block.py:213-220): pre = the four convolutions' outputs stacked on the channel axis
// [B, 4C, T] = (gate1 ; gate2 ; gate3 ; update), h [B, C, T]:
//   g_i = sigmoid(pre_i), u = tanh(pre_4), out = g1 * tanh(g2 * h + g3 * u)
template <typename T>
__global__ void mu_gate_fwd_kernel(int B, int C, int Tn, const T* pre, const T* h, T* out) {
  const long long n = (long long)B * C * Tn, plane = (long long)C * Tn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / plane, r = i - b * plane;
    const T* p = pre + b * 4 * plane + r;
    const float g1 = sigmoid_precise(to_f32<T>(p[0])), g2 = sigmoid_precise(to_f32<T>(p[plane]));
    const float g3 = sigmoid_precise(to_f32<T>(p[2 * plane])), u = tanhf(to_f32<T>(p[3 * plane]));
    out[i] = from_f32<T>(g1 * tanhf(g2 * to_f32<T>(h[i]) + g3 * u));
  }
}

template <typename T>
__global__ void mu_gate_bwd_kernel(int B, int C, int Tn, const T* pre, const T* h, const T* dout, T* dpre, T* dh) {
  const long long n = (long long)B * C * Tn, plane = (long long)C * Tn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / plane, r = i - b * plane;
    const T* p = pre + b * 4 * plane + r;
    T* q = dpre + b * 4 * plane + r;
    const float g1 = sigmoid_precise(to_f32<T>(p[0])), g2 = sigmoid_precise(to_f32<T>(p[plane]));
    const float g3 = sigmoid_precise(to_f32<T>(p[2 * plane])), u = tanhf(to_f32<T>(p[3 * plane]));
    const float hv = to_f32<T>(h[i]), d = to_f32<T>(dout[i]);
    const float ts = tanhf(g2 * hv + g3 * u);
    const float ds = d * g1 * (1.f - ts * ts);
    q[0] = from_f32<T>(d * ts * g1 * (1.f - g1));
    q[plane] = from_f32<T>(ds * hv * g2 * (1.f - g2));
    q[2 * plane] = from_f32<T>(ds * u * g3 * (1.f - g3));
    q[3 * plane] = from_f32<T>(ds * g3 * (1.f - u * u));
    dh[i] = from_f32<T>(ds * g2);
  }
}

static inline int grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  const long long cap = 148ll * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace wnb

using namespace wnb;
typedef __nv_bfloat16 bf16;

#define DISPATCH(dtype, NAME, ...)                                             \
  do {                                                                         \
    if ((dtype) == WNB200_F32) { using T = float; __VA_ARGS__; }               \
    else if ((dtype) == WNB200_BF16) { using T = bf16; __VA_ARGS__; }          \
    else { set_error(NAME ": bad dtype %d", (int)(dtype)); return 1; }         \
  } while (0)

extern "C" int wnb200_gate_bwd(int dtype, int B, int C, int T_, const void* dact, const void* th, const void* sg,
                               void* d_ab, void* stream) {
  WNB_CHECK_ARG(dact && th && sg && d_ab, "gate_bwd: null pointer");
  const long long n = (long long)B * C * T_;
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(dtype, "gate_bwd", (gate_bwd_kernel<T><<<grid_for(n), 256, 0, st>>>(B, C, T_, (const T*)dact,
                                                                              (const T*)th, (const T*)sg, (T*)d_ab)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_gate_fwd(int dtype, int64_t n, const void* a, const void* b, void* out, void* th, void* sg,
                               void* stream) {
  if (n == 0) return 0;
  WNB_CHECK_ARG(a && b && out, "gate_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(dtype, "gate_fwd", (gate_fwd_kernel<T><<<grid_for(n), 256, 0, st>>>(n, (const T*)a, (const T*)b, (T*)out,
                                                                              (T*)th, (T*)sg)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_layernorm_bwd_params(int dtype, int B, int C, int T_, const void* x, const float* stats,
                                           const void* dy, float* dgamma, float* dbeta, void* stream) {
  if (B == 0 || C == 0 || T_ == 0) return 0;
  WNB_CHECK_ARG(x && stats && dy && dgamma && dbeta, "layernorm_bwd_params: null pointer");
  WNB_CHECK_ARG(B <= 65535, "layernorm_bwd_params: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(C, B);
  DISPATCH(dtype, "layernorm_bwd_params", (layernorm_bwd_params_kernel<T><<<grid, 256, 0, st>>>(
                                               C, T_, (const T*)x, stats, (const T*)dy, dgamma, dbeta)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_leaky_bwd(int dtype, int64_t n, const void* dy, const void* ref, void* dx, void* stream) {
  WNB_CHECK_ARG(dy && ref && dx, "leaky_bwd: null pointer");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool aligned = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(ref) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
  const int V = dtype == WNB200_F32 ? 4 : 8;
  const long long nv = aligned ? n / V : 0, done = nv * V;
  if (nv > 0)
    DISPATCH(dtype, "leaky_bwd", (leaky_bwd_vec_kernel<T><<<grid_for(nv), 256, 0, st>>>(nv, (const uint4*)dy, (const uint4*)ref, (uint4*)dx)));
  if (done < n)
    DISPATCH(dtype, "leaky_bwd", (leaky_bwd_kernel<T><<<grid_for(n - done), 256, 0, st>>>(
                                      n - done, (const T*)dy + done, (const T*)ref + done, (T*)dx + done)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_softmax_fwd_col(int dtype, int B, int C, int T_, const void* x, void* y, int log_mode,
                                  void* stream) {
  WNB_CHECK_ARG(x && y && C >= 1, "softmax_fwd: bad args");
  const long long cols = (long long)B * T_;
  if (cols == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(dtype, "softmax_fwd", (softmax_fwd_kernel<T><<<(unsigned)ceil_div64(cols, 128), 128, 0, st>>>(
                                      B, C, T_, (const T*)x, (T*)y, log_mode)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_softmax_bwd_col(int dtype, int B, int C, int T_, const void* y, const void* dy, void* dx,
                                  int log_mode, void* stream) {
  WNB_CHECK_ARG(y && dy && dx, "softmax_bwd: null pointer");
  const long long cols = (long long)B * T_;
  if (cols == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(dtype, "softmax_bwd", (softmax_bwd_kernel<T><<<(unsigned)ceil_div64(cols, 128), 128, 0, st>>>(
                                      B, C, T_, (const T*)y, (const T*)dy, (T*)dx, log_mode)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_xent_fwd_col(int dtype, int B, int C, int T_, const void* logits, const int64_t* target,
                               float* loss_bt, float* lse, void* stream) {
  WNB_CHECK_ARG(logits && target && loss_bt && lse, "xent_fwd: null pointer");
  const long long cols = (long long)B * T_;
  if (cols == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(dtype, "xent_fwd", (xent_fwd_kernel<T><<<(unsigned)ceil_div64(cols, 128), 128, 0, st>>>(
                                   B, C, T_, (const T*)logits, (const long long*)target, loss_bt, lse)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_xent_bwd_col(int dtype, int B, int C, int T_, const void* logits, const int64_t* target,
                               const float* lse, const float* gscale, void* dlogits, void* stream) {
  WNB_CHECK_ARG(logits && target && lse && gscale && dlogits, "xent_bwd: null pointer");
  const long long cols = (long long)B * T_;
  if (cols == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(dtype, "xent_bwd", (xent_bwd_kernel<T><<<(unsigned)ceil_div64(cols, 128), 128, 0, st>>>(
                                   B, C, T_, (const T*)logits, (const long long*)target, lse, gscale, (T*)dlogits)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_sum_f32(int64_t n, const float* x, float* out, float* scratch, void* stream) {
  WNB_CHECK_ARG(x && out && scratch, "sum_f32: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int g = grid_for(n, 256);
  if (g > 1024) g = 1024;
  sum_stage_kernel<<<g, 256, 0, st>>>(n, x, scratch);
  sum_stage_kernel<<<1, 256, 0, st>>>(g, scratch, out);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_avgpool_fwd(int dtype, int B, int C, int T_, int pool, const void* x, void* y, void* stream) {
  WNB_CHECK_ARG(x && y && pool >= 1, "avgpool_fwd: bad args");
  const int To = T_ / pool;
  const long long n = (long long)B * C * To;
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  {
    const int V = dtype == WNB200_F32 ? 4 : 8;
    if (pool <= 4 && T_ % V == 0 && To > 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
      const long long nv = (long long)B * C * ((To + V - 1) / V);
      const bool out_vec = To % V == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
#define AP_VEC(TT, PP) avgpool_fwd_vec_kernel<TT, PP><<<grid_for(nv), 256, 0, st>>>((long long)B * C, T_, To, out_vec, (const TT*)x, (TT*)y)
#define AP_POOL(TT) switch (pool) { case 1: AP_VEC(TT, 1); break; case 2: AP_VEC(TT, 2); break; case 3: AP_VEC(TT, 3); break; default: AP_VEC(TT, 4); break; }
      if (dtype == WNB200_F32) { AP_POOL(float) } else if (dtype == WNB200_BF16) { AP_POOL(bf16) } else { set_error("avgpool_fwd: bad dtype %d", dtype); return 1; }
#undef AP_POOL
#undef AP_VEC
      WNB_LAUNCH_OK();
      return 0;
    }
  }
  DISPATCH(dtype, "avgpool_fwd", (avgpool_fwd_kernel<T><<<grid_for(n), 256, 0, st>>>((long long)B * C, T_, To, pool,
                                                                                    (const T*)x, (T*)y)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_avgpool_bwd(int dtype, int B, int C, int T_, int pool, const void* dy, void* dx, void* stream) {
  WNB_CHECK_ARG(dy && dx && pool >= 1, "avgpool_bwd: bad args");
  const int To = T_ / pool;
  const long long n = (long long)B * C * T_;
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(dtype, "avgpool_bwd", (avgpool_bwd_kernel<T><<<grid_for(n), 256, 0, st>>>((long long)B * C, T_, To, pool,
                                                                                    (const T*)dy, (T*)dx)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_layernorm_fwd_col(int dtype, int B, int C, int T_, const void* x, const float* gamma,
                                    const float* beta, float eps, void* y, float* stats, void* stream) {
  WNB_CHECK_ARG(x && y && gamma && beta && C >= 2, "layernorm_fwd: bad args");
  const long long cols = (long long)B * T_;
  if (cols == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(dtype, "layernorm_fwd", (layernorm_fwd_kernel<T><<<(unsigned)ceil_div64(cols, 128), 128, 0, st>>>(
                                        B, C, T_, (const T*)x, gamma, beta, eps, (T*)y, stats)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_layernorm_bwd_col(int dtype, int B, int C, int T_, const void* x, const float* gamma,
                                    const float* stats, float eps, const void* dy, void* dx, void* stream) {
  WNB_CHECK_ARG(x && gamma && stats && dy && dx, "layernorm_bwd: null pointer");
  const long long cols = (long long)B * T_;
  if (cols == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(dtype, "layernorm_bwd", (layernorm_bwd_kernel<T><<<(unsigned)ceil_div64(cols, 128), 128, 0, st>>>(
                                        B, C, T_, (const T*)x, gamma, stats, eps, (const T*)dy, (T*)dx)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_positions_add(int dtype, int B, int F, int T_, int t0, const float* w, const float* bias,
                                    void* out, void* stream) {
  WNB_CHECK_ARG(w && bias && out, "positions_add: null pointer");
  const long long n = (long long)B * F * T_;
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(dtype, "positions_add", (positions_add_kernel<T><<<grid_for(n), 256, 0, st>>>(B, F, T_, t0, w, bias, (T*)out)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_positions_bwd(int dtype, int B, int F, int T_, int t0, const float* w, const float* bias,
                                    const void* g, float* dw, float* db, void* stream) {
  if (B == 0 || F == 0 || T_ == 0) return 0;
  WNB_CHECK_ARG(w && bias && g && dw && db, "positions_bwd: null pointer");
  WNB_CHECK_ARG(F <= 65535 * 32, "positions_bwd: F too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int tchunk = 4096;
  dim3 grid(F, ceil_div(T_, tchunk));
  DISPATCH(dtype, "positions_bwd", (positions_bwd_kernel<T><<<grid, 256, 0, st>>>(B, F, T_, t0, tchunk, w, bias,
                                                                                  (const T*)g, dw, db)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_argmax_channels_col(int dtype, int B, int C, int T_, const void* x, int64_t* out, void* stream) {
  WNB_CHECK_ARG(x && out && C >= 1, "argmax_channels: bad args");
  const long long cols = (long long)B * T_;
  if (cols == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(dtype, "argmax_channels", (argmax_kernel<T><<<(unsigned)ceil_div64(cols, 128), 128, 0, st>>>(
                                          B, C, T_, (const T*)x, (long long*)out)));
  WNB_LAUNCH_OK();
  return 0;
}

static int gate_bwd_nlc_launch(int64_t rows, int C, const void* dact, const void* th, const void* sg, void* dab,
                               float* dbias, bool th_is_gate, void* stream) {
  WNB_CHECK_ARG(C % 8 == 0 && C <= 2048, "gate_bwd_nlc: C=%d must be a multiple of 8, <= 2048", C);
  if (rows == 0) return 0;
  WNB_CHECK_ARG(dact && th && sg && dab, "gate_bwd_nlc: null pointer");
  const int rpb = 256 / (C / 8);
  long long grid = (rows + rpb - 1) / rpb;
  if (grid > 148 * 8) grid = 148 * 8;
  gate_bwd_nlc_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      rows, C, (const uint4*)dact, (const uint4*)th, (const uint4*)sg, (uint4*)dab, dbias, th_is_gate);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_gate_bwd_nlc(int64_t rows, int C, const void* dact, const void* th, const void* sg, void* dab,
                                   float* dbias, void* stream) {
  return gate_bwd_nlc_launch(rows, C, dact, th, sg, dab, dbias, false, stream);
}

extern "C" int wnb200_gate_bwd_nlc_from_gate(int64_t rows, int C, const void* dact, const void* gate, const void* sg,
                                             void* dab, float* dbias, void* stream) {
  return gate_bwd_nlc_launch(rows, C, dact, gate, sg, dab, dbias, true, stream);
}

extern "C" int wnb200_colsum_nlc(int64_t rows, int C, const void* x, float* out, void* stream) {
  if (rows == 0 || C == 0) return 0;
  WNB_CHECK_ARG(x && out, "colsum_nlc: null pointer");
  WNB_CHECK_ARG(C % 8 == 0 && C <= 2048, "colsum_nlc: C=%d must be a multiple of 8, <= 2048", C);
  const int rpb = 256 / (C / 8);
  long long grid = (rows + rpb - 1) / rpb;
  if (grid > 148 * 8) grid = 148 * 8;
  colsum_nlc_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(rows, C, (const uint4*)x, out);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_featurize_bwd_nlc(int seq_dtype, int B, int T_, int F, int fk, const void* df, const void* fact,
                                        const void* seq, const float* w, float* dw, float* db, float* dseq,
                                        void* stream) {
  WNB_CHECK_ARG(F >= 8 && F <= 256 && (F & (F - 1)) == 0, "featurize_bwd_nlc: F=%d must be a power of two in 8..256", F);
  WNB_CHECK_ARG(fk >= 1 && fk <= 4, "featurize_bwd_nlc: fk=%d not in 1..4", fk);
  if (B == 0 || T_ == 0) return 0;
  WNB_CHECK_ARG(df && fact && seq && w && dw && db, "featurize_bwd_nlc: null pointer");
  WNB_CHECK_ARG(B <= 65535, "featurize_bwd_nlc: batch too large");
  const int To = T_ + fk - 1;
  dim3 grid((To + FB_FRAMES - 1) / FB_FRAMES, B);
  cudaStream_t st = (cudaStream_t)stream;
#define FB_LAUNCH(TS, FK)                                                                                         \
  featurize_bwd_nlc_kernel<TS, FK><<<grid, 256, 0, st>>>(T_, F, (const uint4*)df, (const uint4*)fact, (const TS*)seq, \
                                                         w, dw, db, dseq)
#define FB_DISPATCH(TS)                  \
  switch (fk) {                          \
    case 1: FB_LAUNCH(TS, 1); break;     \
    case 2: FB_LAUNCH(TS, 2); break;     \
    case 3: FB_LAUNCH(TS, 3); break;     \
    default: FB_LAUNCH(TS, 4); break;    \
  }
  if (seq_dtype == WNB200_F32) { FB_DISPATCH(float) }
  else if (seq_dtype == WNB200_BF16) { FB_DISPATCH(bf16) }
  else { set_error("featurize_bwd_nlc: bad dtype %d", seq_dtype); return 1; }
#undef FB_DISPATCH
#undef FB_LAUNCH
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_mu_gate_fwd(int dtype, int B, int C, int T_, const void* pre, const void* h, void* out,
                                  void* stream) {
  const long long n = (long long)B * C * T_;
  if (n == 0) return 0;
  WNB_CHECK_ARG(pre && h && out, "mu_gate_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(dtype, "mu_gate_fwd", (mu_gate_fwd_kernel<T><<<grid_for(n), 256, 0, st>>>(B, C, T_, (const T*)pre,
                                                                                    (const T*)h, (T*)out)));
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_mu_gate_bwd(int dtype, int B, int C, int T_, const void* pre, const void* h, const void* dout,
                                  void* dpre, void* dh, void* stream) {
  const long long n = (long long)B * C * T_;
  if (n == 0) return 0;
  WNB_CHECK_ARG(pre && h && dout && dpre && dh, "mu_gate_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(dtype, "mu_gate_bwd", (mu_gate_bwd_kernel<T><<<grid_for(n), 256, 0, st>>>(
                                      B, C, T_, (const T*)pre, (const T*)h, (const T*)dout, (T*)dpre, (T*)dh)));
  WNB_LAUNCH_OK();
  return 0;
}
