// Channel-reduction kernels on NCL tensors -- channel softmax / log-softmax (wavenet.py:108-109), fused
// log-softmax + NLL (legacy_code/train.py:36-39), LayerNorm (layernorm.py:25-28), per-frame argmax
// (legacy_code/train.py:36) -- as TILED bandwidth kernels.
//
// A CTA stages one read's [C channels x TT frames] tile in shared memory with 16-byte coalesced loads (all of them
// in flight at once), reduces every frame's column out of shared memory (TT columns x 256/TT partial reducers),
// and streams the result back with 16-byte coalesced stores: each input byte is read from HBM once and each output
// byte written once.  (The first version gave every (read, frame) column to one thread that walked the channels with
// stride T, two or three times: 17-21 % of the copy bandwidth.)  TT is chosen so that the tile(s) fit in 64 KB; a
// channel count too large for an 8-frame tile falls back to the per-column kernels in elementwise.cu.
#include <float.h>

#include <stdlib.h>

#include "common.cuh"

namespace wnb {

constexpr int CT_THREADS = 256;

template <typename T>
struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
};

// exp for the softmax family: exact expf for fp32 storage (parity 1e-5), ex2.approx for bf16 storage (4e-3 output ulp)
template <typename T>
__device__ __forceinline__ float exp_t(float x);
template <>
__device__ __forceinline__ float exp_t<float>(float x) { return expf(x); }
template <>
__device__ __forceinline__ float exp_t<__nv_bfloat16>(float x) { return __expf(x); }

template <typename T>
__device__ __forceinline__ float2 load2(const T* p);
template <>
__device__ __forceinline__ float2 load2<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <>
__device__ __forceinline__ float2 load2<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}

// cooperative load of x[b, 0:C, t0:t0+TT] -> s[c * stride + tt]  (frames past Tn read as zero); four 16-byte loads in
// flight per thread
template <typename T>
__device__ __forceinline__ void tile_load(T* s, int stride, const T* xb, int C, int Tn, int t0, int TT, bool vec) {
  constexpr int V = Vec16<T>::N;
  if (vec) {
    const int nv = TT / V, total = C * nv;
    for (int i0 = threadIdx.x; i0 < total; i0 += 4 * CT_THREADS) {
      uint4 val[4];
      int off[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * CT_THREADS;
        val[u] = make_uint4(0, 0, 0, 0);
        off[u] = -1;
        if (i < total) {
          const int c = i / nv, v = i - c * nv, t = t0 + v * V;
          off[u] = c * stride + v * V;
          if (t + V <= Tn) val[u] = __ldg(reinterpret_cast<const uint4*>(xb + (long long)c * Tn + t));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (off[u] >= 0) *reinterpret_cast<uint4*>(s + off[u]) = val[u];
    }
  } else {
    // rows that are not 16-byte aligned (odd T): element loads, lane <-> consecutive frames (coalesced), TT is a power
    // of two, four loads in flight per thread
    const int sh = 31 - __clz(TT), total = C * TT;
    for (int i0 = threadIdx.x; i0 < total; i0 += 4 * CT_THREADS) {
      T val[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * CT_THREADS, c = i >> sh, tt = i & (TT - 1);
        val[u] = (i < total && t0 + tt < Tn) ? xb[(long long)c * Tn + t0 + tt] : from_f32<T>(0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * CT_THREADS;
        if (i < total) s[(i >> sh) * stride + (i & (TT - 1))] = val[u];
      }
    }
  }
}

// cooperative store: y[b, c, t0 + tt] = f(c, tt), V consecutive frames per 16-byte store
template <typename T, typename F>
__device__ __forceinline__ void tile_store(T* yb, int C, int Tn, int t0, int TT, bool vec, F f) {
  constexpr int V = Vec16<T>::N;
  if (vec) {
    const int nv = TT / V;
    for (int i = threadIdx.x; i < C * nv; i += CT_THREADS) {
      const int c = i / nv, v = i - c * nv, t = t0 + v * V;
      if (t + V > Tn) continue;                 // Tn is a multiple of V on this path
      uint4 val;
      T* o = reinterpret_cast<T*>(&val);
#pragma unroll
      for (int k = 0; k < V; ++k) o[k] = from_f32<T>(f(c, v * V + k));
      *reinterpret_cast<uint4*>(yb + (long long)c * Tn + t) = val;
    }
  } else {
    const int sh = 31 - __clz(TT);
    for (int i = threadIdx.x; i < C * TT; i += CT_THREADS) {
      const int c = i >> sh, tt = i & (TT - 1);
      if (t0 + tt < Tn) yb[(long long)c * Tn + t0 + tt] = from_f32<T>(f(c, tt));
    }
  }
}

struct TileGeom {
  int TT, stride, nparts, tiles_per_read;      // nparts = 2 * CT_THREADS / TT partial reducers per frame PAIR
};

// thread -> (frame pair `pr` = frames 2pr, 2pr+1 ; part): every reducer walks channels part, part+nparts, ... with one
// 4- or 8-byte shared load per channel for its two frames
#define CT_PROLOGUE(NTILES)                                                                     \
  extern __shared__ __align__(16) unsigned char ct_smem[];                                      \
  const int b = blockIdx.x / g.tiles_per_read;                                                  \
  const int t0 = (blockIdx.x - b * g.tiles_per_read) * g.TT;                                    \
  const int TT = g.TT, stride = g.stride, nparts = g.nparts;                                    \
  T* s0 = reinterpret_cast<T*>(ct_smem);                                                        \
  T* s1 = s0 + (NTILES > 1 ? (size_t)C * stride : 0);                                           \
  float* red = reinterpret_cast<float*>(ct_smem + (size_t)NTILES * C * stride * sizeof(T));     \
  const int pr = threadIdx.x % (TT / 2), part = threadIdx.x / (TT / 2), f0 = 2 * pr;            \
  const int tt = threadIdx.x;                  /* finalising thread of frame tt (tt < TT) */    \
  const long long rb = (long long)b * C * Tn;                                                   \
  (void)s1; (void)part; (void)nparts; (void)tt; (void)f0

// red layout: K partial arrays [nparts][TT], then the per-frame finals
template <typename T>
__global__ void __launch_bounds__(CT_THREADS)
softmax_fwd_tile(int C, int Tn, TileGeom g, bool vec, const T* x, T* y, int log_mode) {
  CT_PROLOGUE(1);
  tile_load(s0, stride, x + rb, C, Tn, t0, TT, vec);
  __syncthreads();
  float* fin = red + nparts * TT;              // [max | lse or 1/sum]
  float m0 = -INFINITY, m1 = -INFINITY;
  for (int c = part; c < C; c += nparts) {
    const float2 v = load2<T>(s0 + c * stride + f0);
    m0 = fmaxf(m0, v.x);
    m1 = fmaxf(m1, v.y);
  }
  red[part * TT + f0] = m0;
  red[part * TT + f0 + 1] = m1;
  __syncthreads();
  if (tt < TT) {
    float M = -INFINITY;
    for (int p = 0; p < nparts; ++p) M = fmaxf(M, red[p * TT + tt]);
    fin[tt] = M;
  }
  __syncthreads();
  m0 = fin[f0];
  m1 = fin[f0 + 1];
  float a0 = 0.f, a1 = 0.f;
  for (int c = part; c < C; c += nparts) {
    const float2 v = load2<T>(s0 + c * stride + f0);
    a0 += exp_t<T>(v.x - m0);
    a1 += exp_t<T>(v.y - m1);
  }
  red[part * TT + f0] = a0;
  red[part * TT + f0 + 1] = a1;
  __syncthreads();
  if (tt < TT) {
    float S = 0.f;
    for (int p = 0; p < nparts; ++p) S += red[p * TT + tt];
    fin[TT + tt] = log_mode ? fin[tt] + logf(S) : 1.f / S;
  }
  __syncthreads();
  tile_store(y + rb, C, Tn, t0, TT, vec, [&](int c, int j) {
    const float v = to_f32<T>(s0[c * stride + j]);
    return log_mode ? v - fin[TT + j] : exp_t<T>(v - fin[j]) * fin[TT + j];
  });
}

template <typename T>
__global__ void __launch_bounds__(CT_THREADS)
softmax_bwd_tile(int C, int Tn, TileGeom g, bool vec, const T* y, const T* dy, T* dx, int log_mode) {
  CT_PROLOGUE(2);
  tile_load(s0, stride, y + rb, C, Tn, t0, TT, vec);
  tile_load(s1, stride, dy + rb, C, Tn, t0, TT, vec);
  __syncthreads();
  float d0 = 0.f, d1 = 0.f;
  for (int c = part; c < C; c += nparts) {
    const float2 gq = load2<T>(s1 + c * stride + f0);
    if (log_mode) { d0 += gq.x; d1 += gq.y; }
    else {
      const float2 yy = load2<T>(s0 + c * stride + f0);
      d0 += gq.x * yy.x;
      d1 += gq.y * yy.y;
    }
  }
  float* fin = red + nparts * TT;
  red[part * TT + f0] = d0;
  red[part * TT + f0 + 1] = d1;
  __syncthreads();
  if (tt < TT) {
    float d = 0.f;
    for (int p = 0; p < nparts; ++p) d += red[p * TT + tt];
    fin[tt] = d;
  }
  __syncthreads();
  tile_store(dx + rb, C, Tn, t0, TT, vec, [&](int c, int j) {
    const float gq = to_f32<T>(s1[c * stride + j]), yy = to_f32<T>(s0[c * stride + j]);
    return log_mode ? gq - exp_t<T>(yy) * fin[j] : yy * (gq - fin[j]);
  });
}

template <typename T>
__global__ void __launch_bounds__(CT_THREADS)
xent_fwd_tile(int C, int Tn, TileGeom g, bool vec, const T* x, const long long* target, float* loss_bt,
              float* lse_out) {
  CT_PROLOGUE(1);
  tile_load(s0, stride, x + rb, C, Tn, t0, TT, vec);
  __syncthreads();
  float* fin = red + nparts * TT;
  float m0 = -INFINITY, m1 = -INFINITY;
  for (int c = part; c < C; c += nparts) {
    const float2 v = load2<T>(s0 + c * stride + f0);
    m0 = fmaxf(m0, v.x);
    m1 = fmaxf(m1, v.y);
  }
  red[part * TT + f0] = m0;
  red[part * TT + f0 + 1] = m1;
  __syncthreads();
  if (tt < TT) {
    float M = -INFINITY;
    for (int p = 0; p < nparts; ++p) M = fmaxf(M, red[p * TT + tt]);
    fin[tt] = M;
  }
  __syncthreads();
  m0 = fin[f0];
  m1 = fin[f0 + 1];
  float a0 = 0.f, a1 = 0.f;
  for (int c = part; c < C; c += nparts) {
    const float2 v = load2<T>(s0 + c * stride + f0);
    a0 += exp_t<T>(v.x - m0);
    a1 += exp_t<T>(v.y - m1);
  }
  red[part * TT + f0] = a0;
  red[part * TT + f0 + 1] = a1;
  __syncthreads();
  if (tt < TT && t0 + tt < Tn) {
    float S = 0.f;
    for (int p = 0; p < nparts; ++p) S += red[p * TT + tt];
    const float lse = fin[tt] + logf(S);
    const long long col = (long long)b * Tn + t0 + tt;
    long long tg = target[col];
    tg = tg < 0 ? 0 : (tg >= C ? C - 1 : tg);
    loss_bt[col] = lse - to_f32<T>(s0[(int)tg * stride + tt]);
    lse_out[col] = lse;
  }
}

template <typename T>
__global__ void __launch_bounds__(CT_THREADS)
xent_bwd_tile(int C, int Tn, TileGeom g, bool vec, const T* x, const long long* target, const float* lse,
              const float* gscale, T* dx) {
  CT_PROLOGUE(1);
  tile_load(s0, stride, x + rb, C, Tn, t0, TT, vec);
  int* tgs = reinterpret_cast<int*>(red + TT);
  if (tt < TT) {
    const bool ok = t0 + tt < Tn;
    const long long col = (long long)b * Tn + t0 + tt;
    red[tt] = ok ? lse[col] : 0.f;
    // out-of-range targets are clamped exactly as the forward clamps them (the two must describe the same loss)
    long long tgc = ok ? target[col] : -1;
    if (ok) tgc = tgc < 0 ? 0 : (tgc >= C ? C - 1 : tgc);
    tgs[tt] = (int)tgc;
  }
  __syncthreads();
  const float gs = *gscale;
  tile_store(dx + rb, C, Tn, t0, TT, vec, [&](int c, int j) {
    const float p = exp_t<T>(to_f32<T>(s0[c * stride + j]) - red[j]);
    return (p - (c == tgs[j] ? 1.f : 0.f)) * gs;
  });
}

template <typename T>
__global__ void __launch_bounds__(CT_THREADS)
layernorm_fwd_tile(int C, int Tn, TileGeom g, bool vec, const T* x, const float* gamma, const float* beta, float eps,
                   T* y, float* stats) {
  CT_PROLOGUE(1);
  tile_load(s0, stride, x + rb, C, Tn, t0, TT, vec);
  __syncthreads();
  float* fin = red + nparts * TT;              // [mean | r]
  float a0 = 0.f, a1 = 0.f;
  for (int c = part; c < C; c += nparts) {
    const float2 v = load2<T>(s0 + c * stride + f0);
    a0 += v.x;
    a1 += v.y;
  }
  red[part * TT + f0] = a0;
  red[part * TT + f0 + 1] = a1;
  __syncthreads();
  if (tt < TT) {
    float sm = 0.f;
    for (int p = 0; p < nparts; ++p) sm += red[p * TT + tt];
    fin[tt] = sm / (float)C;
  }
  __syncthreads();
  const float mean0 = fin[f0], mean1 = fin[f0 + 1];
  a0 = a1 = 0.f;
  for (int c = part; c < C; c += nparts) {
    const float2 v = load2<T>(s0 + c * stride + f0);
    a0 += (v.x - mean0) * (v.x - mean0);
    a1 += (v.y - mean1) * (v.y - mean1);
  }
  red[part * TT + f0] = a0;
  red[part * TT + f0 + 1] = a1;
  __syncthreads();
  if (tt < TT) {
    float var = 0.f;
    for (int p = 0; p < nparts; ++p) var += red[p * TT + tt];
    const float r = 1.f / (sqrtf(var / (float)(C - 1)) + eps);     // unbiased std, eps on the std (layernorm.py:27)
    fin[TT + tt] = r;
    if (stats && t0 + tt < Tn) {
      const long long col = (long long)b * Tn + t0 + tt;
      stats[col * 2] = fin[tt];
      stats[col * 2 + 1] = r;
    }
  }
  __syncthreads();
  tile_store(y + rb, C, Tn, t0, TT, vec, [&](int c, int j) {
    return gamma[c] * (to_f32<T>(s0[c * stride + j]) - fin[j]) * fin[TT + j] + beta[c];
  });
}

template <typename T>
__global__ void __launch_bounds__(CT_THREADS)
layernorm_bwd_tile(int C, int Tn, TileGeom g, bool vec, const T* x, const float* gamma, const float* stats, float eps,
                   const T* dy, T* dx) {
  CT_PROLOGUE(2);
  tile_load(s0, stride, x + rb, C, Tn, t0, TT, vec);
  tile_load(s1, stride, dy + rb, C, Tn, t0, TT, vec);
  float* fin = red + 2 * nparts * TT;           // [mean | r | k | mu]
  if (tt < TT) {
    const bool ok = t0 + tt < Tn;
    const long long col = (long long)b * Tn + t0 + tt;
    fin[tt] = ok ? stats[col * 2] : 0.f;
    fin[TT + tt] = ok ? stats[col * 2 + 1] : 1.f;
  }
  __syncthreads();
  const float mean0 = fin[f0], mean1 = fin[f0 + 1];
  float p0 = 0.f, p1 = 0.f, q0 = 0.f, q1 = 0.f;       // sum(g), sum(g * xc),  g = dy * gamma
  for (int c = part; c < C; c += nparts) {
    const float gm = gamma[c];
    const float2 gq = load2<T>(s1 + c * stride + f0), v = load2<T>(s0 + c * stride + f0);
    p0 += gq.x * gm;
    p1 += gq.y * gm;
    q0 += gq.x * gm * (v.x - mean0);
    q1 += gq.y * gm * (v.y - mean1);
  }
  red[part * TT + f0] = p0;
  red[part * TT + f0 + 1] = p1;
  red[(nparts + part) * TT + f0] = q0;
  red[(nparts + part) * TT + f0 + 1] = q1;
  __syncthreads();
  if (tt < TT) {
    float a1 = 0.f, a2 = 0.f;
    for (int p = 0; p < nparts; ++p) { a1 += red[p * TT + tt]; a2 += red[(nparts + p) * TT + tt]; }
    const float r = fin[TT + tt];
    const float sd = 1.f / r - eps;
    fin[2 * TT + tt] = (sd > 0.f) ? (-r * r * a2 / ((float)(C - 1) * sd)) : 0.f;
    fin[3 * TT + tt] = r * a1 / (float)C;
  }
  __syncthreads();
  tile_store(dx + rb, C, Tn, t0, TT, vec, [&](int c, int j) {
    const float gq = to_f32<T>(s1[c * stride + j]) * gamma[c];
    const float xc = to_f32<T>(s0[c * stride + j]) - fin[j];
    return gq * fin[TT + j] + fin[2 * TT + j] * xc - fin[3 * TT + j];
  });
}

template <typename T>
__global__ void __launch_bounds__(CT_THREADS)
argmax_tile(int C, int Tn, TileGeom g, bool vec, const T* x, long long* out) {
  CT_PROLOGUE(1);
  tile_load(s0, stride, x + rb, C, Tn, t0, TT, vec);
  __syncthreads();
  float m0 = -INFINITY, m1 = -INFINITY;
  int g0 = C, g1 = C;                            // ascending c within a part: strict > keeps the lowest index
  for (int c = part; c < C; c += nparts) {
    const float2 v = load2<T>(s0 + c * stride + f0);
    if (v.x > m0 || g0 == C) { m0 = v.x; g0 = c; }
    if (v.y > m1 || g1 == C) { m1 = v.y; g1 = c; }
  }
  int* args = reinterpret_cast<int*>(red + nparts * TT);
  red[part * TT + f0] = m0;
  red[part * TT + f0 + 1] = m1;
  args[part * TT + f0] = g0;
  args[part * TT + f0 + 1] = g1;
  __syncthreads();
  if (tt < TT && t0 + tt < Tn) {
    float M = red[tt];
    int A = args[tt];
    for (int p = 1; p < nparts; ++p) {
      const float v = red[p * TT + tt];
      const int a = args[p * TT + tt];
      if (a < C && (A == C || v > M || (v == M && a < A))) { M = v; A = a; }
    }
    out[(long long)b * Tn + t0 + tt] = A;
  }
}

// ------------------------------------------------------------------------------------------ register tiles
// For C <= 256 (the networks' 256 signal levels) the tile never touches shared memory: a CTA of 256 threads owns
// [C channels x 8 vectors of 16 bytes] (64 frames in bf16, 32 in fp32); thread (g = tid / 8, v = tid % 8) keeps vector v of
// channels g, g+32, ..., g+32(KC-1) in registers (KC 16-byte loads in flight per thread), reduces its channels
// locally, meets the 31 other channel groups of its frames through an 8 KB shared array, and stores straight from
// registers.  ~8 instructions per element instead of ~15 for the shared-memory tile, whose budget at 2 B/element is
// ~10 per element at HBM speed.
template <typename T, int KC>
struct RegTile {
  static constexpr int V = 16 / sizeof(T);
  static constexpr int TT = 8 * V;
  uint4 raw[KC];
  bool cok[KC];
  bool tok;
  int g, v;
  __device__ __forceinline__ void load(const T* xb, int C, int Tn, int t0, uint32_t fill) {
    g = threadIdx.x >> 3;
    v = threadIdx.x & 7;
    const int t = t0 + v * V;
    tok = t + V <= Tn;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int c = g + 32 * k;
      cok[k] = c < C;
      raw[k] = make_uint4(fill, fill, fill, fill);
      if (cok[k] && tok) raw[k] = __ldg(reinterpret_cast<const uint4*>(xb + (long long)c * Tn + t));
      else if (!tok) raw[k] = make_uint4(0, 0, 0, 0);
    }
  }
  __device__ __forceinline__ float get(int k, int i) const { return to_f32<T>(reinterpret_cast<const T*>(&raw[k])[i]); }
  // Between two passes over the tile: makes the packed registers opaque to the compiler, so that it re-unpacks them in
  // the next pass instead of carrying 8 KC fp32 copies across (bf16 softmax: 128 registers and two CTAs per SM with the
  // copies, three CTAs without -- these kernels are latency-bound on their loads at 16 warps per SM).
  __device__ __forceinline__ void repack() const {
    uint4* q = const_cast<uint4*>(raw);
#pragma unroll
    for (int k = 0; k < KC; ++k) asm volatile("" : "+r"(q[k].x), "+r"(q[k].y), "+r"(q[k].z), "+r"(q[k].w));
  }
};

// One tile per CTA: tile = blockIdx.x.  (A persistent variant -- 2 CTAs per SM walking tiles with the next tile's
// loads issued ahead of the current tile's reduction -- measured SLOWER, 0.35-0.47 of the copy bandwidth against
// 0.38-0.57: at 2 bytes per element these reductions sit on the issue / MUFU rate, ~6 instructions and one ex2 per
// element against a budget of ~5 per element at HBM speed, and many short CTAs overlap better than a few long ones.)
template <typename T, int KC, typename Body>
__device__ __forceinline__ void reg_tiles(int C, int Tn, int tiles_per_read, int ntiles, const T* x, uint32_t fill,
                                          Body body) {
  using RT = RegTile<T, KC>;
  const int tile = blockIdx.x;
  if (tile >= ntiles) return;
  const int b = tile / tiles_per_read, t0 = (tile - b * tiles_per_read) * RT::TT;
  RT r;
  r.load(x + (long long)b * C * Tn, C, Tn, t0, fill);
  body(r, b, t0);
}

template <typename T> struct NegInf;
template <> struct NegInf<float> { static constexpr uint32_t bits = 0xFF800000u; };
template <> struct NegInf<__nv_bfloat16> { static constexpr uint32_t bits = 0xFF80FF80u; };

// exp(x - m) with one FFMA feeding the ex2: e = 2^(x * log2e - m * log2e); `ml` = m * log2e is formed once per frame.
// bf16 storage only (ex2.approx, like exp_t<bf16>); fp32 storage keeps the exact expf for its 1e-5 parity.
template <typename T>
struct ExpSub {
  static __device__ __forceinline__ float scale(float m) { return m; }
  static __device__ __forceinline__ float eval(float x, float ml) { return expf(x - ml); }
};
template <>
struct ExpSub<__nv_bfloat16> {
  static __device__ __forceinline__ float scale(float m) { return m * 1.4426950408889634f; }
  static __device__ __forceinline__ float eval(float x, float ml) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaf(x, 1.4426950408889634f, -ml)));
    return y;
  }
};

// per-frame maximum over this thread's KC channels; bf16 tiles compare packed pairs (max is exact in any format)
template <typename T, int KC>
__device__ __forceinline__ void tile_max(const RegTile<T, KC>& r, float (&m)[RegTile<T, KC>::V]) {
  constexpr int V = RegTile<T, KC>::V;
#pragma unroll
  for (int i = 0; i < V; ++i) m[i] = -INFINITY;
#pragma unroll
  for (int k = 0; k < KC; ++k)
#pragma unroll
    for (int i = 0; i < V; ++i) m[i] = fmaxf(m[i], r.get(k, i));
}
template <int KC>
__device__ __forceinline__ void tile_max(const RegTile<__nv_bfloat16, KC>& r, float (&m)[8]) {
  __nv_bfloat162 a[4];
  const __nv_bfloat162* p0 = reinterpret_cast<const __nv_bfloat162*>(&r.raw[0]);
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = p0[j];
#pragma unroll
  for (int k = 1; k < KC; ++k) {
    const __nv_bfloat162* pk = reinterpret_cast<const __nv_bfloat162*>(&r.raw[k]);
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = __hmax2(a[j], pk[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(a[j]);
    m[2 * j] = f.x;
    m[2 * j + 1] = f.y;
  }
}

// p[i] (one partial per frame of this thread) -> total over all channels.  Two shuffle steps merge the four channel
// groups of a warp (lanes v, v+8, v+16, v+24), the eight warps meet in shared memory (red: 8 x TT floats), one thread
// per frame merges them.  MAXOP: max instead of sum.
template <int V, bool MAXOP>
__device__ __forceinline__ void frames_combine(float (&p)[V], float* red, float* fin, int v) {
  constexpr int TT = 8 * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < V; ++i) {
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      const float q = __shfl_xor_sync(0xffffffffu, p[i], o);
      p[i] = MAXOP ? fmaxf(p[i], q) : p[i] + q;
    }
    if (lane < 8) red[warp * TT + v * V + i] = p[i];
  }
  __syncthreads();
  if (threadIdx.x < TT) {
    float a = red[threadIdx.x];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float q = red[w * TT + threadIdx.x];
      a = MAXOP ? fmaxf(a, q) : a + q;
    }
    fin[threadIdx.x] = a;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < V; ++i) p[i] = fin[v * V + i];
}

// global max then global sum of exp(x - max) of this thread's V frames (two combines; exps only on the elements)
template <typename T, int KC>
__device__ __forceinline__ void softmax_stats(const RegTile<T, KC>& r, float (&m)[RegTile<T, KC>::V],
                                              float (&s)[RegTile<T, KC>::V], float* red, float* fin) {
  constexpr int V = RegTile<T, KC>::V;
  tile_max(r, m);
  frames_combine<V, true>(m, red, fin, r.v);
  r.repack();
  float ml[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s[i] = 0.f; ml[i] = ExpSub<T>::scale(m[i]); }
#pragma unroll
  for (int k = 0; k < KC; ++k)
#pragma unroll
    for (int i = 0; i < V; ++i) s[i] += ExpSub<T>::eval(r.get(k, i), ml[i]);
  frames_combine<V, false>(s, red + 8 * 8 * V, fin + 8 * V, r.v);
}

template <typename T, int KC>
__global__ void __launch_bounds__(CT_THREADS, 4)
softmax_fwd_reg(int C, int Tn, int tiles_per_read, int ntiles, const T* x, T* y, int log_mode) {
  using RT = RegTile<T, KC>;
  constexpr int V = RT::V, TT = RT::TT;
  __shared__ float red[16 * TT], fin[2 * TT];
  reg_tiles<T, KC>(C, Tn, tiles_per_read, ntiles, x, NegInf<T>::bits, [&](const RT& r, int b, int t0) {
    const long long rb = (long long)b * C * Tn;
    float m[V], sum[V];
    softmax_stats<T, KC>(r, m, sum, red, fin);
    if (!r.tok) return;
    r.repack();
#pragma unroll
    for (int i = 0; i < V; ++i) {
      sum[i] = log_mode ? m[i] + logf(sum[i]) : 1.f / sum[i];
      m[i] = ExpSub<T>::scale(m[i]);
    }
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      if (!r.cok[k]) continue;
      uint4 val;
      T* o = reinterpret_cast<T*>(&val);
#pragma unroll
      for (int i = 0; i < V; ++i)
        o[i] = from_f32<T>(log_mode ? r.get(k, i) - sum[i] : ExpSub<T>::eval(r.get(k, i), m[i]) * sum[i]);
      *reinterpret_cast<uint4*>(y + rb + (long long)(r.g + 32 * k) * Tn + t0 + r.v * V) = val;
    }
  });
}

template <typename T, int KC>
__global__ void __launch_bounds__(CT_THREADS, 4)
xent_fwd_reg(int C, int Tn, int tiles_per_read, int ntiles, const T* x, const long long* target, float* loss_bt,
             float* lse_out) {
  using RT = RegTile<T, KC>;
  constexpr int V = RT::V, TT = RT::TT;
  __shared__ float red[16 * TT], fin[2 * TT], xt[TT];
  reg_tiles<T, KC>(C, Tn, tiles_per_read, ntiles, x, NegInf<T>::bits, [&](const RT& r, int b, int t0) {
    // the target's logit comes straight from global memory (the line is in L1: this CTA has just loaded it)
    if (threadIdx.x < TT) {
      const int t = t0 + threadIdx.x;
      float val = 0.f;
      if (t < Tn) {
        long long tg = target[(long long)b * Tn + t];
        tg = tg < 0 ? 0 : (tg >= C ? C - 1 : tg);
        val = to_f32<T>(__ldg(x + (long long)b * C * Tn + tg * Tn + t));
      }
      xt[threadIdx.x] = val;
    }
    float m[V], sum[V];
    softmax_stats<T, KC>(r, m, sum, red, fin);      // (its barriers also publish xt)
    if (r.g == 0 && r.tok) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float lse = m[i] + logf(sum[i]);
        const long long col = (long long)b * Tn + t0 + r.v * V + i;
        loss_bt[col] = lse - xt[r.v * V + i];
        lse_out[col] = lse;
      }
    }
  });
}

// LayerNorm over channels (layernorm.py:25-28: unbiased std, eps added to the std), register tile
template <typename T, int KC>
__global__ void __launch_bounds__(CT_THREADS, 4)
layernorm_fwd_reg(int C, int Tn, int tiles_per_read, int ntiles, const T* x, const float* gamma, const float* beta,
                  float eps, T* y, float* stats) {
  using RT = RegTile<T, KC>;
  constexpr int V = RT::V, TT = RT::TT;
  __shared__ float red[16 * TT], fin[2 * TT];
  // (gamma / beta are read where they are used: a CTA normalises ONE tile, keeping 2 KC of them in registers from the
  // start only cost occupancy -- 128 registers with spills, 2 CTAs per SM)
  const float invC = 1.f / (float)C, invC1 = 1.f / (float)(C - 1);
  reg_tiles<T, KC>(C, Tn, tiles_per_read, ntiles, x, 0u, [&](const RT& r, int b, int t0) {
    const long long rb = (long long)b * C * Tn;
    float mean[V], rr[V];
#pragma unroll
    for (int i = 0; i < V; ++i) mean[i] = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k)
#pragma unroll
      for (int i = 0; i < V; ++i) mean[i] += r.get(k, i);            // channels >= C were filled with zeros
    frames_combine<V, false>(mean, red, fin, r.v);
    r.repack();
#pragma unroll
    for (int i = 0; i < V; ++i) { mean[i] *= invC; rr[i] = 0.f; }
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      if (!r.cok[k]) continue;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float d = r.get(k, i) - mean[i];
        rr[i] = fmaf(d, d, rr[i]);
      }
    }
    frames_combine<V, false>(rr, red + 8 * 8 * V, fin + 8 * V, r.v);
    if (!r.tok) return;
    r.repack();
#pragma unroll
    for (int i = 0; i < V; ++i) rr[i] = 1.f / (sqrtf(rr[i] * invC1) + eps);
    if (stats && r.g == 0) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const long long col = (long long)b * Tn + t0 + r.v * V + i;
        stats[col * 2] = mean[i];
        stats[col * 2 + 1] = rr[i];
      }
    }
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      if (!r.cok[k]) continue;
      uint4 val;
      T* o = reinterpret_cast<T*>(&val);
#pragma unroll
      const float gk = __ldg(gamma + r.g + 32 * k), bk = __ldg(beta + r.g + 32 * k);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = from_f32<T>(gk * (r.get(k, i) - mean[i]) * rr[i] + bk);
      *reinterpret_cast<uint4*>(y + rb + (long long)(r.g + 32 * k) * Tn + t0 + r.v * V) = val;
    }
  });
}

template <typename T, int KC>
__global__ void __launch_bounds__(CT_THREADS, 3)
xent_bwd_reg(int C, int Tn, int tiles_per_read, const T* x, const long long* target, const float* lse,
             const float* gscale, T* dx) {
  using RT = RegTile<T, KC>;
  constexpr int V = RT::V, TT = RT::TT;
  const int b = blockIdx.x / tiles_per_read, t0 = (blockIdx.x - b * tiles_per_read) * TT;
  const long long rb = (long long)b * C * Tn;
  RT r;
  r.load(x + rb, C, Tn, t0, 0u);
  if (!r.tok) return;
  float l[V];
  int tg[V];
  const long long col0 = (long long)b * Tn + t0 + r.v * V;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    l[i] = lse[col0 + i];
    const long long tv = target[col0 + i];
    tg[i] = (int)(tv < 0 ? 0 : (tv >= C ? C - 1 : tv));       // clamped like the forward
  }
  const float gs = *gscale;
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    if (!r.cok[k]) continue;
    const int c = r.g + 32 * k;
    uint4 val;
    T* o = reinterpret_cast<T*>(&val);
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = from_f32<T>((exp_t<T>(r.get(k, i) - l[i]) - (c == tg[i] ? 1.f : 0.f)) * gs);
    *reinterpret_cast<uint4*>(dx + rb + (long long)c * Tn + t0 + r.v * V) = val;
  }
}

template <typename T, int KC>
__global__ void __launch_bounds__(CT_THREADS, 4)
argmax_reg(int C, int Tn, int tiles_per_read, int ntiles, const T* x, long long* out) {
  using RT = RegTile<T, KC>;
  constexpr int V = RT::V, TT = RT::TT;
  __shared__ float redv[2][8 * TT];              // double-buffered: one barrier per tile
  __shared__ int redc[2][8 * TT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int par = 0;
  reg_tiles<T, KC>(C, Tn, tiles_per_read, ntiles, x, NegInf<T>::bits, [&](const RT& r, int b, int t0) {
    float* rv = redv[par];
    int* rc = redc[par];
    par ^= 1;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float m = r.cok[0] ? r.get(0, i) : -INFINITY;
      int a = r.cok[0] ? r.g : C;                   // ascending channels within a thread: strict > keeps the lowest
#pragma unroll
      for (int k = 1; k < KC; ++k) {
        const float val = r.get(k, i);
        if (r.cok[k] && val > m) { m = val; a = r.g + 32 * k; }
      }
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
        const int a2 = __shfl_xor_sync(0xffffffffu, a, o);
        if (a2 < C && (a >= C || m2 > m || (m2 == m && a2 < a))) { m = m2; a = a2; }
      }
      if (lane < 8) { rv[warp * TT + r.v * V + i] = m; rc[warp * TT + r.v * V + i] = a; }
    }
    __syncthreads();
    if (threadIdx.x < TT && t0 + threadIdx.x < Tn) {
      float M = rv[threadIdx.x];
      int A = rc[threadIdx.x];
#pragma unroll
      for (int q = 1; q < 8; ++q) {
        const float val = rv[q * TT + threadIdx.x];
        const int a = rc[q * TT + threadIdx.x];
        if (a < C && (A >= C || val > M || (val == M && a < A))) { M = val; A = a; }
      }
      out[(long long)b * Tn + t0 + threadIdx.x] = A;
    }
  });
}

// Wide variant for bf16 and C <= 256: [C channels x 128 frames] per CTA, thread (g = tid / 16, v = tid % 16) keeps vector
// v of channels g, g+16, ... (16 loads in flight per thread).  A channel row is 2 T bytes from the next one, so a tile
// touches C different DRAM pages; 256 bytes per visit instead of 128.
template <int KC>
__global__ void __launch_bounds__(CT_THREADS, 2)
argmax_reg_wide(int C, int Tn, int tiles_per_read, int ntiles, const __nv_bfloat16* x, long long* out) {
  constexpr int V = 8, TT = 128;
  __shared__ float redv[8 * TT];
  __shared__ int redc[8 * TT];
  const int tile = blockIdx.x;
  if (tile >= ntiles) return;
  const int b = tile / tiles_per_read, t0 = (tile - b * tiles_per_read) * TT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = threadIdx.x >> 4, v = threadIdx.x & 15;
  const int t = t0 + v * V;
  const bool tok = t + V <= Tn;
  const __nv_bfloat16* xb = x + (long long)b * C * Tn;
  uint4 raw[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int c = g + 16 * k;
    raw[k] = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
    if (c < C && tok) raw[k] = __ldg(reinterpret_cast<const uint4*>(xb + (long long)c * Tn + t));
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float m = -INFINITY;
    int a = C;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const float val = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&raw[k])[i]);
      if (g + 16 * k < C && (a >= C || val > m)) { m = val; a = g + 16 * k; }
    }
    const float m2 = __shfl_xor_sync(0xffffffffu, m, 16);
    const int a2 = __shfl_xor_sync(0xffffffffu, a, 16);
    if (a2 < C && (a >= C || m2 > m || (m2 == m && a2 < a))) { m = m2; a = a2; }
    if (lane < 16) { redv[warp * TT + v * V + i] = m; redc[warp * TT + v * V + i] = a; }
  }
  __syncthreads();
  if (threadIdx.x < TT && t0 + threadIdx.x < Tn) {
    float M = redv[threadIdx.x];
    int A = redc[threadIdx.x];
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      const float val = redv[q * TT + threadIdx.x];
      const int a = redc[q * TT + threadIdx.x];
      if (a < C && (A >= C || val > M || (val == M && a < A))) { M = val; A = a; }
    }
    out[(long long)b * Tn + t0 + threadIdx.x] = A;
  }
}

// ------------------------------------------------------------------------------------------ register tiles, unaligned rows
// Rows whose length is not a multiple of 16 bytes (the train step's T - 1 = 16383 frames, legacy_code/train.py:30) start
// at a different 2-byte phase in every channel, so no 16-byte load lines up with a frame boundary.  The loads stay
// ALIGNED instead and the frames shift: channel c's row starts at element e_c = (b C + c) Tn; a thread loads the aligned
// vector that contains its frames, and position i of vector v holds frame 8v + i - (e_c + t0) mod V of the tile.  That
// shift is the same for channels 8 apart (8 Tn is a multiple of V), so a thread's channels g, g+32, ... share it, and so
// do the four channel groups of a WARP when they are w, w+8, w+16, w+24 -- the in-register reduction over channels and
// the two shuffle steps work unchanged; only the cross-warp step indexes shared memory by frame.  A tile owns 7 vectors
// of frames (56 in bf16) and loads 8 per channel (12.5 % more L2 -> SM traffic, the same HBM traffic).
template <typename T, int KC>
struct RegTileU {
  static constexpr int V = 16 / sizeof(T);
  static constexpr int TT = 8 * V;               // positions loaded per channel (and the shared-memory stride)
  static constexpr int TV = 7 * V;               // frames a tile owns
  uint4 raw[KC];
  bool cok[KC];
  int g, v, shift;
  long long voff;                                // element offset of this thread's vector of channel g
  __device__ __forceinline__ void load(const T* x, long long total, int b, int C, int Tn, int t0, uint32_t fill) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    g = warp + 8 * (lane >> 3);
    v = lane & 7;
    const long long e = ((long long)b * C + g) * Tn + t0;
    shift = (int)(e & (V - 1));
    voff = e - shift + v * V;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int c = g + 32 * k;
      cok[k] = c < C;
      const long long o = voff + (long long)32 * k * Tn;
      raw[k] = make_uint4(fill, fill, fill, fill);
      if (cok[k] && o + V <= total) raw[k] = __ldg(reinterpret_cast<const uint4*>(x + o));
      else if (cok[k]) {                          // the last vector of the tensor: element by element
        T* e1 = reinterpret_cast<T*>(&raw[k]);
        for (int i = 0; i < V; ++i)
          if (o + i < total) e1[i] = x[o + i];
      }
    }
  }
  // frame (within the tile) of position i, or -1 when it belongs to a neighbouring tile / lies past the row
  __device__ __forceinline__ void slots(int (&sl)[V], int Tn, int t0) const {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int j = v * V + i - shift;
      sl[i] = (j >= 0 && j < TV && t0 + j < Tn) ? j : -1;
    }
  }
  __device__ __forceinline__ float get(int k, int i) const { return to_f32<T>(reinterpret_cast<const T*>(&raw[k])[i]); }
  __device__ __forceinline__ void repack() const {        // see RegTile::repack
    uint4* q = const_cast<uint4*>(raw);
#pragma unroll
    for (int k = 0; k < KC; ++k) asm volatile("" : "+r"(q[k].x), "+r"(q[k].y), "+r"(q[k].z), "+r"(q[k].w));
  }
};

template <typename T, int KC>
__device__ __forceinline__ void tile_max_u(const RegTileU<T, KC>& r, float (&m)[RegTileU<T, KC>::V]) {
  constexpr int V = RegTileU<T, KC>::V;
#pragma unroll
  for (int i = 0; i < V; ++i) m[i] = -INFINITY;
#pragma unroll
  for (int k = 0; k < KC; ++k)
#pragma unroll
    for (int i = 0; i < V; ++i) m[i] = fmaxf(m[i], r.get(k, i));        // channels past C hold the -inf fill
}

template <int V, bool MAXOP>
__device__ __forceinline__ void frames_combine_u(float (&p)[V], const int (&sl)[V], float* red, float* fin) {
  constexpr int TT = 8 * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < V; ++i) {
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      const float q = __shfl_xor_sync(0xffffffffu, p[i], o);
      p[i] = MAXOP ? fmaxf(p[i], q) : p[i] + q;
    }
    if (lane < 8 && sl[i] >= 0) red[warp * TT + sl[i]] = p[i];
  }
  __syncthreads();
  if (threadIdx.x < 7 * V) {                     // every warp has written every owned frame (shift < V)
    float a = red[threadIdx.x];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float q = red[w * TT + threadIdx.x];
      a = MAXOP ? fmaxf(a, q) : a + q;
    }
    fin[threadIdx.x] = a;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < V; ++i) p[i] = sl[i] >= 0 ? fin[sl[i]] : (MAXOP ? 0.f : 1.f);
}

template <typename T, int KC>
__device__ __forceinline__ void softmax_stats_u(const RegTileU<T, KC>& r, const int (&sl)[RegTileU<T, KC>::V],
                                                float (&m)[RegTileU<T, KC>::V], float (&s)[RegTileU<T, KC>::V],
                                                float* red, float* fin) {
  constexpr int V = RegTileU<T, KC>::V;
  tile_max_u(r, m);
  frames_combine_u<V, true>(m, sl, red, fin);
  r.repack();
  float ml[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s[i] = 0.f; ml[i] = ExpSub<T>::scale(m[i]); }
#pragma unroll
  for (int k = 0; k < KC; ++k)
#pragma unroll
    for (int i = 0; i < V; ++i) s[i] += ExpSub<T>::eval(r.get(k, i), ml[i]);
  frames_combine_u<V, false>(s, sl, red + 8 * 8 * V, fin + 8 * V);
}

template <typename T, int KC>
__global__ void __launch_bounds__(CT_THREADS, 4)
xent_fwd_regu(int C, int Tn, int tiles_per_read, int ntiles, long long total, const T* x, const long long* target,
              float* loss_bt, float* lse_out) {
  using RT = RegTileU<T, KC>;
  constexpr int V = RT::V, TT = RT::TT, TV = RT::TV;
  __shared__ float red[16 * TT], fin[2 * TT];
  const int tile = blockIdx.x;
  if (tile >= ntiles) return;
  const int b = tile / tiles_per_read, t0 = (tile - b * tiles_per_read) * TV;
  RT r;
  r.load(x, total, b, C, Tn, t0, NegInf<T>::bits);
  // the target's logit comes straight from global memory (the line is in L1 / L2: this CTA loads it)
  float xt = 0.f;
  const int tf = t0 + (int)threadIdx.x;
  if (threadIdx.x < TV && tf < Tn) {
    long long tg = target[(long long)b * Tn + tf];
    tg = tg < 0 ? 0 : (tg >= C ? C - 1 : tg);
    xt = to_f32<T>(__ldg(x + ((long long)b * C + tg) * Tn + tf));
  }
  int sl[V];
  r.slots(sl, Tn, t0);
  float m[V], sum[V];
  softmax_stats_u<T, KC>(r, sl, m, sum, red, fin);
  if (threadIdx.x < TV && tf < Tn) {              // one thread per owned frame: fin = (max | sum) per frame
    const float lse = fin[threadIdx.x] + logf(fin[8 * V + threadIdx.x]);
    const long long col = (long long)b * Tn + tf;
    loss_bt[col] = lse - xt;
    lse_out[col] = lse;
  }
}

template <typename T, int KC>
__global__ void __launch_bounds__(CT_THREADS, 3)
xent_bwd_regu(int C, int Tn, int tiles_per_read, int ntiles, long long total, const T* x, const long long* target,
              const float* lse, const float* gscale, T* dx) {
  using RT = RegTileU<T, KC>;
  constexpr int V = RT::V, TV = RT::TV;
  const int tile = blockIdx.x;
  if (tile >= ntiles) return;
  const int b = tile / tiles_per_read, t0 = (tile - b * tiles_per_read) * TV;
  RT r;
  r.load(x, total, b, C, Tn, t0, 0u);
  int sl[V];
  r.slots(sl, Tn, t0);
  float l[V];
  int tg[V];
  bool any = false, all = true;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    l[i] = 0.f;
    tg[i] = -1;
    if (sl[i] >= 0) {
      const long long col = (long long)b * Tn + t0 + sl[i];
      l[i] = lse[col];
      const long long tv = target[col];
      tg[i] = (int)(tv < 0 ? 0 : (tv >= C ? C - 1 : tv));       // clamped like the forward
      any = true;
    } else {
      all = false;
    }
  }
  if (!any) return;
  const float gs = *gscale;
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    if (!r.cok[k]) continue;
    const int c = r.g + 32 * k;
    uint4 val;
    T* o = reinterpret_cast<T*>(&val);
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = from_f32<T>((exp_t<T>(r.get(k, i) - l[i]) - (c == tg[i] ? 1.f : 0.f)) * gs);
    T* dst = dx + r.voff + (long long)32 * k * Tn;
    if (all) {
      *reinterpret_cast<uint4*>(dst) = val;       // dx has x's layout and alignment
    } else {                                      // a vector that straddles a tile or row boundary: its own frames only
#pragma unroll
      for (int i = 0; i < V; ++i)
        if (sl[i] >= 0) dst[i] = o[i];
    }
  }
}

template <typename T, int KC>
__global__ void __launch_bounds__(CT_THREADS, 4)
argmax_regu(int C, int Tn, int tiles_per_read, int ntiles, long long total, const T* x, long long* out) {
  using RT = RegTileU<T, KC>;
  constexpr int V = RT::V, TT = RT::TT, TV = RT::TV;
  __shared__ float redv[8 * TT];
  __shared__ int redc[8 * TT];
  const int tile = blockIdx.x;
  if (tile >= ntiles) return;
  const int b = tile / tiles_per_read, t0 = (tile - b * tiles_per_read) * TV;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RT r;
  r.load(x, total, b, C, Tn, t0, NegInf<T>::bits);
  int sl[V];
  r.slots(sl, Tn, t0);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float m = r.cok[0] ? r.get(0, i) : -INFINITY;
    int a = r.cok[0] ? r.g : C;                     // ascending channels within a thread: strict > keeps the lowest
#pragma unroll
    for (int k = 1; k < KC; ++k) {
      const float val = r.get(k, i);
      if (r.cok[k] && val > m) { m = val; a = r.g + 32 * k; }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
      const int a2 = __shfl_xor_sync(0xffffffffu, a, o);
      if (a2 < C && (a >= C || m2 > m || (m2 == m && a2 < a))) { m = m2; a = a2; }
    }
    if (lane < 8 && sl[i] >= 0) { redv[warp * TT + sl[i]] = m; redc[warp * TT + sl[i]] = a; }
  }
  __syncthreads();
  if (threadIdx.x < TV && t0 + threadIdx.x < Tn) {
    float M = redv[threadIdx.x];
    int A = redc[threadIdx.x];
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      const float val = redv[q * TT + threadIdx.x];
      const int a = redc[q * TT + threadIdx.x];
      if (a < C && (A >= C || val > M || (val == M && a < A))) { M = val; A = a; }
    }
    out[(long long)b * Tn + t0 + threadIdx.x] = A;
  }
}

// ------------------------------------------------------------------------------------------ host side
// largest TT in {256 .. 8} such that ntiles tiles of [C x (TT + pad)] elements + scratch fit in `budget` bytes
static bool pick_geom(int C, int Tn, int esize, int ntiles, TileGeom* g, size_t* smem) {
  const int V = 16 / esize;
  for (int TT = 256; TT >= 8; TT >>= 1) {
    if (TT > 8 && TT / 2 >= Tn) continue;                       // do not over-tile short reads
    const int stride = TT + V;
    const int nparts = 2 * CT_THREADS / TT;
    const size_t bytes = (size_t)ntiles * C * stride * esize + sizeof(float) * (size_t)((2 * nparts + 4) * TT);
    if (bytes <= 64 * 1024 || (TT == 8 && bytes <= 200 * 1024)) {
      g->TT = TT; g->stride = stride; g->nparts = nparts;
      g->tiles_per_read = (Tn + TT - 1) / TT;
      *smem = bytes;
      return true;
    }
  }
  return false;
}

static inline bool vec_ok(const void* p0, const void* p1, const void* p2, int Tn, int esize) {
  auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return ((long long)Tn * esize) % 16 == 0 && al(p0) && al(p1) && al(p2);
}

}  // namespace wnb

using namespace wnb;
typedef __nv_bfloat16 bf16;

// fallbacks (per-column kernels, elementwise.cu) for channel counts whose 8-frame tile does not fit in shared memory
extern "C" int wnb200_softmax_fwd_col(int, int, int, int, const void*, void*, int, void*);
extern "C" int wnb200_softmax_bwd_col(int, int, int, int, const void*, const void*, void*, int, void*);
extern "C" int wnb200_xent_fwd_col(int, int, int, int, const void*, const int64_t*, float*, float*, void*);
extern "C" int wnb200_xent_bwd_col(int, int, int, int, const void*, const int64_t*, const float*, const float*, void*,
                                   void*);
extern "C" int wnb200_layernorm_fwd_col(int, int, int, int, const void*, const float*, const float*, float, void*,
                                        float*, void*);
extern "C" int wnb200_layernorm_bwd_col(int, int, int, int, const void*, const float*, const float*, float,
                                        const void*, void*, void*);
extern "C" int wnb200_argmax_channels_col(int, int, int, int, const void*, int64_t*, void*);

// register-tile path: C <= 256, 16-byte aligned rows
#define RT_TRY(KERNEL, P0, P1, ...)                                                                        \
  do {                                                                                                     \
    const int esize_ = dtype == WNB200_F32 ? 4 : 2;                                                        \
    if (C <= 256 && (dtype == WNB200_F32 || dtype == WNB200_BF16) && vec_ok(P0, P1, nullptr, T_, esize_)) { \
      const int TT_ = 8 * (16 / esize_);                                                                   \
      const int tiles_ = (T_ + TT_ - 1) / TT_;                                                             \
      const unsigned grid_ = (unsigned)((long long)B * tiles_);                                            \
      cudaStream_t st_ = (cudaStream_t)stream;                                                             \
      const int kc_ = C <= 64 ? 2 : (C <= 128 ? 4 : 8);                                                    \
      if (dtype == WNB200_F32) {                                                                           \
        using T = float;                                                                                   \
        if (kc_ == 2) KERNEL<T, 2><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, __VA_ARGS__);              \
        else if (kc_ == 4) KERNEL<T, 4><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, __VA_ARGS__);         \
        else KERNEL<T, 8><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, __VA_ARGS__);                       \
      } else {                                                                                             \
        using T = bf16;                                                                                    \
        if (kc_ == 2) KERNEL<T, 2><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, __VA_ARGS__);              \
        else if (kc_ == 4) KERNEL<T, 4><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, __VA_ARGS__);         \
        else KERNEL<T, 8><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, __VA_ARGS__);                       \
      }                                                                                                    \
      WNB_LAUNCH_OK();                                                                                     \
      return 0;                                                                                            \
    }                                                                                                      \
  } while (0)

static inline int rt_persistent_grid(long long ntiles) { return (int)ntiles; }
#define RT_TRY_P(KERNEL, P0, P1, ...)                                                                      \
  do {                                                                                                     \
    const int esize_ = dtype == WNB200_F32 ? 4 : 2;                                                        \
    const int TT_ = 8 * (16 / esize_);                                                                     \
    const long long ntiles_ = (long long)B * ((T_ + TT_ - 1) / TT_);                                       \
    if (C <= 256 && (dtype == WNB200_F32 || dtype == WNB200_BF16) && vec_ok(P0, P1, nullptr, T_, esize_) && \
        ntiles_ < (1LL << 31)) {                                                                           \
      const int tiles_ = (T_ + TT_ - 1) / TT_;                                                             \
      const unsigned grid_ = (unsigned)rt_persistent_grid(ntiles_);                                        \
      const int nt_ = (int)ntiles_;                                                                        \
      cudaStream_t st_ = (cudaStream_t)stream;                                                             \
      const int kc_ = C <= 64 ? 2 : (C <= 128 ? 4 : 8);                                                    \
      if (dtype == WNB200_F32) {                                                                           \
        using T = float;                                                                                   \
        if (kc_ == 2) KERNEL<T, 2><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, nt_, __VA_ARGS__);         \
        else if (kc_ == 4) KERNEL<T, 4><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, nt_, __VA_ARGS__);    \
        else KERNEL<T, 8><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, nt_, __VA_ARGS__);                  \
      } else {                                                                                             \
        using T = bf16;                                                                                    \
        if (kc_ == 2) KERNEL<T, 2><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, nt_, __VA_ARGS__);         \
        else if (kc_ == 4) KERNEL<T, 4><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, nt_, __VA_ARGS__);    \
        else KERNEL<T, 8><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, nt_, __VA_ARGS__);                  \
      }                                                                                                    \
      WNB_LAUNCH_OK();                                                                                     \
      return 0;                                                                                            \
    }                                                                                                      \
  } while (0)

// register tiles on rows that are NOT 16-byte multiples (aligned loads, shifted frames: RegTileU)
#define RT_TRY_U(KERNEL, P0, P1, ...)                                                                      \
  do {                                                                                                     \
    const int esize_ = dtype == WNB200_F32 ? 4 : 2;                                                        \
    const int TV_ = 7 * (16 / esize_);                                                                     \
    const long long ntiles_ = (long long)B * ((T_ + TV_ - 1) / TV_);                                       \
    auto al_ = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };    \
    if (C <= 256 && (dtype == WNB200_F32 || dtype == WNB200_BF16) && al_(P0) && al_(P1) &&                 \
        ((long long)T_ * esize_) % 16 != 0 && ntiles_ < (1LL << 31)) {                                     \
      const int tiles_ = (T_ + TV_ - 1) / TV_;                                                             \
      const unsigned grid_ = (unsigned)ntiles_;                                                            \
      const int nt_ = (int)ntiles_;                                                                        \
      const long long total_ = (long long)B * C * T_;                                                      \
      cudaStream_t st_ = (cudaStream_t)stream;                                                             \
      const int kc_ = C <= 64 ? 2 : (C <= 128 ? 4 : 8);                                                    \
      if (dtype == WNB200_F32) {                                                                           \
        using T = float;                                                                                   \
        if (kc_ == 2) KERNEL<T, 2><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, nt_, total_, __VA_ARGS__);  \
        else if (kc_ == 4) KERNEL<T, 4><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, nt_, total_, __VA_ARGS__); \
        else KERNEL<T, 8><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, nt_, total_, __VA_ARGS__);           \
      } else {                                                                                             \
        using T = bf16;                                                                                    \
        if (kc_ == 2) KERNEL<T, 2><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, nt_, total_, __VA_ARGS__);  \
        else if (kc_ == 4) KERNEL<T, 4><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, nt_, total_, __VA_ARGS__); \
        else KERNEL<T, 8><<<grid_, CT_THREADS, 0, st_>>>(C, T_, tiles_, nt_, total_, __VA_ARGS__);           \
      }                                                                                                    \
      WNB_LAUNCH_OK();                                                                                     \
      return 0;                                                                                            \
    }                                                                                                      \
  } while (0)

#define CT_LAUNCH(KERNEL, NTILES, P0, P1, P2, FALLBACK, ...)                                               \
  do {                                                                                                     \
    const int esize = dtype == WNB200_F32 ? 4 : 2;                                                         \
    TileGeom g;                                                                                            \
    size_t smem;                                                                                           \
    if (dtype != WNB200_F32 && dtype != WNB200_BF16) { set_error(#KERNEL ": bad dtype %d", dtype); return 1; } \
    if (!pick_geom(C, T_, esize, NTILES, &g, &smem)) return FALLBACK;                                      \
    const bool vec = vec_ok(P0, P1, P2, T_, esize);                                                        \
    const unsigned grid = (unsigned)((long long)B * g.tiles_per_read);                                     \
    cudaStream_t st = (cudaStream_t)stream;                                                                \
    if (dtype == WNB200_F32) {                                                                             \
      using T = float;                                                                                     \
      if (smem > 48 * 1024)                                                                                \
        WNB_CUDA_OK(cudaFuncSetAttribute(KERNEL<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      KERNEL<T><<<grid, CT_THREADS, smem, st>>>(C, T_, g, vec, __VA_ARGS__);                                \
    } else {                                                                                               \
      using T = bf16;                                                                                      \
      if (smem > 48 * 1024)                                                                                \
        WNB_CUDA_OK(cudaFuncSetAttribute(KERNEL<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      KERNEL<T><<<grid, CT_THREADS, smem, st>>>(C, T_, g, vec, __VA_ARGS__);                                \
    }                                                                                                      \
    WNB_LAUNCH_OK();                                                                                       \
    return 0;                                                                                              \
  } while (0)

extern "C" int wnb200_softmax_fwd(int dtype, int B, int C, int T_, const void* x, void* y, int log_mode,
                                  void* stream) {
  WNB_CHECK_ARG(x && y && C >= 1, "softmax_fwd: bad args");
  if ((long long)B * T_ == 0) return 0;
  RT_TRY_P(softmax_fwd_reg, x, y, (const T*)x, (T*)y, log_mode);
  CT_LAUNCH(softmax_fwd_tile, 1, x, y, nullptr, wnb200_softmax_fwd_col(dtype, B, C, T_, x, y, log_mode, stream),
            (const T*)x, (T*)y, log_mode);
}

extern "C" int wnb200_softmax_bwd(int dtype, int B, int C, int T_, const void* y, const void* dy, void* dx,
                                  int log_mode, void* stream) {
  WNB_CHECK_ARG(y && dy && dx, "softmax_bwd: null pointer");
  if ((long long)B * T_ == 0) return 0;
  CT_LAUNCH(softmax_bwd_tile, 2, y, dy, dx, wnb200_softmax_bwd_col(dtype, B, C, T_, y, dy, dx, log_mode, stream),
            (const T*)y, (const T*)dy, (T*)dx, log_mode);
}

extern "C" int wnb200_xent_fwd(int dtype, int B, int C, int T_, const void* logits, const int64_t* target,
                               float* loss_bt, float* lse, void* stream) {
  WNB_CHECK_ARG(logits && target && loss_bt && lse, "xent_fwd: null pointer");
  if ((long long)B * T_ == 0) return 0;
  RT_TRY_P(xent_fwd_reg, logits, nullptr, (const T*)logits, (const long long*)target, loss_bt, lse);
  RT_TRY_U(xent_fwd_regu, logits, nullptr, (const T*)logits, (const long long*)target, loss_bt, lse);
  CT_LAUNCH(xent_fwd_tile, 1, logits, nullptr, nullptr,
            wnb200_xent_fwd_col(dtype, B, C, T_, logits, target, loss_bt, lse, stream), (const T*)logits,
            (const long long*)target, loss_bt, lse);
}

extern "C" int wnb200_xent_bwd(int dtype, int B, int C, int T_, const void* logits, const int64_t* target,
                               const float* lse, const float* gscale, void* dlogits, void* stream) {
  WNB_CHECK_ARG(logits && target && lse && gscale && dlogits, "xent_bwd: null pointer");
  if ((long long)B * T_ == 0) return 0;
  RT_TRY(xent_bwd_reg, logits, dlogits, (const T*)logits, (const long long*)target, lse, gscale, (T*)dlogits);
  RT_TRY_U(xent_bwd_regu, logits, dlogits, (const T*)logits, (const long long*)target, lse, gscale, (T*)dlogits);
  CT_LAUNCH(xent_bwd_tile, 1, logits, dlogits, nullptr,
            wnb200_xent_bwd_col(dtype, B, C, T_, logits, target, lse, gscale, dlogits, stream), (const T*)logits,
            (const long long*)target, lse, gscale, (T*)dlogits);
}

extern "C" int wnb200_layernorm_fwd(int dtype, int B, int C, int T_, const void* x, const float* gamma,
                                    const float* beta, float eps, void* y, float* stats, void* stream) {
  WNB_CHECK_ARG(x && y && gamma && beta && C >= 2, "layernorm_fwd: bad args");
  if ((long long)B * T_ == 0) return 0;
  RT_TRY_P(layernorm_fwd_reg, x, y, (const T*)x, gamma, beta, eps, (T*)y, stats);
  CT_LAUNCH(layernorm_fwd_tile, 1, x, y, nullptr,
            wnb200_layernorm_fwd_col(dtype, B, C, T_, x, gamma, beta, eps, y, stats, stream), (const T*)x, gamma, beta,
            eps, (T*)y, stats);
}

extern "C" int wnb200_layernorm_bwd(int dtype, int B, int C, int T_, const void* x, const float* gamma,
                                    const float* stats, float eps, const void* dy, void* dx, void* stream) {
  WNB_CHECK_ARG(x && gamma && stats && dy && dx, "layernorm_bwd: null pointer");
  if ((long long)B * T_ == 0) return 0;
  CT_LAUNCH(layernorm_bwd_tile, 2, x, dy, dx,
            wnb200_layernorm_bwd_col(dtype, B, C, T_, x, gamma, stats, eps, dy, dx, stream), (const T*)x, gamma, stats,
            eps, (const T*)dy, (T*)dx);
}

extern "C" int wnb200_argmax_channels(int dtype, int B, int C, int T_, const void* x, int64_t* out, void* stream) {
  WNB_CHECK_ARG(x && out && C >= 1, "argmax_channels: bad args");
  if ((long long)B * T_ == 0) return 0;
  // A/B switch: the 64-frame register tile at four CTAs per SM measured 0.53 of the copy peak, the wide tile 0.60
  static const bool use_wide = getenv("WNB200_ARGMAX_NARROW") == nullptr;
  if (use_wide && dtype == WNB200_BF16 && C > 128 && C <= 256 && vec_ok(x, nullptr, nullptr, T_, 2) && T_ >= 1024 &&
      (long long)B * ((T_ + 127) / 128) < (1LL << 31)) {
    const int tiles = (T_ + 127) / 128;
    argmax_reg_wide<16><<<(unsigned)((long long)B * tiles), CT_THREADS, 0, (cudaStream_t)stream>>>(
        C, T_, tiles, B * tiles, (const bf16*)x, (long long*)out);
    WNB_LAUNCH_OK();
    return 0;
  }
  RT_TRY_P(argmax_reg, x, nullptr, (const T*)x, (long long*)out);
  RT_TRY_U(argmax_regu, x, nullptr, (const T*)x, (long long*)out);
  CT_LAUNCH(argmax_tile, 1, x, nullptr, nullptr, wnb200_argmax_channels_col(dtype, B, C, T_, x, out, stream),
            (const T*)x, (long long*)out);
}
