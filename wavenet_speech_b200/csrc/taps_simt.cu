// Generic tap-sum contraction on CUDA cores (fp32 FFMA accumulate), NCL layout.
//
//   out[b, m, t] (+)= epi( bias[m] + sum_s sum_c w_s[m, c] * pre(x_s[b, c, t + off_s]) )
//
// This is the any-shape path behind CausalConv1d / NonCausalConv1d / ResidualBlock /
// output stacks / LinearConv1d (reference: modules/conv_ops.py:39-44,74-79; block.py:54-82;
// wavenet.py:93-103; linear_conv_ops.py:39-68) and -- with transposed weight slabs and negated
// offsets -- their data gradients.  fp32 storage gives the <=1e-5 parity mode (TF32 tensor
// cores would not); bf16 storage shares the kernel.  The tcgen05 path lives in resblock_tc.cu.
//
// Tiling: 128 weight rows x 128 frames per CTA, K slabs of 8 channels, 256 threads, 8x8
// register micro-tiles, register-prefetch double buffering through shared memory.
// Zero padding of the reference (conv1d padding at the true sequence ends) = the bounds check
// on t + off_s; the columns the reference computes and slices away are never computed.
#include <string.h>

#include "common.cuh"

namespace wnb {

constexpr int BM = 128, BN = 128, BK = 8, NT = 256;

struct SrcDev {
  const void* x;
  const void* w;
  long long bs, cs;
  int C, T_src, t_off, pre_act;
  // pre_act == WNB200_PRE_LNRELU: relu(gamma[c] * (x - mean[b,t]) * r[b,t] + beta[c]) as the source is loaded
  const float* ln_stats;   // [B, T_src, 2] = (mean, 1 / (std + eps))
  const float* ln_gamma;   // [C]
  const float* ln_beta;    // [C]
};

struct TapsParams {
  int B, T_out, M, rows, nsrc;
  SrcDev src[WNB200_MAX_SRC];
  const float* bias;
  int accumulate;
  void* out;
  void* th;
  void* sg;
  const void* residual;    // [B, M, T_out] added to the result (ByteNet blocks' `seq + stack(seq)`), or null
  const void* mu_h;        // EPI_MU: h of g1 * tanh(g2 * h + g3 * u), [B, M, T_out]
  int mu_h_ln;             // EPI_MU: mu_h is source 0's raw tensor, its LayerNorm + ReLU is applied on the fly
};

// the transform a source undergoes on its way into the contraction
__device__ __forceinline__ float pre_transform(const SrcDev& S, float v, int b, int c, int t) {
  if (S.pre_act == WNB200_PRE_LEAKY) return leaky(v);
  if (S.pre_act == WNB200_PRE_LNRELU) {
    const float2 st = *reinterpret_cast<const float2*>(S.ln_stats + ((long long)b * S.T_src + t) * 2);
    return fmaxf(S.ln_gamma[c] * (v - st.x) * st.y + S.ln_beta[c], 0.f);
  }
  return v;
}

template <typename T>
__device__ __forceinline__ void add_row4(const T* base, long long idx, int t, int T_out, float v[4], bool vec_ok) {
  if (t >= T_out) return;
  const T* p = base + idx;
  if (vec_ok && t + 3 < T_out) {
    if constexpr (sizeof(T) == 4) {
      const float4 r = *reinterpret_cast<const float4*>(p);
      v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] += to_f32<T>(p[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (t + i < T_out) v[i] += to_f32<T>(p[i]);
  }
}

template <typename T>
__device__ __forceinline__ void store_row4(T* base, long long idx, int t, int T_out, const float v[4],
                                           bool accumulate, bool vec_ok) {
  // base[idx + 0..3] for frames t..t+3 (those < T_out)
  if (t >= T_out) return;
  T* p = base + idx;
  if (vec_ok && t + 3 < T_out) {
    if constexpr (sizeof(T) == 4) {
      float4 o = make_float4(v[0], v[1], v[2], v[3]);
      if (accumulate) {
        float4 old = *reinterpret_cast<const float4*>(p);
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      *reinterpret_cast<float4*>(p) = o;
    } else {
      float w[4] = {v[0], v[1], v[2], v[3]};
      if (accumulate) {
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] += to_f32<T>(p[i]);
      }
      __nv_bfloat162 lo = __floats2bfloat162_rn(w[0], w[1]);
      __nv_bfloat162 hi = __floats2bfloat162_rn(w[2], w[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(p) = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (t + i < T_out) {
        float o = v[i];
        if (accumulate) o += to_f32<T>(p[i]);
        p[i] = from_f32<T>(o);
      }
    }
  }
}

template <typename T, int EPI, bool LN>
__global__ void __launch_bounds__(NT, 2) taps_fwd_kernel(const TapsParams p) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int b = blockIdx.z;
  const int row0 = blockIdx.y * BM;
  const int t0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int a_row = tid >> 1, a_k = (tid & 1) * 4;
  const int b_k = tid >> 5, b_t = (tid & 31) * 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // Register prefetch: the loads of slab i+1 are ISSUED before the FFMA loop of slab i and CONSUMED (converted,
  // transformed, written to shared memory) after it.  Nothing between the issue and the loop may read the loaded
  // registers -- round 2's profile (profiles/r2_ncu_full_taps_*.csv) showed the kernel at 38 % of the FFMA pipe with
  // its largest stall on the LeakyReLU compare that used to sit right behind the load.
  T rat[4], rbt[4];
  unsigned rb_ok = 0;
  int ld_s = 0, ld_c = 0;
  float2 lst[4];          // LN only: (mean, 1/(std+eps)) of the four frames
  float lg = 0.f, lb = 0.f;

  auto load_regs = [&](int s, int k0) {
    const SrcDev& S = p.src[s];
    const int grow = row0 + a_row;
    const T* w = reinterpret_cast<const T*>(S.w) + (long long)grow * S.C + k0 + a_k;
    const bool row_ok = grow < p.rows;
    if (row_ok && (S.C & 3) == 0 && k0 + a_k + 3 < S.C) {
      if constexpr (sizeof(T) == 4) {
        const float4 q = *reinterpret_cast<const float4*>(w);
        rat[0] = q.x; rat[1] = q.y; rat[2] = q.z; rat[3] = q.w;
      } else {
        const uint2 q = *reinterpret_cast<const uint2*>(w);
        *reinterpret_cast<uint2*>(&rat[0]) = q;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        rat[i] = from_f32<T>(0.f);
        if (row_ok && k0 + a_k + i < S.C) rat[i] = w[i];
      }
    }
    const int c = k0 + b_k;
    const int ts = t0 + b_t + S.t_off;
    const T* x = reinterpret_cast<const T*>(S.x) + (long long)b * S.bs + (long long)c * S.cs + ts;
    ld_s = s; ld_c = c;
    rb_ok = 0;
    if (c < S.C) {
      if (ts >= 0 && ts + 3 < S.T_src && ((reinterpret_cast<uintptr_t>(x) & (4 * sizeof(T) - 1)) == 0)) {
        rb_ok = 15u;
        if constexpr (sizeof(T) == 4) {
          const float4 q = *reinterpret_cast<const float4*>(x);
          rbt[0] = q.x; rbt[1] = q.y; rbt[2] = q.z; rbt[3] = q.w;
        } else {
          *reinterpret_cast<uint2*>(&rbt[0]) = *reinterpret_cast<const uint2*>(x);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          rbt[i] = from_f32<T>(0.f);
          if (ts + i >= 0 && ts + i < S.T_src) { rbt[i] = x[i]; rb_ok |= 1u << i; }
        }
      }
      if constexpr (LN) {
        if (S.pre_act == WNB200_PRE_LNRELU) {
          lg = S.ln_gamma[c]; lb = S.ln_beta[c];
          const float2* stp = reinterpret_cast<const float2*>(S.ln_stats) + (long long)b * S.T_src + ts;
#pragma unroll
          for (int i = 0; i < 4; ++i) lst[i] = (rb_ok >> i & 1u) ? stp[i] : make_float2(0.f, 0.f);
        }
      }
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) As[buf][a_k + i][a_row] = to_f32<T>(rat[i]);
    const int pre = p.src[ld_s].pre_act;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float q = (rb_ok >> i & 1u) ? to_f32<T>(rbt[i]) : 0.f;
      if (pre == WNB200_PRE_LEAKY) q = leaky(q);
      if constexpr (LN) {
        if (pre == WNB200_PRE_LNRELU) q = (rb_ok >> i & 1u) ? fmaxf(lg * (q - lst[i].x) * lst[i].y + lb, 0.f) : 0.f;
      }
      v[i] = q;
    }
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_t]) = make_float4(v[0], v[1], v[2], v[3]);
  };

  int niter = 0;
  for (int s = 0; s < p.nsrc; ++s) niter += (p.src[s].C + BK - 1) / BK;

  int s = 0, k0 = 0;
  load_regs(s, k0);
  store_smem(0);
  __syncthreads();

  for (int it = 0; it < niter; ++it) {
    const int buf = it & 1;
    const bool more = (it + 1 < niter);
    if (more) {
      k0 += BK;
      if (k0 >= p.src[s].C) { k0 = 0; ++s; }
      load_regs(s, k0);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) store_smem(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue ----
  const bool vec_ok = (p.T_out % 4 == 0);
  T* out = reinterpret_cast<T*>(p.out);
  if constexpr (EPI == WNB200_EPI_GATE) {
    T* th = reinterpret_cast<T*>(p.th);
    T* sg = reinterpret_cast<T*>(p.sg);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ch = blockIdx.y * 64 + ty * 4 + i;
      if (ch >= p.M) continue;
      const float bt = p.bias ? p.bias[row0 + ty * 4 + i] : 0.f;
      const float bsg = p.bias ? p.bias[row0 + 64 + ty * 4 + i] : 0.f;
      const long long rowbase = ((long long)b * p.M + ch) * p.T_out;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int t = t0 + h * 64 + tx * 4;
        float vt[4], vs[4], vo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          vt[j] = tanhf(acc[i][h * 4 + j] + bt);
          vs[j] = sigmoid_precise(acc[i + 4][h * 4 + j] + bsg);
          vo[j] = vt[j] * vs[j];
        }
        store_row4<T>(out, rowbase + t, t, p.T_out, vo, p.accumulate != 0, vec_ok);
        if (th) store_row4<T>(th, rowbase + t, t, p.T_out, vt, false, vec_ok);
        if (sg) store_row4<T>(sg, rowbase + t, t, p.T_out, vs, false, vec_ok);
      }
    }
  } else if constexpr (EPI == WNB200_EPI_MU) {
    // MultiplicativeUnit (block.py:213-220): a 128-row tile holds 32 channels x (gate1, gate2, gate3, update); a thread's
    // rows [ty*4, ty*4+4) are the four pre-activations of channel ty, rows [64+ty*4, ..+4) those of channel 16+ty
    const T* hsrc = reinterpret_cast<const T*>(p.mu_h);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int ch = blockIdx.y * 32 + half * 16 + ty;
      if (ch >= p.M) continue;
      const int r0 = row0 + half * 64 + ty * 4;
      float bi[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) bi[q] = p.bias ? p.bias[r0 + q] : 0.f;
      const long long rowbase = ((long long)b * p.M + ch) * p.T_out;
      const SrcDev& S0 = p.src[0];
      const T* hrow = p.mu_h_ln ? hsrc + (long long)b * S0.bs + (long long)ch * S0.cs : hsrc + rowbase;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int t = t0 + h * 64 + tx * 4;
        float vo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          vo[j] = 0.f;
          if (t + j < p.T_out) {
            float hv = to_f32<T>(hrow[t + j]);
            if (p.mu_h_ln) hv = pre_transform(S0, hv, b, ch, t + j);
            const float g1 = sigmoid_precise(acc[half * 4 + 0][h * 4 + j] + bi[0]);
            const float g2 = sigmoid_precise(acc[half * 4 + 1][h * 4 + j] + bi[1]);
            const float g3 = sigmoid_precise(acc[half * 4 + 2][h * 4 + j] + bi[2]);
            const float u = tanhf(acc[half * 4 + 3][h * 4 + j] + bi[3]);
            vo[j] = g1 * tanhf(g2 * hv + g3 * u);
          }
        }
        if (p.residual) add_row4<T>(reinterpret_cast<const T*>(p.residual), rowbase + t, t, p.T_out, vo, vec_ok);
        store_row4<T>(out, rowbase + t, t, p.T_out, vo, p.accumulate != 0, vec_ok);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (m >= p.M) continue;
      const float bi = p.bias ? p.bias[m] : 0.f;
      const long long rowbase = ((long long)b * p.M + m) * p.T_out;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int t = t0 + h * 64 + tx * 4;
        float vo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = acc[i][h * 4 + j] + bi;
          if (EPI == WNB200_EPI_LEAKY) v = leaky(v);
          vo[j] = v;
        }
        if (p.residual) add_row4<T>(reinterpret_cast<const T*>(p.residual), rowbase + t, t, p.T_out, vo, vec_ok);
        store_row4<T>(out, rowbase + t, t, p.T_out, vo, p.accumulate != 0, vec_ok);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// weight gradient: dW[m, c] += sum_{b,t} dout[b, m, t] * pre(x[b, c, t + off])
// 64 x 64 tile of dW per CTA, reduction over a (batch, time-chunk) split, fp32 atomics.
// ------------------------------------------------------------------------------------------
constexpr int WG_T = 16;      // frames per smem slab
constexpr int WG_CHUNK = 2048;  // frames reduced by one CTA

struct WgradParams {
  int B, T_out, M, nchunk;
  SrcDev src;
  const void* dout;
  float* dw;
};

template <typename T>
__global__ void __launch_bounds__(256) taps_wgrad_kernel(const WgradParams p) {
  __shared__ __align__(16) float Ds[WG_T][64 + 4];
  __shared__ __align__(16) float Xs[WG_T][64 + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int c0 = blockIdx.x * 64, m0 = blockIdx.y * 64;
  const int b = blockIdx.z / p.nchunk, chunk = blockIdx.z % p.nchunk;
  const int t_begin = chunk * WG_CHUNK;
  const int t_end = min(p.T_out, t_begin + WG_CHUNK);
  const int l_row = tid >> 2, l_t = (tid & 3) * 4;

  const T* dout = reinterpret_cast<const T*>(p.dout) + ((long long)b * p.M + (m0 + l_row)) * p.T_out;
  const T* x = reinterpret_cast<const T*>(p.src.x) + (long long)b * p.src.bs + (long long)(c0 + l_row) * p.src.cs;
  const bool m_ok = (m0 + l_row) < p.M, c_ok = (c0 + l_row) < p.src.C;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int ts = t_begin; ts < t_end; ts += WG_T) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = ts + l_t + i;
      float dv = 0.f, xv = 0.f;
      if (t < t_end) {
        if (m_ok) dv = to_f32<T>(dout[t]);
        const int tsrc = t + p.src.t_off;
        if (c_ok && tsrc >= 0 && tsrc < p.src.T_src) {
          xv = to_f32<T>(x[tsrc]);
          if (p.src.pre_act) xv = pre_transform(p.src, xv, b, c0 + l_row, tsrc);
        }
      }
      Ds[l_t + i][l_row] = dv;
      Xs[l_t + i][l_row] = xv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < WG_T; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&Ds[kk][ty * 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c < p.src.C) atomicAdd(&p.dw[(long long)m * p.src.C + c], acc[i][j]);
    }
  }
}

// out[c] += sum_{b,t} a[b,c,t] * (bb ? bb[b,c,t] : 1)
template <typename T>
__global__ void __launch_bounds__(256) channel_reduce_kernel(int B, int C, int Tn, const T* a, const T* bb,
                                                             float* out) {
  const int c = blockIdx.x, b = blockIdx.y;
  const long long base = ((long long)b * C + c) * Tn;
  float s = 0.f;
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
    float v = to_f32<T>(a[base + t]);
    if (bb) v *= to_f32<T>(bb[base + t]);
    s += v;
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) atomicAdd(&out[c], s);
  }
}

static int fill_src(SrcDev& d, const wnb200_src_t& s, const wnb200_ln_t* ln) {
  d.x = s.x; d.w = s.w; d.bs = s.batch_stride; d.cs = s.chan_stride;
  d.C = s.C; d.T_src = s.T_src; d.t_off = s.t_off; d.pre_act = s.pre_act;
  d.ln_stats = d.ln_gamma = d.ln_beta = nullptr;
  WNB_CHECK_ARG(s.pre_act >= WNB200_PRE_NONE && s.pre_act <= WNB200_PRE_LNRELU, "taps: bad pre_act %d", s.pre_act);
  if (s.pre_act == WNB200_PRE_LNRELU) {
    WNB_CHECK_ARG(ln && ln->stats && ln->gamma && ln->beta, "taps: a source with PRE_LNRELU needs its wnb200_ln_t");
    d.ln_stats = ln->stats; d.ln_gamma = ln->gamma; d.ln_beta = ln->beta;
  }
  return 0;
}

template <typename T>
static int launch_taps(const TapsParams& p, int epilogue, cudaStream_t st) {
  dim3 grid(ceil_div(p.T_out, BN), ceil_div(p.rows, BM), p.B);
  bool ln = false;
  for (int s = 0; s < p.nsrc; ++s) ln = ln || p.src[s].pre_act == WNB200_PRE_LNRELU;
  if (ln && epilogue != WNB200_EPI_NONE && epilogue != WNB200_EPI_MU) {
    set_error("taps_fwd: PRE_LNRELU sources go with EPI_NONE or EPI_MU");
    return 1;
  }
  switch (epilogue) {
    case WNB200_EPI_NONE:
      if (ln) taps_fwd_kernel<T, WNB200_EPI_NONE, true><<<grid, NT, 0, st>>>(p);
      else taps_fwd_kernel<T, WNB200_EPI_NONE, false><<<grid, NT, 0, st>>>(p);
      break;
    case WNB200_EPI_LEAKY: taps_fwd_kernel<T, WNB200_EPI_LEAKY, false><<<grid, NT, 0, st>>>(p); break;
    case WNB200_EPI_GATE: taps_fwd_kernel<T, WNB200_EPI_GATE, false><<<grid, NT, 0, st>>>(p); break;
    case WNB200_EPI_MU:
      if (ln) taps_fwd_kernel<T, WNB200_EPI_MU, true><<<grid, NT, 0, st>>>(p);
      else taps_fwd_kernel<T, WNB200_EPI_MU, false><<<grid, NT, 0, st>>>(p);
      break;
    default: set_error("taps_fwd: bad epilogue %d", epilogue); return 1;
  }
  WNB_LAUNCH_OK();
  return 0;
}

}  // namespace wnb

using namespace wnb;

extern "C" int wnb200_taps_fwd_ex(const wnb200_taps_t* a, void* stream) {
  WNB_CHECK_ARG(a != nullptr, "taps_fwd_ex: null args");
  WNB_CHECK_STRUCT(a, wnb200_taps_t, "taps_fwd_ex");
  const int dtype = a->dtype, B = a->B, T_out = a->T_out, M = a->M, nsrc = a->nsrc, epilogue = a->epilogue;
  WNB_CHECK_ARG(dtype == WNB200_F32 || dtype == WNB200_BF16, "taps_fwd: bad dtype %d", dtype);
  WNB_CHECK_ARG(nsrc >= 1 && nsrc <= WNB200_MAX_SRC, "taps_fwd: nsrc %d out of range", nsrc);
  WNB_CHECK_ARG(B >= 0 && T_out >= 0 && M >= 1, "taps_fwd: bad shape B=%d T=%d M=%d", B, T_out, M);
  if (B == 0 || T_out == 0) return 0;
  WNB_CHECK_ARG(a->out != nullptr && a->srcs != nullptr, "taps_fwd: null pointer");
  WNB_CHECK_ARG(B <= 65535, "taps_fwd: batch %d > 65535", B);
  TapsParams p;
  p.B = B; p.T_out = T_out; p.M = M; p.nsrc = nsrc;
  p.rows = (epilogue == WNB200_EPI_GATE) ? 2 * ceil_div(M, 64) * 64
         : (epilogue == WNB200_EPI_MU) ? 4 * ceil_div(M, 32) * 32 : M;
  for (int s = 0; s < nsrc; ++s) {
    WNB_CHECK_ARG(a->srcs[s].x && a->srcs[s].w && a->srcs[s].C >= 1, "taps_fwd: source %d invalid", s);
    if (fill_src(p.src[s], a->srcs[s], a->ln ? a->ln + s : nullptr)) return 1;
  }
  p.bias = a->bias; p.accumulate = a->accumulate; p.out = a->out; p.th = a->th; p.sg = a->sg;
  p.residual = a->residual; p.mu_h = a->mu_h; p.mu_h_ln = a->mu_h_ln;
  if (epilogue == WNB200_EPI_MU) {
    WNB_CHECK_ARG(a->mu_h != nullptr, "taps_fwd: EPI_MU needs mu_h");
    WNB_CHECK_ARG(!a->mu_h_ln || (a->srcs[0].pre_act == WNB200_PRE_LNRELU && a->srcs[0].C == M &&
                                  a->srcs[0].T_src == T_out && a->mu_h == a->srcs[0].x),
                  "taps_fwd: mu_h_ln wants mu_h == source 0 (C == M, T_src == T_out) with PRE_LNRELU");
  }
  WNB_CHECK_ARG(!(a->residual && epilogue == WNB200_EPI_GATE), "taps_fwd: residual is not available with EPI_GATE");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dtype == WNB200_F32 ? launch_taps<float>(p, epilogue, st) : launch_taps<__nv_bfloat16>(p, epilogue, st);
}

extern "C" int wnb200_taps_fwd(int dtype, int B, int T_out, int M, int nsrc, const wnb200_src_t* srcs,
                               const float* bias, int epilogue, int accumulate, void* out, void* th,
                               void* sg, void* stream) {
  wnb200_taps_t a;
  memset(&a, 0, sizeof(a));
  a.struct_size = sizeof(a);
  a.dtype = dtype; a.B = B; a.T_out = T_out; a.M = M; a.nsrc = nsrc; a.epilogue = epilogue; a.accumulate = accumulate;
  a.srcs = srcs; a.bias = bias; a.out = out; a.th = th; a.sg = sg;
  WNB_CHECK_ARG(epilogue != WNB200_EPI_MU, "taps_fwd: EPI_MU goes through wnb200_taps_fwd_ex");
  return wnb200_taps_fwd_ex(&a, stream);
}

extern "C" int wnb200_taps_wgrad_ex(int dtype, int B, int T_out, int M, const wnb200_src_t* src, const wnb200_ln_t* ln,
                                    const void* dout, float* dw, void* stream) {
  WNB_CHECK_ARG(dtype == WNB200_F32 || dtype == WNB200_BF16, "taps_wgrad: bad dtype %d", dtype);
  if (B == 0 || T_out == 0) return 0;
  WNB_CHECK_ARG(src && src->x && dout && dw, "taps_wgrad: null pointer");
  WgradParams p;
  p.B = B; p.T_out = T_out; p.M = M; p.nchunk = ceil_div(T_out, WG_CHUNK);
  if (fill_src(p.src, *src, ln)) return 1;
  p.dout = dout; p.dw = dw;
  WNB_CHECK_ARG((long long)B * p.nchunk <= 65535, "taps_wgrad: too many splits");
  dim3 grid(ceil_div(src->C, 64), ceil_div(M, 64), B * p.nchunk);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == WNB200_F32) taps_wgrad_kernel<float><<<grid, 256, 0, st>>>(p);
  else taps_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_taps_wgrad(int dtype, int B, int T_out, int M, const wnb200_src_t* src, const void* dout,
                                 float* dw, void* stream) {
  return wnb200_taps_wgrad_ex(dtype, B, T_out, M, src, nullptr, dout, dw, stream);
}

extern "C" int wnb200_channel_reduce(int dtype, int B, int C, int T, const void* a, const void* b_or_null,
                                     float* out, void* stream) {
  WNB_CHECK_ARG(dtype == WNB200_F32 || dtype == WNB200_BF16, "channel_reduce: bad dtype %d", dtype);
  if (B == 0 || T == 0 || C == 0) return 0;
  WNB_CHECK_ARG(a && out, "channel_reduce: null pointer");
  WNB_CHECK_ARG(B <= 65535, "channel_reduce: batch too large");
  dim3 grid(C, B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == WNB200_F32)
    channel_reduce_kernel<float><<<grid, 256, 0, st>>>(B, C, T, (const float*)a, (const float*)b_or_null, out);
  else
    channel_reduce_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(B, C, T, (const __nv_bfloat16*)a,
                                                               (const __nv_bfloat16*)b_or_null, out);
  WNB_LAUNCH_OK();
  return 0;
}
