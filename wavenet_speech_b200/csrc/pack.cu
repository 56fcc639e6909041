// Weight packing for the tensor-core kernels, straight from the reference's parameter tensors (state_dict layout):
// one launch per residual block (+ its skip bottleneck) / per output head builds every operand matrix the forward and
// backward kernels stream by TMA -- tap-major gate weights in the pipelined row order, [Wres | Wproj], the folded
// skip -> bottleneck product Wbn * Wskip (wavenet.py:100 applies the bottleneck straight to conv1x1_skip's output), the
// fused biases, and the transposed matrices of the two data-gradient contractions.  A host that is not Python binds
// these instead of re-deriving the layouts (they were torch.cat / bmm / permute glue in round 1).
//
// Also here: the weight-space algebra that turns M = dskips (x) gate into the gradients of the fold's two factors
// (a small batched fp32 GEMM pair per layer, all layers in one launch) -- the train step calls no library GEMM.
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace wnb {
typedef __nv_bfloat16 bf16;

template <typename T>
__device__ __forceinline__ float ldw(const void* p, long long i) {
  return to_f32<T>(reinterpret_cast<const T*>(p)[i]);
}
// 2-byte operand word in the activation format
__device__ __forceinline__ uint16_t to_fmt(bool f16, float v) {
  if (f16) {
    const __half h = __float2half_rn(v);
    return *reinterpret_cast<const uint16_t*>(&h);
  }
  const bf16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<const uint16_t*>(&h);
}

struct PackBlockDev {
  int C, k, f16, row_order;
  const void *wt, *bt, *ws, *bs, *wres, *bres, *wskip, *bskip, *wproj, *bproj, *wbn, *bbn;
  uint16_t* w1; float* b1; uint16_t* w2; float* b2;
  uint16_t *wdg, *wdg_skip, *wdx, *wdx_taps;
};

// fold[m][n] = sum_j Wbn[m][j] * Wskip[j][n], always summed in the same order (the forward and backward packs must agree)
template <typename T>
__device__ __forceinline__ float fold_mn(const PackBlockDev& p, int m, int n) {
  float s = 0.f;
  const int C = p.C;
  for (int j = 0; j < C; ++j) s = fmaf(ldw<T>(p.wbn, (long long)m * C + j), ldw<T>(p.wskip, (long long)j * C + n), s);
  return s;
}

// source row of w1 row r: (gate 0 = tanh / 1 = sigmoid, output channel m)
__device__ __forceinline__ void w1_row(int row_order, int C, int r, int& gate, int& m) {
  if (row_order == 0) { gate = r / C; m = r - gate * C; return; }
  const int hc = C / 2, blk = r / hc;          // [tanh 0:hc ; sig 0:hc ; tanh hc:C ; sig hc:C]
  gate = blk & 1;
  m = (blk >> 1) * hc + (r - blk * hc);
}

// blockIdx.y selects the segment; blockIdx.x / threadIdx.x stride over its elements (consecutive threads -> consecutive
// output columns: coalesced stores; the transposed reads of the backward packs go through L1/L2, 1 MB per block in all)
template <typename T>
__global__ void __launch_bounds__(256) pack_block_kernel(const PackBlockDev p) {
  const int C = p.C, k = p.k;
  const bool f16 = p.f16 != 0;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  switch (blockIdx.y) {
    case 0: {   // w1 [2C][k*C], tap-major columns
      const long long n = 2LL * C * k * C;
      for (long long i = tid; i < n; i += nth) {
        const int r = (int)(i / (k * C)), col = (int)(i - (long long)r * k * C);
        const int j = col / C, c = col - j * C;
        int gate, m;
        w1_row(p.row_order, C, r, gate, m);
        p.w1[i] = to_fmt(f16, ldw<T>(gate ? p.ws : p.wt, ((long long)m * C + c) * k + j));
      }
      break;
    }
    case 1: {   // b1 [2C] (pre-scaled for the ex2-based gate of the fp16 format), b2 [2C]
      for (long long i = tid; i < 2LL * C; i += nth) {
        int gate, m;
        w1_row(p.row_order, C, (int)i, gate, m);
        float v = ldw<T>(gate ? p.bs : p.bt, m);
        if (f16) v *= gate ? -1.4426950408889634f : -2.885390081777927f;
        p.b1[i] = v;
        if (i < C) {
          p.b2[i] = ldw<T>(p.bres, i) + ldw<T>(p.bproj, i);
        } else {
          const int mm = (int)i - C;
          float s = 0.f;
          for (int j = 0; j < C; ++j) s = fmaf(ldw<T>(p.wbn, (long long)mm * C + j), ldw<T>(p.bskip, j), s);
          p.b2[i] = s + ldw<T>(p.bbn, mm);
        }
      }
      break;
    }
    case 2: {   // w2 [2C][2C] = [[Wres, Wproj], [fold, 0]]
      const long long n = 4LL * C * C;
      for (long long i = tid; i < n; i += nth) {
        const int r = (int)(i / (2 * C)), col = (int)(i - (long long)r * 2 * C);
        float v;
        if (r < C) v = col < C ? ldw<T>(p.wres, (long long)r * C + col) : ldw<T>(p.wproj, (long long)r * C + col - C);
        else v = col < C ? fold_mn<T>(p, r - C, col) : 0.f;
        p.w2[i] = to_fmt(f16, v);
      }
      break;
    }
    case 3: {   // wdg [C][2C] = [Wres^T | fold^T], wdg_skip [C][C] = fold^T        (backward, bf16)
      if (!p.wdg) break;
      const long long n = 2LL * C * C;
      for (long long i = tid; i < n; i += nth) {
        const int r = (int)(i / (2 * C)), col = (int)(i - (long long)r * 2 * C);
        float v;
        if (col < C) v = ldw<T>(p.wres, (long long)col * C + r);
        else {
          v = fold_mn<T>(p, col - C, r);
          p.wdg_skip[(long long)r * C + col - C] = to_fmt(false, v);
        }
        p.wdg[i] = to_fmt(false, v);
      }
      break;
    }
    default: {  // wdx [C][k*2C + C] = [Wt_0^T | Ws_0^T | ... | Wproj^T], wdx_taps = the same without Wproj^T
      if (!p.wdx) break;
      const int W = k * 2 * C + C;
      const long long n = (long long)C * W;
      for (long long i = tid; i < n; i += nth) {
        const int r = (int)(i / W), col = (int)(i - (long long)r * W);
        float v;
        if (col < k * 2 * C) {
          const int j = col / (2 * C), q = col - j * 2 * C;
          v = q < C ? ldw<T>(p.wt, ((long long)q * C + r) * k + j) : ldw<T>(p.ws, ((long long)(q - C) * C + r) * k + j);
          p.wdx_taps[(long long)r * (k * 2 * C) + col] = to_fmt(false, v);
        } else {
          v = ldw<T>(p.wproj, (long long)(col - k * 2 * C) * C + r);
        }
        p.wdx[i] = to_fmt(false, v);
      }
      break;
    }
  }
}

struct PackHeadDev {
  int C, n_out, n2, npad, f16;
  const void *w1, *b1, *w3, *b3;
  uint16_t* pw1; float* pb1; uint16_t* pw2; float* pb2; uint16_t* w3t; uint16_t* w1t;
};

template <typename T>
__global__ void __launch_bounds__(256) pack_head_kernel(const PackHeadDev p) {
  const int C = p.C;
  const bool f16 = p.f16 != 0;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < (long long)C * C; i += nth) {
    const int r = (int)(i / C), c = (int)(i - (long long)r * C);
    p.pw1[i] = to_fmt(f16, ldw<T>(p.w1, i));
    if (p.w1t) p.w1t[i] = to_fmt(false, ldw<T>(p.w1, (long long)c * C + r));
  }
  for (long long i = tid; i < (long long)p.n2 * C; i += nth) {
    const int r = (int)(i / C);
    p.pw2[i] = to_fmt(f16, r < p.n_out ? ldw<T>(p.w3, i) : 0.f);
  }
  if (p.w3t)
    for (long long i = tid; i < (long long)C * p.npad; i += nth) {
      const int r = (int)(i / p.npad), c = (int)(i - (long long)r * p.npad);
      p.w3t[i] = to_fmt(false, c < p.n_out ? ldw<T>(p.w3, (long long)c * C + r) : 0.f);
    }
  for (long long i = tid; i < C; i += nth) p.pb1[i] = ldw<T>(p.b1, i);
  for (long long i = tid; i < p.n2; i += nth) p.pb2[i] = i < p.n_out ? ldw<T>(p.b3, i) : 0.f;
}

// ---------------------------------------------------------------------------------------------------------------
// Fold gradients, all layers in one launch.  Per layer l (C x C matrices, fp32 accumulation):
//   dwskip = Wbn^T M        dwbn = M Wskip^T + csk (x) bskip        dbskip = Wbn^T csk
// 64 x 64 output tiles, 16-deep K slices through shared memory, 4 x 4 outputs per thread.
constexpr int FG_MAXL = 64;
struct FoldGradDev {
  int L, C;
  const void* wbn[FG_MAXL];
  const void* wskip[FG_MAXL];
  const void* bskip[FG_MAXL];
  const float* M;       // [L][C][C]
  const float* csk;     // [C]
  float* dwskip;        // [L][C][C]
  float* dwbn;          // [L][C][C]
  float* dbskip;        // [L][C]
};

template <typename T>
__global__ void __launch_bounds__(256) fold_grads_kernel(const __grid_constant__ FoldGradDev p) {
  const int C = p.C, l = blockIdx.z >> 1, which = blockIdx.z & 1;
  const int tiles = C / 64;
  const int ti = blockIdx.x / tiles, tj = blockIdx.x - ti * tiles;
  const int i0 = ti * 64, j0 = tj * 64;
  __shared__ float As[16][64 + 4];     // [kk][i]
  __shared__ float Bs[16][64 + 4];     // [kk][j]
  const float* M = p.M + (long long)l * C * C;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < C; k0 += 16) {
    // which 0: out[i][j] = sum_k Wbn[k][i] * M[k][j]          A(i,k) = Wbn[k][i],  B(k,j) = M[k][j]
    // which 1: out[i][j] = sum_k M[i][k] * Wskip[j][k]        A(i,k) = M[i][k],    B(k,j) = Wskip[j][k]
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      if (which == 0) {
        const int kk = e >> 6, ii = e & 63;                      // consecutive threads -> consecutive i / j: coalesced
        As[kk][ii] = ldw<T>(p.wbn[l], (long long)(k0 + kk) * C + i0 + ii);
        Bs[kk][ii] = M[(long long)(k0 + kk) * C + j0 + ii];
      } else {
        const int ii = e >> 4, kk = e & 15;                      // consecutive threads -> consecutive k
        As[kk][ii] = M[(long long)(i0 + ii) * C + k0 + kk];
        Bs[kk][ii] = ldw<T>(p.wskip[l], (long long)(j0 + ii) * C + k0 + kk);
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = As[kk][ty * 4 + u]; b[u] = Bs[kk][tx * 4 + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
  float* out = (which == 0 ? p.dwskip : p.dwbn) + (long long)l * C * C;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + ty * 4 + u;
    float4 v = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
    if (which == 1) {                                            // + csk[i] * bskip[j]
      const float ci = p.csk[i];
      const int j = j0 + tx * 4;
      v.x = fmaf(ci, ldw<T>(p.bskip[l], j), v.x);
      v.y = fmaf(ci, ldw<T>(p.bskip[l], j + 1), v.y);
      v.z = fmaf(ci, ldw<T>(p.bskip[l], j + 2), v.z);
      v.w = fmaf(ci, ldw<T>(p.bskip[l], j + 3), v.w);
    }
    *reinterpret_cast<float4*>(out + (long long)i * C + j0 + tx * 4) = v;
  }
  // dbskip[i] = sum_k Wbn[k][i] csk[k]: by the first tile column of the `which == 0` grid
  if (which == 0 && tj == 0 && threadIdx.x < 64) {
    const int i = i0 + threadIdx.x;
    float s = 0.f;
    for (int kx = 0; kx < C; ++kx) s = fmaf(ldw<T>(p.wbn[l], (long long)kx * C + i), p.csk[kx], s);
    p.dbskip[(long long)l * C + i] = s;
  }
}

}  // namespace wnb

using namespace wnb;

extern "C" int wnb200_pack_block(const wnb200_pack_block_t* a, void* stream) {
  WNB_CHECK_ARG(a != nullptr, "pack_block: null argument");
  WNB_CHECK_STRUCT(a, wnb200_pack_block_t, "pack_block");
  WNB_CHECK_ARG(a->C >= 64 && a->C % 64 == 0 && a->C <= 1024, "pack_block: C=%d must be a multiple of 64", a->C);
  WNB_CHECK_ARG(a->k >= 1 && a->k <= 3, "pack_block: k=%d not in 1..3", a->k);
  WNB_CHECK_ARG(a->act_fmt == WNB200_ACT_BF16 || a->act_fmt == WNB200_ACT_F16X2, "pack_block: bad act_fmt");
  WNB_CHECK_ARG(a->w_dtype == WNB200_F32 || a->w_dtype == WNB200_BF16, "pack_block: bad w_dtype");
  WNB_CHECK_ARG(a->row_order == 0 || a->row_order == 1, "pack_block: bad row_order");
  WNB_CHECK_ARG(a->wt && a->bt && a->ws && a->bs && a->wres && a->bres && a->wskip && a->bskip && a->wproj &&
                    a->bproj && a->wbn && a->bbn, "pack_block: null parameter pointer");
  WNB_CHECK_ARG(a->w1 && a->b1 && a->w2 && a->b2, "pack_block: null output pointer");
  const bool bwd = a->wdg || a->wdg_skip || a->wdx || a->wdx_taps;
  WNB_CHECK_ARG(!bwd || (a->wdg && a->wdg_skip && a->wdx && a->wdx_taps), "pack_block: backward packs come as all four");
  PackBlockDev p;
  memset(&p, 0, sizeof(p));
  p.C = a->C; p.k = a->k; p.f16 = a->act_fmt == WNB200_ACT_F16X2; p.row_order = a->row_order;
  p.wt = a->wt; p.bt = a->bt; p.ws = a->ws; p.bs = a->bs; p.wres = a->wres; p.bres = a->bres; p.wskip = a->wskip;
  p.bskip = a->bskip; p.wproj = a->wproj; p.bproj = a->bproj; p.wbn = a->wbn; p.bbn = a->bbn;
  p.w1 = (uint16_t*)a->w1; p.b1 = a->b1; p.w2 = (uint16_t*)a->w2; p.b2 = a->b2;
  p.wdg = (uint16_t*)a->wdg; p.wdg_skip = (uint16_t*)a->wdg_skip; p.wdx = (uint16_t*)a->wdx;
  p.wdx_taps = (uint16_t*)a->wdx_taps;
  dim3 grid(64, bwd ? 5 : 3);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->w_dtype == WNB200_F32) pack_block_kernel<float><<<grid, 256, 0, st>>>(p);
  else pack_block_kernel<bf16><<<grid, 256, 0, st>>>(p);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_pack_head(const wnb200_pack_head_t* a, void* stream) {
  WNB_CHECK_ARG(a != nullptr, "pack_head: null argument");
  WNB_CHECK_STRUCT(a, wnb200_pack_head_t, "pack_head");
  WNB_CHECK_ARG(a->C >= 64 && a->C % 64 == 0 && a->n_out >= 1 && a->n_out <= 256, "pack_head: bad sizes");
  WNB_CHECK_ARG(a->act_fmt == WNB200_ACT_BF16 || a->act_fmt == WNB200_ACT_F16X2, "pack_head: bad act_fmt");
  WNB_CHECK_ARG(a->w_dtype == WNB200_F32 || a->w_dtype == WNB200_BF16, "pack_head: bad w_dtype");
  WNB_CHECK_ARG(a->w1 && a->b1 && a->w3 && a->b3 && a->pw1 && a->pb1 && a->pw2 && a->pb2, "pack_head: null pointer");
  PackHeadDev p;
  memset(&p, 0, sizeof(p));
  p.C = a->C; p.n_out = a->n_out; p.n2 = (a->n_out + 15) / 16 * 16; p.npad = (a->n_out + 63) / 64 * 64;
  p.f16 = a->act_fmt == WNB200_ACT_F16X2;
  p.w1 = a->w1; p.b1 = a->b1; p.w3 = a->w3; p.b3 = a->b3;
  p.pw1 = (uint16_t*)a->pw1; p.pb1 = a->pb1; p.pw2 = (uint16_t*)a->pw2; p.pb2 = a->pb2;
  p.w3t = (uint16_t*)a->w3t; p.w1t = (uint16_t*)a->w1t;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->w_dtype == WNB200_F32) pack_head_kernel<float><<<64, 256, 0, st>>>(p);
  else pack_head_kernel<bf16><<<64, 256, 0, st>>>(p);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_fold_grads(int w_dtype, int L, int C, const void* const* wbn, const void* const* wskip,
                                 const void* const* bskip, const float* M, const float* csk, float* dwskip, float* dwbn,
                                 float* dbskip, void* stream) {
  WNB_CHECK_ARG(w_dtype == WNB200_F32 || w_dtype == WNB200_BF16, "fold_grads: bad w_dtype");
  WNB_CHECK_ARG(L >= 0 && L <= FG_MAXL, "fold_grads: L=%d layers per call, at most %d", L, FG_MAXL);
  WNB_CHECK_ARG(C >= 64 && C % 64 == 0, "fold_grads: C=%d must be a multiple of 64", C);
  if (L == 0) return 0;
  WNB_CHECK_ARG(wbn && wskip && bskip && M && csk && dwskip && dwbn && dbskip, "fold_grads: null pointer");
  FoldGradDev p;
  memset(&p, 0, sizeof(p));
  p.L = L; p.C = C;
  for (int l = 0; l < L; ++l) {
    WNB_CHECK_ARG(wbn[l] && wskip[l] && bskip[l], "fold_grads: null parameter pointer (layer %d)", l);
    p.wbn[l] = wbn[l]; p.wskip[l] = wskip[l]; p.bskip[l] = bskip[l];
  }
  p.M = M; p.csk = csk; p.dwskip = dwskip; p.dwbn = dwbn; p.dbskip = dbskip;
  dim3 grid((C / 64) * (C / 64), 1, 2 * L);
  cudaStream_t st = (cudaStream_t)stream;
  if (w_dtype == WNB200_F32) fold_grads_kernel<float><<<grid, 256, 0, st>>>(p);
  else fold_grads_kernel<bf16><<<grid, 256, 0, st>>>(p);
  WNB_LAUNCH_OK();
  return 0;
}

// Bytes of the device buffers one forward of a residual stack needs besides its input and output, for a host that
// allocates them itself (the Python side lets torch's allocator do it): the NLC stream (x2 when it is an fp16 (hi, lo)
// pair, x2 again for the ping-pong), the fp32 running skip sum, the head's two NLC activations.
extern "C" size_t wnb200_workspace_bytes(int B, int T, int C, int act_fmt) {
  if (B <= 0 || T <= 0 || C <= 0) return 0;
  const size_t frames = (size_t)B * (size_t)T;
  const size_t stream = frames * C * 2 * 2 * (act_fmt == WNB200_ACT_F16X2 ? 2 : 1);
  const size_t skips = frames * C * 4;
  const size_t head = frames * C * 2 * 2;
  return stream + skips + head + 4096;
}
