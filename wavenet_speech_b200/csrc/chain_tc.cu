// Tensor-core path of the residual stack: one persistent, warp-specialised kernel that evaluates a chain of
// up to two channel contractions per 128-frame tile with the intermediate kept on chip.
//
//   stage 1:  acc1[128 x n1] = sum over taps j, channel blocks:  X(t + off_j)[128 x 64] * W1[n1 x 64]^T
//   epi  1:   GATE   act = tanh(acc1[:, 0:C] + b) * sigmoid(acc1[:, C:2C] + b)        (block.py:185)
//             LEAKY  act = LeakyReLU(acc1 + b)           LINEAR  act = acc1 + b
//             -> bf16, written to shared memory in the swizzled K-major layout stage 2 reads
//   stage 2:  acc2[128 x n2] = act * W2a^T  (+ X(t) * W2x^T into columns [0, C))
//   epi  2:   RESBLOCK  res = acc2[:, 0:C] + b -> bf16 NLC;  skips (+)= acc2[:, C:2C] + b (fp32 NLC)
//             HEAD      (softmax over) acc2[:, 0:n_out] + b -> NCL
//
// Data movement: activations are NLC bf16; a 3-D TMA tensor map (C, T, B) loads [128 frames x 64 channels]
// boxes at frame coordinate t0 + off_j -- out-of-range frames are zero-filled by the TMA unit, which IS the
// reference's zero padding (conv_ops.py:31-34,65-68), and the batch coordinate stops any bleed between reads.
// Weights stream as [rows x 64] K-major boxes.  Both land in the 128B-swizzled layout tcgen05.mma consumes.
// Accumulators live in TMEM (128 lanes x up to 512 fp32 columns); one elected thread issues tcgen05.mma;
// four epilogue warps read TMEM with tcgen05.ld (one frame per thread, channels along columns, so the gate
// and the channel softmax are thread-local).
//
// Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

#ifdef WNB200_TIMELINE
#define WNB_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && it < 8) p.dbg[it * 16 + (slot)] = clock64(); } while (0)
#else
#define WNB_STAMP(slot) do { } while (0)
#endif

namespace wnb {
using namespace tc;
typedef __nv_bfloat16 bf16;

enum { EPI1_GATE = 0, EPI1_LEAKY = 1, EPI1_LINEAR = 2 };
enum { EPI2_NONE = 0, EPI2_RESBLOCK = 1, EPI2_HEAD = 2 };

struct ChainDev {
  int B, T, tiles_per_seq, num_tiles;
  int ntaps, t_off[3];
  int epi1, n1;
  int n2, use_x2, epi2;
  int w1_boxrows, w2a_boxrows, w2x_boxrows;
  const float* bias1;
  const float* bias2;
  bf16* y_nlc;       // stage-1-only output, or res
  float* skips;      // fp32 NLC running skip sum
  int skips_init;    // 1: write, 0: accumulate
  bf16* skips_act;   // optional: LeakyReLU(skips) as bf16 NLC (input of the head)
  void* out_ncl;     // HEAD output, NCL
  int out_f32, n_out, softmax;
  long long* dbg;   // optional timeline buffer (block 0): [iter][16] clock64 stamps
};

constexpr int TILE_M = 128;
constexpr int A_BYTES = TILE_M * 128;  // one [128 x 64] bf16 block
constexpr int NUM_THREADS = 192;

template <int C>
struct Cfg {
  static constexpr int KB = C / 64;                     // 64-channel blocks per tensor
  static constexpr int B_ROWS = 2 * C;                  // max weight rows per stage
  static constexpr int STAGE_BYTES = A_BYTES + B_ROWS * 128;
  static constexpr int ACT_BYTES = KB * A_BYTES;
  static constexpr int NSTAGE = (C == 256) ? 2 : (C == 128 ? 3 : 4);
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + ACT_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void issue_mma_block(uint32_t tmem_base, uint32_t a_addr, uint32_t b_addr, int n,
                                                bool first) {
  // one 64-channel K block = 4 UMMA K-steps of 16; N split into <=256-column instructions
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4) {
    const uint64_t adesc = make_smem_desc_sw128(a_addr + k4 * 32);
    for (int n0 = 0; n0 < n; n0 += 256) {
      const int nn = (n - n0) < 256 ? (n - n0) : 256;
      const uint64_t bdesc = make_smem_desc_sw128(b_addr + n0 * 128 + k4 * 32);
      umma_bf16(tmem_base + n0, adesc, bdesc, make_idesc_bf16(TILE_M, nn), (first && k4 == 0) ? 0u : 1u);
    }
  }
}

template <int C>
__global__ void __launch_bounds__(NUM_THREADS, 1)
chain_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
             const __grid_constant__ CUtensorMap map_w2a, const __grid_constant__ CUtensorMap map_w2x,
             const ChainDev p) {
  using K = Cfg<C>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base;
  const uint32_t act_base = smem_base + K::NSTAGE * K::STAGE_BYTES;
  const uint32_t bar_base = act_base + K::ACT_BYTES;
  // barriers (8 B each)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (K::NSTAGE + s); };
  const uint32_t g1_full = bar_base + 8u * (2 * K::NSTAGE);
  const uint32_t act_ready = g1_full + 8;
  const uint32_t g2_full = g1_full + 16;
  const uint32_t tmem_empty = g1_full + 24;
  const uint32_t tmem_slot = g1_full + 32;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));   // generic pointer to the aligned base

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has2 = p.n2 > 0;
  int tmem_cols = 32;
  {
    const int need = p.n1 > p.n2 ? p.n1 : p.n2;
    while (tmem_cols < need) tmem_cols <<= 1;
  }

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_w1);
    if (has2) prefetch_tensormap(&map_w2a);
    if (has2 && p.use_x2) prefetch_tensormap(&map_w2x);
    for (int s = 0; s < K::NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(g1_full, 1);
    mbar_init(act_ready, 128);
    mbar_init(g2_full, 1);
    mbar_init(tmem_empty, 128);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == K::NSTAGE) { stage = 0; phase ^= 1; }
      };
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int b = tile / p.tiles_per_seq;
        const int t0 = (tile - b * p.tiles_per_seq) * TILE_M;
        for (int kb = 0; kb < p.ntaps * K::KB; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = stage_base + stage * K::STAGE_BYTES, sb = sa + A_BYTES;
          mbar_expect_tx(full_bar(stage), A_BYTES + p.n1 * 128);
          const int tap = kb / K::KB, cb = kb - tap * K::KB;
          tma_load_3d(sa, &map_x, full_bar(stage), cb * 64, t0 + p.t_off[tap], b);
          for (int r = 0; r < p.n1; r += p.w1_boxrows)
            tma_load_2d(sb + r * 128, &map_w1, full_bar(stage), kb * 64, r);
          advance();
        }
        if (has2) {
          for (int kb = 0; kb < K::KB; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t sb = stage_base + stage * K::STAGE_BYTES + A_BYTES;
            mbar_expect_tx(full_bar(stage), p.n2 * 128);
            for (int r = 0; r < p.n2; r += p.w2a_boxrows)
              tma_load_2d(sb + r * 128, &map_w2a, full_bar(stage), kb * 64, r);
            advance();
          }
          if (p.use_x2) {
            for (int kb = 0; kb < K::KB; ++kb) {
              mbar_wait(empty_bar(stage), phase ^ 1);
              const uint32_t sa = stage_base + stage * K::STAGE_BYTES, sb = sa + A_BYTES;
              mbar_expect_tx(full_bar(stage), A_BYTES + C * 128);
              tma_load_3d(sa, &map_x, full_bar(stage), kb * 64, t0, b);
              for (int r = 0; r < C; r += p.w2x_boxrows)
                tma_load_2d(sb + r * 128, &map_w2x, full_bar(stage), C + kb * 64, r);
              advance();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == K::NSTAGE) { stage = 0; phase ^= 1; }
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        if (it > 0) {
          mbar_wait(tmem_empty, (uint32_t)((it - 1) & 1));
          tc_fence_after();
        }
        WNB_STAMP(0);
        for (int kb = 0; kb < p.ntaps * K::KB; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = stage_base + stage * K::STAGE_BYTES;
          issue_mma_block(tmem_base, sa, sa + A_BYTES, p.n1, kb == 0);
          umma_commit(empty_bar(stage));
          advance();
        }
        umma_commit(g1_full);
        WNB_STAMP(1);
        if (has2) {
          mbar_wait(act_ready, (uint32_t)(it & 1));
          tc_fence_after();
          WNB_STAMP(2);
          for (int kb = 0; kb < K::KB; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sb = stage_base + stage * K::STAGE_BYTES + A_BYTES;
            issue_mma_block(tmem_base, act_base + kb * A_BYTES, sb, p.n2, kb == 0);
            umma_commit(empty_bar(stage));
            advance();
          }
          if (p.use_x2) {
            for (int kb = 0; kb < K::KB; ++kb) {
              mbar_wait(full_bar(stage), phase);
              tc_fence_after();
              const uint32_t sa = stage_base + stage * K::STAGE_BYTES;
              issue_mma_block(tmem_base, sa, sa + A_BYTES, C, false);
              umma_commit(empty_bar(stage));
              advance();
            }
          }
          umma_commit(g2_full);
          WNB_STAMP(3);
        }
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;          // frame inside the tile
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int b = tile / p.tiles_per_seq;
      const int t0 = (tile - b * p.tiles_per_seq) * TILE_M;
      const int t = t0 + row;
      const bool valid = t < p.T;
      const long long grow = ((long long)b * p.T + t) * C;   // NLC row offset (elements)

      mbar_wait(g1_full, (uint32_t)(it & 1));
      tc_fence_after();
      if (threadIdx.x == 64) WNB_STAMP(4);
      if (has2) {
        // ---- epilogue 1: -> act (bf16, swizzled K-major blocks in smem) ----
        for (int c0 = 0; c0 < C; c0 += 16) {
          float a[16], g[16];
          tmem_ld16(lane_addr + c0, a);
          if (p.epi1 == EPI1_GATE) tmem_ld16(lane_addr + C + c0, g);
          tmem_wait_ld();
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float v0, v1;
            if (p.epi1 == EPI1_GATE) {
              v0 = tanh_approx(a[i] + __ldg(p.bias1 + c0 + i)) * sigmoid_approx(g[i] + __ldg(p.bias1 + C + c0 + i));
              v1 = tanh_approx(a[i + 1] + __ldg(p.bias1 + c0 + i + 1)) *
                   sigmoid_approx(g[i + 1] + __ldg(p.bias1 + C + c0 + i + 1));
            } else {
              v0 = a[i] + __ldg(p.bias1 + c0 + i);
              v1 = a[i + 1] + __ldg(p.bias1 + c0 + i + 1);
              if (p.epi1 == EPI1_LEAKY) { v0 = leaky(v0); v1 = leaky(v1); }
            }
            pk[i >> 1] = pack_bf16x2(v0, v1);
          }
          const int kb = c0 >> 6, ci = (c0 & 63) >> 3;   // 16-byte chunk index inside the 128-byte row
          uint8_t* blk = smem_gen + (act_base - smem_base) + kb * A_BYTES + row * 128;
          *reinterpret_cast<uint4*>(blk + (((ci) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(blk + (((ci + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the tensor core (async proxy)
        tc_fence_before();
        mbar_arrive(act_ready);
        if (threadIdx.x == 64) WNB_STAMP(5);

        mbar_wait(g2_full, (uint32_t)(it & 1));
        tc_fence_after();
        if (threadIdx.x == 64) WNB_STAMP(6);
        if (p.epi2 == EPI2_RESBLOCK) {
          if (p.y_nlc) {
            for (int c0 = 0; c0 < C; c0 += 16) {
              float a[16];
              tmem_ld16(lane_addr + c0, a);
              tmem_wait_ld();
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 16; i += 2)
                pk[i >> 1] = pack_bf16x2(a[i] + __ldg(p.bias2 + c0 + i), a[i + 1] + __ldg(p.bias2 + c0 + i + 1));
              if (valid) {
                uint4* dst = reinterpret_cast<uint4*>(p.y_nlc + grow + c0);
                dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
              }
            }
          }
          for (int c0 = 0; c0 < C; c0 += 16) {
            float a[16];
            tmem_ld16(lane_addr + C + c0, a);
            tmem_wait_ld();
            if (valid) {
              float4* sp = reinterpret_cast<float4*>(p.skips + grow + c0);
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = a[i] + __ldg(p.bias2 + C + c0 + i);
              if (!p.skips_init) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 o = sp[j];
                  v[4 * j] += o.x; v[4 * j + 1] += o.y; v[4 * j + 2] += o.z; v[4 * j + 3] += o.w;
                }
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) sp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              if (p.skips_act) {
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) pk[i >> 1] = pack_bf16x2(leaky(v[i]), leaky(v[i + 1]));
                uint4* dst = reinterpret_cast<uint4*>(p.skips_act + grow + c0);
                dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
              }
            }
          }
        } else {   // EPI2_HEAD: (softmax over) n_out channels -> NCL
          float mx = -INFINITY, sum = 0.f;
          if (p.softmax) {
            for (int c0 = 0; c0 < p.n2; c0 += 16) {
              float a[16];
              tmem_ld16(lane_addr + c0, a);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c0 + i < p.n_out) mx = fmaxf(mx, a[i] + __ldg(p.bias2 + c0 + i));
            }
            for (int c0 = 0; c0 < p.n2; c0 += 16) {
              float a[16];
              tmem_ld16(lane_addr + c0, a);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c0 + i < p.n_out) sum += __expf(a[i] + __ldg(p.bias2 + c0 + i) - mx);
            }
            sum = 1.f / sum;
          }
          for (int c0 = 0; c0 < p.n2; c0 += 16) {
            float a[16];
            tmem_ld16(lane_addr + c0, a);
            tmem_wait_ld();
            if (valid) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int c = c0 + i;
                if (c < p.n_out) {
                  float v = a[i] + __ldg(p.bias2 + c);
                  if (p.softmax) v = __expf(v - mx) * sum;
                  const long long o = ((long long)b * p.n_out + c) * p.T + t;
                  if (p.out_f32) reinterpret_cast<float*>(p.out_ncl)[o] = v;
                  else reinterpret_cast<bf16*>(p.out_ncl)[o] = __float2bfloat16_rn(v);
                }
              }
            }
          }
        }
      } else {
        // ---- single-stage: epilogue 1 writes NLC bf16 straight to global ----
        for (int c0 = 0; c0 < p.n1; c0 += 16) {
          float a[16];
          tmem_ld16(lane_addr + c0, a);
          tmem_wait_ld();
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float v0 = a[i] + __ldg(p.bias1 + c0 + i), v1 = a[i + 1] + __ldg(p.bias1 + c0 + i + 1);
            if (p.epi1 == EPI1_LEAKY) { v0 = leaky(v0); v1 = leaky(v1); }
            pk[i >> 1] = pack_bf16x2(v0, v1);
          }
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(p.y_nlc + ((long long)b * p.T + t) * p.n1 + c0);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty);
      if (threadIdx.x == 64) WNB_STAMP(7);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D bf16 row-major matrix [rows][cols] -> boxes of [boxrows][64], 128B swizzle.
static int make_map_2d(CUtensorMap* m, const void* ptr, int rows, int cols, int boxrows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return 5; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)boxrows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(2d %dx%d box %d) failed: %d", rows, cols, boxrows, (int)r); return 5; }
  return 0;
}

// NLC bf16 activations [B][T][C] -> boxes of [1][128 frames][64 channels], 128B swizzle, zero OOB fill.
static int make_map_x(CUtensorMap* m, const void* ptr, int B, int T, int C) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return 5; }
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2};
  cuuint32_t box[3] = {64, TILE_M, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(x %dx%dx%d) failed: %d", B, T, C, (int)r); return 5; }
  return 0;
}

template <int C>
static int launch_chain(const CUtensorMap& mx, const CUtensorMap& mw1, const CUtensorMap& mw2a,
                        const CUtensorMap& mw2x, const ChainDev& p, cudaStream_t st) {
  using K = Cfg<C>;
  WNB_SET_SMEM_ATTR(K::SMEM_BYTES, chain_kernel<C>);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  chain_kernel<C><<<grid, NUM_THREADS, K::SMEM_BYTES, st>>>(mx, mw1, mw2a, mw2x, p);
  WNB_LAUNCH_OK();
  return 0;
}

}  // namespace wnb

using namespace wnb;

extern "C" int wnb200_chain_fwd_tc(const wnb200_chain_t* a, void* stream) {
  WNB_CHECK_ARG(a != nullptr, "chain_fwd_tc: null argument");
  WNB_CHECK_STRUCT(a, wnb200_chain_t, "chain_fwd_tc");
  const int C = a->C;
  WNB_CHECK_ARG(C == 64 || C == 128 || C == 256, "chain_fwd_tc: C=%d not in {64,128,256}", C);
  WNB_CHECK_ARG(a->ntaps >= 1 && a->ntaps <= 3, "chain_fwd_tc: ntaps=%d not in 1..3", a->ntaps);
  WNB_CHECK_ARG(a->x && a->w1 && a->bias1, "chain_fwd_tc: null stage-1 pointer");
  WNB_CHECK_ARG(a->n1 % 16 == 0 && a->n1 >= 16 && a->n1 <= 2 * C, "chain_fwd_tc: bad n1=%d", a->n1);
  WNB_CHECK_ARG(a->epi1 != EPI1_GATE || (a->n1 == 2 * C && a->n2 > 0),
                "chain_fwd_tc: GATE needs n1 = 2C and a second stage");
  WNB_CHECK_ARG(a->n2 == 0 || (a->n2 % 16 == 0 && a->n2 <= 2 * C && (a->n2 <= 256 || a->n2 % 256 == 0) && a->w2 &&
                               a->bias2),
                "chain_fwd_tc: bad stage 2 (n2=%d)", a->n2);
  WNB_CHECK_ARG(a->n1 <= 256 || a->n1 % 256 == 0, "chain_fwd_tc: n1=%d must be <= 256 or a multiple of 256", a->n1);
  WNB_CHECK_ARG(a->n2 == 0 || a->epi1 == EPI1_GATE || a->n1 == C, "chain_fwd_tc: stage 2 needs n1 == C");
  if (a->B == 0 || a->T == 0) return 0;
  ChainDev p;
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.T = a->T;
  p.tiles_per_seq = ceil_div(a->T, TILE_M);
  p.num_tiles = p.tiles_per_seq * a->B;
  p.ntaps = a->ntaps;
  for (int j = 0; j < 3; ++j) p.t_off[j] = a->t_off[j];
  p.epi1 = a->epi1; p.n1 = a->n1; p.n2 = a->n2; p.use_x2 = a->use_x2; p.epi2 = a->epi2;
  p.bias1 = a->bias1; p.bias2 = a->bias2;
  p.y_nlc = (bf16*)a->y_nlc; p.skips = a->skips; p.skips_init = a->skips_init; p.skips_act = (bf16*)a->skips_act;
  p.out_ncl = a->out_ncl; p.out_f32 = a->out_f32; p.n_out = a->n_out; p.softmax = a->softmax;
  p.dbg = (long long*)a->dbg;
  if (a->n2 > 0) {
    WNB_CHECK_ARG(a->epi2 == EPI2_RESBLOCK || a->epi2 == EPI2_HEAD, "chain_fwd_tc: bad epi2");
    if (a->epi2 == EPI2_RESBLOCK)
      WNB_CHECK_ARG(a->n2 == 2 * C && a->skips, "chain_fwd_tc: RESBLOCK needs n2 = 2C and a skips buffer");
    if (a->epi2 == EPI2_HEAD)
      WNB_CHECK_ARG(a->out_ncl && a->n_out >= 1 && a->n_out <= a->n2, "chain_fwd_tc: HEAD needs out_ncl / n_out");
  } else {
    WNB_CHECK_ARG(a->y_nlc, "chain_fwd_tc: single-stage call needs y_nlc");
  }
  p.w1_boxrows = a->n1 < 256 ? a->n1 : 256;
  p.w2a_boxrows = a->n2 < 256 ? (a->n2 > 0 ? a->n2 : 16) : 256;
  p.w2x_boxrows = C < 256 ? C : 256;
  CUtensorMap mx, mw1, mw2a, mw2x;
  int rc;
  if ((rc = make_map_x(&mx, a->x, a->B, a->T, C))) return rc;
  if ((rc = make_map_2d(&mw1, a->w1, a->n1, a->ntaps * C, p.w1_boxrows))) return rc;
  mw2a = mw1; mw2x = mw1;
  if (a->n2 > 0) {
    const int k2 = a->use_x2 ? 2 * C : C;
    if ((rc = make_map_2d(&mw2a, a->w2, a->n2, k2, p.w2a_boxrows))) return rc;
    if (a->use_x2 && (rc = make_map_2d(&mw2x, a->w2, a->n2, k2, p.w2x_boxrows))) return rc;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (C) {
    case 64: return launch_chain<64>(mx, mw1, mw2a, mw2x, p, st);
    case 128: return launch_chain<128>(mx, mw1, mw2a, mw2x, p, st);
    default: return launch_chain<256>(mx, mw1, mw2a, mw2x, p, st);
  }
}

// ------------------------------------------------------------------------------------------ layout changes
namespace wnb {
// NCL [B][C][T] (fp32 or bf16) -> NLC bf16 [B][T][C], 32x32 tiles through shared memory.
template <typename T>
__global__ void ncl_to_nlc_kernel(int C, int Tn, const T* x, bf16* y) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    tile[i][tx] = (c < C && t < Tn) ? to_f32<T>(x[((long long)b * C + c) * Tn + t]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (t < Tn && c < C) y[((long long)b * Tn + t) * C + c] = __float2bfloat16_rn(tile[tx][i]);
  }
}
// NLC [B][T][C] (fp32 or bf16) -> NCL [B][C][T] (fp32 or bf16): 64 x 64 (frames x channels) tiles through shared
// memory; loads are 16 consecutive channels per thread, stores 16 consecutive frames per thread (16-byte accesses on
// both sides when the row pitches allow it).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) nlc_to_ncl_kernel(int C, int Tn, const TI* x, TO* y) {
  __shared__ float tile[64][65];              // [channel][frame]
  const int b = blockIdx.z, c0 = blockIdx.y * 64, t0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
  {   // load: thread -> (frame = tid / 4, 16 consecutive channels)
    const int t = tid >> 2, cq = (tid & 3) * 16;
    float v[16];
    const TI* src = x + ((long long)b * Tn + t0 + t) * C + c0 + cq;
    const bool tok = t0 + t < Tn;
    if (tok && c0 + cq + 16 <= C && ((long long)C * sizeof(TI)) % 16 == 0) {
      constexpr int NV = 16 * sizeof(TI) / 16;             // 16-byte vectors for 16 channels
      uint4 raw[NV];
#pragma unroll
      for (int q = 0; q < NV; ++q) raw[q] = __ldg(reinterpret_cast<const uint4*>(src) + q);
      const TI* e = reinterpret_cast<const TI*>(raw);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = to_f32<TI>(e[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (tok && c0 + cq + i < C) ? to_f32<TI>(src[i]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) tile[cq + i][t] = v[i];
  }
  __syncthreads();
  {   // store: thread -> (channel = tid / 4, 16 consecutive frames)
    const int c = tid >> 2, tq = (tid & 3) * 16;
    if (c0 + c < C) {
      TO* dst = y + ((long long)b * C + c0 + c) * Tn + t0 + tq;
      if (t0 + tq + 16 <= Tn && ((long long)Tn * sizeof(TO)) % 16 == 0) {
        constexpr int NV = 16 * sizeof(TO) / 16;
        uint4 raw[NV];
        TO* e = reinterpret_cast<TO*>(raw);
#pragma unroll
        for (int i = 0; i < 16; ++i) e[i] = from_f32<TO>(tile[c][tq + i]);
#pragma unroll
        for (int q = 0; q < NV; ++q) reinterpret_cast<uint4*>(dst)[q] = raw[q];
      } else {
        for (int i = 0; i < 16; ++i)
          if (t0 + tq + i < Tn) dst[i] = from_f32<TO>(tile[c][tq + i]);
      }
    }
  }
}
}  // namespace wnb

namespace wnb { int nlc_to_ncl_bf16_fast_launch(int B, int C, int T_, const void* x, void* y, cudaStream_t st); }

extern "C" int wnb200_nlc_to_ncl(int out_dtype, int src_is_f32, int B, int C, int T_, const void* x, void* y,
                                 void* stream) {
  WNB_CHECK_ARG(x && y, "nlc_to_ncl: null pointer");
  if (B == 0 || C == 0 || T_ == 0) return 0;
  dim3 grid(ceil_div(T_, 64), ceil_div(C, 64), B), block(256);
  cudaStream_t st = (cudaStream_t)stream;
  if (!src_is_f32 && out_dtype == WNB200_BF16 && C % 64 == 0 && B <= 65535 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    // bf16 -> bf16: the word-tile kernel of bytenet.cu (two channels per word, 16-byte accesses on both sides)
    nlc_to_ncl_bf16_fast_launch(B, C, T_, x, y, st);
    WNB_LAUNCH_OK();
    return 0;
  }
  if (src_is_f32) {
    if (out_dtype == WNB200_F32) nlc_to_ncl_kernel<float, float><<<grid, block, 0, st>>>(C, T_, (const float*)x, (float*)y);
    else nlc_to_ncl_kernel<float, bf16><<<grid, block, 0, st>>>(C, T_, (const float*)x, (bf16*)y);
  } else {
    if (out_dtype == WNB200_F32) nlc_to_ncl_kernel<bf16, float><<<grid, block, 0, st>>>(C, T_, (const bf16*)x, (float*)y);
    else nlc_to_ncl_kernel<bf16, bf16><<<grid, block, 0, st>>>(C, T_, (const bf16*)x, (bf16*)y);
  }
  WNB_LAUNCH_OK();
  return 0;
}
