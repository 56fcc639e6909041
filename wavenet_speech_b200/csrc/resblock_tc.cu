// Fused residual block + skip bottleneck, tensor-core path v2 (C = 128 or 256, NLC bf16).
//
// One persistent CTA per SM walks 128-frame tiles.  TMEM (2C fp32 columns) is split into two regions A|B of C
// columns so that the epilogue of one region overlaps the MMAs into the other:
//
//   MMA order per tile                       TMEM     epilogue (8 warps, overlapped)
//   G1a  x taps  * W1[half 0]  (N = C)   ->   A       E1a: gate channels [0, C/2)   -> act blocks (smem)
//   G1b  x taps  * W1[half 1]  (N = C)   ->   B       E1b: gate channels [C/2, C)   -> act blocks (smem)
//   G2r  x(t)*Wproj + act*Wres (N = C)   ->   A       E2a: res  + bias -> bf16 -> smem -> TMA store
//   G2s  act * (Wbn*Wskip)     (N = C)   ->   B       E2b: skip + bias -> fp32 -> smem -> TMA reduce-add
//
// W1 rows are packed per half as [tanh C/2 ; sigmoid C/2] so a thread finds the two pre-activations of one
// channel in the same TMEM lane.  The skip sum is accumulated in HBM by cp.reduce.async.bulk (L2 does the add):
// the SM never reads the running sum.  Weight and activation blocks stream through a ring of (16 KB + C*128 B)
// stages filled by TMA; out-of-range frames are zero-filled by TMA (= the reference's conv padding).
//
// Roles: warp 0 TMA producer, warp 1 TMEM alloc + MMA issue, warps 2..9 epilogue (lane quarter = warp % 4,
// column half = (warp - 2) / 4).
// Reference semantics: modules/block.py:54-82 (ResidualBlock.forward) + modules/wavenet.py:100 (bottleneck add).
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace wnb {
using namespace tc;
typedef __nv_bfloat16 bf16;

struct ResDev {
  int B, T, tiles_per_seq, num_tiles;
  int ntaps, t_off[3];
  const float* bias1;   // [2C] packed like W1 rows
  const float* bias2;   // [2C] = [res ; skip]
  int write_res, skips_init;
  int final_act;    // last layer (CTA-pair kernel, inference): emit LeakyReLU(running sum + contribution) as bf16 instead
  long long* dbg;
  bf16* save_act;   // training: gate / tanh / sigmoid, NLC bf16 (CTA-pair kernel only)
  bf16* save_th;
  bf16* save_sg;
  int has_lo;       // PREC kernels: x_lo is present (the stream's fp16 low half joins the projection)
  int xflags;       // WNB200_TIMELINE builds only: experiment switches: 1 approx gate, 2 no lo store
};

constexpr int RB_THREADS = 320;
constexpr int RB_EPI_THREADS = 256;
constexpr int RB_ABYTES = RB_TILE * 128;

template <int C>
struct RCfg {
  static constexpr int KB = C / 64;
  static constexpr int BBYTES = C * 128;                       // [C rows x 64] bf16
  static constexpr int STAGE = RB_ABYTES + BBYTES;
  static constexpr int ACT = KB * RB_ABYTES;
  static constexpr int STAGING = RB_ABYTES;                    // 16 KB output staging
  static constexpr int NSTAGE = (C == 256) ? 3 : 5;
  static constexpr int SMEM = NSTAGE * STAGE + ACT + STAGING + 1024 + 256;
};

// one 64-channel K block (4 UMMA K-steps), N = C columns
template <int C>
__device__ __forceinline__ void mma_kblock(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, bool first) {
  constexpr uint32_t idesc = make_idesc_bf16(RB_TILE, C);
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4)
    umma_bf16(tmem_d, make_smem_desc_sw128(a_addr + k4 * 32), make_smem_desc_sw128(b_addr + k4 * 32), idesc,
              (first && k4 == 0) ? 0u : 1u);
}

// Per-phase clock64 stamps of CTA 0 (scripts/timeline.py): compiled in only with -DWNB200_TIMELINE (a separate
// diagnostic build, WNB200_TIMELINE=1 python -m wavenet_speech_b200.csrc.build); the shipped kernels carry none.
#ifdef WNB200_TIMELINE
#define RB_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && it < 8) p.dbg[it * 16 + (slot)] = clock64(); } while (0)
#else
#define RB_STAMP(slot) do { } while (0)
#endif

template <int C>
__global__ void __launch_bounds__(RB_THREADS, 1)
resblock_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_res,
                const __grid_constant__ CUtensorMap map_skips, const ResDev p) {
  using K = RCfg<C>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t act_base = smem_base + K::NSTAGE * K::STAGE;
  const uint32_t stg_base = act_base + K::ACT;
  const uint32_t bar_base = stg_base + K::STAGING;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (K::NSTAGE + s); };
  const uint32_t bb = bar_base + 8u * (2 * K::NSTAGE);
  const uint32_t accA_full = bb, accB_full = bb + 8, e1a_done = bb + 16, e1b_done = bb + 24, e2a_done = bb + 32,
                 e2b_done = bb + 40, tmem_slot = bb + 48;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_w1);
    prefetch_tensormap(&map_w2);
    prefetch_tensormap(&map_res);
    prefetch_tensormap(&map_skips);
    for (int s = 0; s < K::NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accA_full, 1);
    mbar_init(accB_full, 1);
    mbar_init(e1a_done, RB_EPI_THREADS);
    mbar_init(e1b_done, RB_EPI_THREADS);
    mbar_init(e2a_done, RB_EPI_THREADS);
    mbar_init(e2b_done, RB_EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * C);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  const uint32_t tmemA = tmem_base, tmemB = tmem_base + C;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto next = [&]() {
        if (++stage == K::NSTAGE) { stage = 0; phase ^= 1; }
      };
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int b = tile / p.tiles_per_seq;
        const int t0 = (tile - b * p.tiles_per_seq) * RB_TILE;
        // G1a, G1b: x taps + W1 half
        for (int half = 0; half < 2; ++half) {
          for (int kb = 0; kb < p.ntaps * K::KB; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t sa = smem_base + stage * K::STAGE;
            mbar_expect_tx(full_bar(stage), K::STAGE);
            const int tap = kb / K::KB, cb = kb - tap * K::KB;
            tma_load_3d(sa, &map_x, full_bar(stage), cb * 64, t0 + p.t_off[tap], b);
            tma_load_2d(sa + RB_ABYTES, &map_w1, full_bar(stage), kb * 64, half * C);
            next();
          }
        }
        // G2r, x(t) part: x block + Wproj block
        for (int kb = 0; kb < K::KB; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * K::STAGE;
          mbar_expect_tx(full_bar(stage), K::STAGE);
          tma_load_3d(sa, &map_x, full_bar(stage), kb * 64, t0, b);
          tma_load_2d(sa + RB_ABYTES, &map_w2, full_bar(stage), C + kb * 64, 0);
          next();
        }
        // G2r act part (Wres) then G2s (folded skip weights): weights only
        for (int part = 0; part < 2; ++part) {
          for (int kb = 0; kb < K::KB; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t sa = smem_base + stage * K::STAGE;
            mbar_expect_tx(full_bar(stage), K::BBYTES);
            tma_load_2d(sa + RB_ABYTES, &map_w2, full_bar(stage), kb * 64, part * C);
            next();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto next = [&]() {
        if (++stage == K::NSTAGE) { stage = 0; phase ^= 1; }
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t prev = (uint32_t)((it - 1) & 1), cur = (uint32_t)(it & 1);
        RB_STAMP(0);
        for (int half = 0; half < 2; ++half) {
          if (it > 0) {
            mbar_wait(half == 0 ? e2a_done : e2b_done, prev);   // region drained by the previous tile's epilogue 2
            tc_fence_after();
          }
          for (int kb = 0; kb < p.ntaps * K::KB; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * K::STAGE;
            mma_kblock<C>(half == 0 ? tmemA : tmemB, sa, sa + RB_ABYTES, kb == 0);
            umma_commit(empty_bar(stage));
            next();
          }
          umma_commit(half == 0 ? accA_full : accB_full);
          RB_STAMP(1 + half);
        }
        // G2r: x(t) * Wproj into region A (needs E1a to have drained A)
        mbar_wait(e1a_done, cur);
        tc_fence_after();
        RB_STAMP(3);
        for (int kb = 0; kb < K::KB; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * K::STAGE;
          mma_kblock<C>(tmemA, sa, sa + RB_ABYTES, kb == 0);
          umma_commit(empty_bar(stage));
          next();
        }
        // G2r: act * Wres (act blocks of half 0 are ready; half 1 after E1b)
        for (int kb = 0; kb < K::KB; ++kb) {
          if (kb == K::KB / 2) {
            mbar_wait(e1b_done, cur);
            tc_fence_after();
            RB_STAMP(4);
          }
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * K::STAGE;
          mma_kblock<C>(tmemA, act_base + kb * RB_ABYTES, sa + RB_ABYTES, false);
          umma_commit(empty_bar(stage));
          next();
        }
        umma_commit(accA_full);
        RB_STAMP(5);
        // G2s: act * (Wbn Wskip) into region B (drained by E1b, waited above)
        for (int kb = 0; kb < K::KB; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * K::STAGE;
          mma_kblock<C>(tmemB, act_base + kb * RB_ABYTES, sa + RB_ABYTES, kb == 0);
          umma_commit(empty_bar(stage));
          next();
        }
        umma_commit(accB_full);
        RB_STAMP(6);
      }
    }
  } else {
    // ================================ epilogue warps ================================
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;                 // column half handled by this warp
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const bool issuer = (threadIdx.x == 64);
    uint8_t* stg_row = smem_gen + (stg_base - smem_base) + row * 128;
    const int sw = row & 7;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int b = tile / p.tiles_per_seq;
      const int t0 = (tile - b * p.tiles_per_seq) * RB_TILE;

      // ---------------- E1a / E1b: gate -> act (bf16, swizzled K-major blocks) ----------------
      for (int half = 0; half < 2; ++half) {
        mbar_wait(half == 0 ? accA_full : accB_full, 0u);
        tc_fence_after();
        if (issuer) RB_STAMP(8 + half);
        const uint32_t treg = (half == 0 ? tmemA : tmemB) + lane_off;
#pragma unroll 1
        for (int cc = 0; cc < C / 4; cc += 16) {
          const int col = h * (C / 4) + cc;        // channel inside the half; tanh col = col, sigmoid col = C/2+col
          float a[16], g[16];
          tmem_ld16(treg + col, a);
          tmem_ld16(treg + C / 2 + col, g);
          const float4* bt = reinterpret_cast<const float4*>(p.bias1 + half * C + col);
          const float4* bs = reinterpret_cast<const float4*>(p.bias1 + half * C + C / 2 + col);
          float bta[16], bsa[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 u = __ldg(bt + j), v = __ldg(bs + j);
            bta[4 * j] = u.x; bta[4 * j + 1] = u.y; bta[4 * j + 2] = u.z; bta[4 * j + 3] = u.w;
            bsa[4 * j] = v.x; bsa[4 * j + 1] = v.y; bsa[4 * j + 2] = v.z; bsa[4 * j + 3] = v.w;
          }
          tmem_wait_ld();
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float v0 = tanh_approx(a[i] + bta[i]) * sigmoid_approx(g[i] + bsa[i]);
            const float v1 = tanh_approx(a[i + 1] + bta[i + 1]) * sigmoid_approx(g[i + 1] + bsa[i + 1]);
            pk[i >> 1] = pack_bf16x2(v0, v1);
          }
          const int ch = half * (C / 2) + col;     // global channel of a[0]
          const int kb = ch >> 6, ci = (ch & 63) >> 3;
          uint8_t* blk = smem_gen + (act_base - smem_base) + kb * RB_ABYTES + row * 128;
          *reinterpret_cast<uint4*>(blk + ((ci ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(blk + (((ci + 1) ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(half == 0 ? e1a_done : e1b_done);
      }

      // ---------------- E2a: res = acc + bias -> bf16 -> staging -> TMA store ----------------
      mbar_wait(accA_full, 1u);
      tc_fence_after();
      if (issuer) RB_STAMP(10);
      if (p.write_res) {
#pragma unroll 1
        for (int c = 0; c < C / 64; ++c) {
          const int col = c * 64 + h * 32;
          float a[32];
          tmem_ld16(tmemA + lane_off + col, a);
          tmem_ld16(tmemA + lane_off + col + 16, a + 16);
          const float4* bp = reinterpret_cast<const float4*>(p.bias2 + col);
          float bv[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 u = __ldg(bp + j);
            bv[4 * j] = u.x; bv[4 * j + 1] = u.y; bv[4 * j + 2] = u.z; bv[4 * j + 3] = u.w;
          }
          tmem_wait_ld();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) pk[i >> 1] = pack_bf16x2(a[i] + bv[i], a[i + 1] + bv[i + 1]);
          if (issuer) bulk_wait_read0();            // previous TMA op has finished reading the staging buffer
          epi_bar();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(stg_row + (((4 * h + j) ^ sw) << 4)) =
                make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          fence_proxy_async_smem();
          epi_bar();
          if (issuer) {
            tma_store_3d(&map_res, stg_base, c * 64, t0, b);
            bulk_commit();
          }
        }
      }
      tc_fence_before();
      mbar_arrive(e2a_done);

      // ---------------- E2b: skips (+)= acc + bias (fp32) -> staging -> TMA reduce-add ----------------
      mbar_wait(accB_full, 1u);
      tc_fence_after();
      if (issuer) RB_STAMP(11);
#pragma unroll 1
      for (int c = 0; c < C / 32; ++c) {
        const int col = c * 32 + h * 16;
        float a[16];
        tmem_ld16(tmemB + lane_off + col, a);
        const float4* bp = reinterpret_cast<const float4*>(p.bias2 + C + col);
        float bv[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 u = __ldg(bp + j);
          bv[4 * j] = u.x; bv[4 * j + 1] = u.y; bv[4 * j + 2] = u.z; bv[4 * j + 3] = u.w;
        }
        tmem_wait_ld();
        if (issuer) bulk_wait_read0();
        epi_bar();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(stg_row + (((4 * h + j) ^ sw) << 4)) =
              make_float4(a[4 * j] + bv[4 * j], a[4 * j + 1] + bv[4 * j + 1], a[4 * j + 2] + bv[4 * j + 2],
                          a[4 * j + 3] + bv[4 * j + 3]);
        fence_proxy_async_smem();
        epi_bar();
        if (issuer) {
          if (p.skips_init) tma_store_3d(&map_skips, stg_base, c * 32, t0, b);
          else tma_reduce_add_3d(&map_skips, stg_base, c * 32, t0, b);
          bulk_commit();
        }
      }
      tc_fence_before();
      mbar_arrive(e2b_done);
      if (issuer) RB_STAMP(12);
    }
    if (issuer) bulk_wait0();      // all output traffic issued by this CTA has completed
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * C);
  }
}


// =========================================================================================================
// CTA-pair variant (cta_group::2): two CTAs of a cluster process 256 consecutive frames.  Each CTA owns 128
// frames (its own A tiles, act tile, TMEM lanes and outputs); every weight block is split between the two CTAs
// (each loads C/2 of its C rows) and the leader CTA's single MMA thread issues M=256 tcgen05.mma instructions
// that read both shared memories and write both tensor memories.  Compared with the single-CTA kernel this
// halves the weight bytes each SM pulls from L2 (the limiter there) and shrinks a stage to 32 KB.
// Barriers: full[s] lives in the leader (count 2 = leader's expect_tx + peer's arrive; both CTAs' TMA bytes are
// signalled on it); empty[s] / acc*_full are per CTA and receive the multicast tcgen05.commit; the four
// epilogue->MMA barriers live in the leader and collect 2 x 256 arrivals.
// =========================================================================================================
template <int C>
struct R2Cfg {
  static constexpr int KB = C / 64;
  static constexpr int BHBYTES = (C / 2) * 128;                // this CTA's half of a [C x 64] weight block
  static constexpr int STAGE = RB_ABYTES + BHBYTES;
  static constexpr int ACT = KB * RB_ABYTES;
  static constexpr int STAGING = 2 * RB_ABYTES;                // double-buffered output staging
  static constexpr int NSTAGE = (C == 256) ? 4 : 6;
  static constexpr bool RETAIN = (NSTAGE == KB);               // a phase of KB K-blocks walks the whole ring once
  static constexpr int SMEM = NSTAGE * STAGE + ACT + STAGING + 1024 + 256;
};

template <int C, bool F16 = false>
__device__ __forceinline__ void mma_kblock_2sm(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, bool first) {
  constexpr uint32_t idesc = make_idesc(F16, 2 * RB_TILE, C);
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4)
    umma_bf16_2sm(tmem_d, make_smem_desc_sw128(a_addr + k4 * 32), make_smem_desc_sw128(b_addr + k4 * 32), idesc,
                  (first && k4 == 0) ? 0u : 1u);
}

// SAVE (training): the gate and its two factors are kept for backward.  E1 then works on whole 64-channel blocks:
// tanh / sigmoid go through the two output staging buffers, the gate is stored straight from the shared tile that
// feeds G2, all three by TMA (thread-level global stores from the TMEM-lane layout cost a 128-byte line per lane).
//
// PREC (WNB200_ACT_F16X2, inference): the accuracy mode that holds the stated bf16-class tolerance (2e-2 on the logits)
// at the depth of the benchmarked stacks.  With the reference's initialisation the residual projection is a random
// nn.Linear (block.py:48,77-78), the stream grows ~1.45x per block and every rounding of it is amplified by the blocks
// that follow; a bf16 stream + bf16 gate + tanh.approx is 5e-2..9e-2 off at 16-20 blocks (so is PyTorch's own bf16
// evaluation).  Here operands are fp16 (same tensor-core rate, 3 more mantissa bits; weights that were rounded to bf16
// are exact in fp16), the stream is carried between layers as an fp16 (hi, lo) pair -- the dilated taps read hi, the
// projection contracts hi AND lo against the same Wproj blocks (one extra K slab: 8 C^2 instead of 7 C^2 MAC per
// frame; its MMAs run while E1b is still producing the second half of the gate) -- and the gate is evaluated with
// ex2/rcp (1e-6) instead of tanh.approx (5e-4).  map_act / map_th double as the x_lo / res_lo maps.
template <int C, bool SAVE, bool PREC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(RB_THREADS, 1)
resblock2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_res,
                 const __grid_constant__ CUtensorMap map_skips, const __grid_constant__ CUtensorMap map_act,
                 const __grid_constant__ CUtensorMap map_th, const __grid_constant__ CUtensorMap map_sg,
                 const ResDev p) {
  static_assert(!(SAVE && PREC), "the training forward keeps the bf16 format");
  using K = R2Cfg<C>;
  const CUtensorMap& map_xlo = map_act;      // PREC: fp16 low half of the input stream / of the output stream
  const CUtensorMap& map_reslo = map_th;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t act_base = smem_base + K::NSTAGE * K::STAGE;
  const uint32_t stg_base = act_base + K::ACT;
  const uint32_t bar_base = stg_base + K::STAGING;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (K::NSTAGE + s); };
  const uint32_t bb = bar_base + 8u * (2 * K::NSTAGE);
  const uint32_t accA_full = bb, accB_full = bb + 8, e1a_done = bb + 16, e1b_done = bb + 24, e2a_done = bb + 32,
                 e2b_done = bb + 40, tmem_slot = bb + 48, ld_bar = bb + 56;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_w1);
    prefetch_tensormap(&map_w2);
    prefetch_tensormap(&map_res);
    prefetch_tensormap(&map_skips);
    if (PREC) {
      prefetch_tensormap(&map_xlo);
      prefetch_tensormap(&map_reslo);
    }
    for (int s = 0; s < K::NSTAGE; ++s) {
      mbar_init(full_bar(s), 2);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accA_full, 1);
    mbar_init(accB_full, 1);
    mbar_init(e1a_done, 2 * RB_EPI_THREADS);
    mbar_init(e1b_done, 2 * RB_EPI_THREADS);
    mbar_init(e2a_done, 2 * RB_EPI_THREADS);
    mbar_init(e2b_done, 2 * RB_EPI_THREADS);
    mbar_init(ld_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, 2 * C);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();          // barriers of both CTAs initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  const uint32_t tmemA = tmem_base, tmemB = tmem_base + C;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto next = [&]() {
        if (++stage == K::NSTAGE) { stage = 0; phase ^= 1; }
      };
      const int wrow = (int)rank * (C / 2);       // this CTA's rows inside a [C x 64] weight block
      int ord[3] = {0, 1, 2};                     // tap order: the zero-offset tap (x(t)) first
      for (int j = 1; j < p.ntaps; ++j)
        if (p.t_off[j] == 0) { ord[0] = j; for (int i = 1; i <= j; ++i) ord[i] = i - 1; }
      auto begin_stage = [&](uint32_t bytes_per_cta) -> uint32_t {
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * bytes_per_cta);
        return mapa_shared(full_bar(stage), 0);
      };
      auto end_stage = [&](uint32_t lfull) {
        if (rank != 0) mbar_arrive_cluster(lfull);
        next();
      };
      for (int pt = pair; pt < p.num_tiles; pt += npairs) {
        const int b = pt / p.tiles_per_seq;
        const int t0 = (pt - b * p.tiles_per_seq) * (2 * RB_TILE) + (int)rank * RB_TILE;
        // G1a walks the taps as ord[0..ntaps), G1b backwards, the projection follows G1b's last tap (= ord[0],
        // the zero-offset tap when there is one).  When the ring is exactly KB stages deep, the activation blocks
        // of the tap a phase starts with are still sitting in the stages it is about to use: only the weight half
        // of those stages is reloaded (12 instead of 20 activation blocks per tile at k = 2, C = 256).
        for (int half = 0; half < 2; ++half) {
          for (int jj = 0; jj < p.ntaps; ++jj) {
            const int tap = half == 0 ? ord[jj] : ord[p.ntaps - 1 - jj];
            const bool keep = K::RETAIN && half == 1 && jj == 0;
            for (int cb = 0; cb < K::KB; ++cb) {
              const uint32_t sa = smem_base + stage * K::STAGE;
              const uint32_t lfull = begin_stage(keep ? K::BHBYTES : K::STAGE);
              if (!keep) tma_load_3d_2sm(sa, &map_x, lfull, cb * 64, t0 + p.t_off[tap], b);
              tma_load_2d_2sm(sa + RB_ABYTES, &map_w1, lfull, (tap * K::KB + cb) * 64, half * C + wrow);
              end_stage(lfull);
            }
          }
        }
        const bool keep_x = K::RETAIN && p.t_off[ord[0]] == 0;
        for (int kb = 0; kb < K::KB; ++kb) {
          const uint32_t sa = smem_base + stage * K::STAGE;
          const uint32_t lfull = begin_stage(keep_x ? K::BHBYTES : K::STAGE);
          if (!keep_x) tma_load_3d_2sm(sa, &map_x, lfull, kb * 64, t0, b);
          tma_load_2d_2sm(sa + RB_ABYTES, &map_w2, lfull, C + kb * 64, wrow);
          end_stage(lfull);
        }
        if (PREC && p.has_lo) {
          // the stream's low half against the same Wproj blocks: with a ring exactly KB deep the weight half of each
          // stage still holds the block the hi pass used -- only the 16 KB activation block is loaded
          for (int kb = 0; kb < K::KB; ++kb) {
            const uint32_t sa = smem_base + stage * K::STAGE;
            const uint32_t lfull = begin_stage(K::RETAIN ? RB_ABYTES : K::STAGE);
            tma_load_3d_2sm(sa, &map_xlo, lfull, kb * 64, t0, b);
            if (!K::RETAIN) tma_load_2d_2sm(sa + RB_ABYTES, &map_w2, lfull, C + kb * 64, wrow);
            end_stage(lfull);
          }
        }
        for (int part = 0; part < 2; ++part) {
          for (int kb = 0; kb < K::KB; ++kb) {
            const uint32_t sa = smem_base + stage * K::STAGE;
            const uint32_t lfull = begin_stage(K::BHBYTES);
            tma_load_2d_2sm(sa + RB_ABYTES, &map_w2, lfull, kb * 64, part * C + wrow);
            end_stage(lfull);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto next = [&]() {
        if (++stage == K::NSTAGE) { stage = 0; phase ^= 1; }
      };
      int it = 0;
      for (int pt = pair; pt < p.num_tiles; pt += npairs, ++it) {
        const uint32_t prev = (uint32_t)((it - 1) & 1), cur = (uint32_t)(it & 1);
#ifdef WNB200_TIMELINE
        long long wfull = 0;
#endif
        RB_STAMP(0);
        for (int half = 0; half < 2; ++half) {
          if (it > 0) {
            mbar_wait(half == 0 ? e2a_done : e2b_done, prev);
            tc_fence_after();
          }
          for (int kb = 0; kb < p.ntaps * K::KB; ++kb) {
#ifdef WNB200_TIMELINE
            { const long long c0 = clock64(); mbar_wait(full_bar(stage), phase); wfull += clock64() - c0; }
#else
            mbar_wait(full_bar(stage), phase);
#endif
            tc_fence_after();
            const uint32_t sa = smem_base + stage * K::STAGE;
            mma_kblock_2sm<C, PREC>(half == 0 ? tmemA : tmemB, sa, sa + RB_ABYTES, kb == 0);
            umma_commit_2sm(empty_bar(stage));
            next();
          }
          umma_commit_2sm(half == 0 ? accA_full : accB_full);
          RB_STAMP(1 + half);
#ifdef WNB200_TIMELINE
          if (p.dbg && blockIdx.x == 0 && it < 8) p.dbg[it * 16 + 13 + half] = wfull;
#endif
        }
        mbar_wait(e1a_done, cur);
        tc_fence_after();
        RB_STAMP(3);
        for (int kb = 0; kb < K::KB; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * K::STAGE;
          mma_kblock_2sm<C, PREC>(tmemA, sa, sa + RB_ABYTES, kb == 0);
          umma_commit_2sm(empty_bar(stage));
          next();
        }
        if (PREC && p.has_lo) {
          for (int kb = 0; kb < K::KB; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * K::STAGE;
            mma_kblock_2sm<C, PREC>(tmemA, sa, sa + RB_ABYTES, false);
            umma_commit_2sm(empty_bar(stage));
            next();
          }
        }
        for (int kb = 0; kb < K::KB; ++kb) {
          if (kb == K::KB / 2) {
            mbar_wait(e1b_done, cur);
            tc_fence_after();
            RB_STAMP(4);
          }
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * K::STAGE;
          mma_kblock_2sm<C, PREC>(tmemA, act_base + kb * RB_ABYTES, sa + RB_ABYTES, false);
          umma_commit_2sm(empty_bar(stage));
          next();
        }
        umma_commit_2sm(accA_full);
        RB_STAMP(5);
        for (int kb = 0; kb < K::KB; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * K::STAGE;
          mma_kblock_2sm<C, PREC>(tmemB, act_base + kb * RB_ABYTES, sa + RB_ABYTES, kb == 0);
          umma_commit_2sm(empty_bar(stage));
          next();
        }
        umma_commit_2sm(accB_full);
        RB_STAMP(6);
      }
    }
  } else {
    // ================================ epilogue warps (both CTAs) ================================
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const bool issuer = (threadIdx.x == 64);
    const int sw = row & 7;
    const uint32_t r_e1a = mapa_shared(e1a_done, 0), r_e1b = mapa_shared(e1b_done, 0),
                   r_e2a = mapa_shared(e2a_done, 0), r_e2b = mapa_shared(e2b_done, 0);
    uint32_t nchunk = 0, ld_phase = 0;
    int it = 0;
    for (int pt = pair; pt < p.num_tiles; pt += npairs, ++it) {
      const int b = pt / p.tiles_per_seq;
      const int t0 = (pt - b * p.tiles_per_seq) * (2 * RB_TILE) + (int)rank * RB_TILE;

      for (int half = 0; half < 2; ++half) {
        // The biases come from L2 (the unified L1 is all shared memory here): ~400 cycles per load round if they are
        // fetched where they are used (ncu, r2: 12 % of the epilogue warps' samples sat on the first FADD/FFMA after
        // them).  Every epilogue phase therefore fetches the NEXT step's biases while it works on the current one, the
        // first step's before it waits for its accumulator.
        float4 nb_t[4], nb_s[4];
        if constexpr (!SAVE) {
          const float4* bt = reinterpret_cast<const float4*>(p.bias1 + half * C + h * (C / 4));
          const float4* bs = reinterpret_cast<const float4*>(p.bias1 + half * C + C / 2 + h * (C / 4));
#pragma unroll
          for (int j = 0; j < 4; ++j) { nb_t[j] = __ldg(bt + j); nb_s[j] = __ldg(bs + j); }
        }
        mbar_wait(half == 0 ? accA_full : accB_full, 0u);
        tc_fence_after();
        if (issuer && rank == 0) RB_STAMP(8 + half);
        const uint32_t treg = (half == 0 ? tmemA : tmemB) + lane_off;
        if constexpr (SAVE) {
#pragma unroll 1
          for (int c = 0; c < C / 128; ++c) {          // one 64-channel block of this half per pass
            const int col = c * 64 + h * 32;
            float a[32], g[32];
            tmem_ld16(treg + col, a);
            tmem_ld16(treg + col + 16, a + 16);
            tmem_ld16(treg + C / 2 + col, g);
            tmem_ld16(treg + C / 2 + col + 16, g + 16);
            const float* bt = p.bias1 + half * C + col;
            const float* bs = bt + C / 2;
            tmem_wait_ld();
            uint32_t pk[16], pt[16], ps[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float t0v = tanh_approx(a[i] + __ldg(bt + i)), s0v = sigmoid_approx(g[i] + __ldg(bs + i));
              const float t1v = tanh_approx(a[i + 1] + __ldg(bt + i + 1));
              const float s1v = sigmoid_approx(g[i + 1] + __ldg(bs + i + 1));
              pk[i >> 1] = pack_bf16x2(t0v * s0v, t1v * s1v);
              pt[i >> 1] = pack_bf16x2(t0v, t1v);
              ps[i >> 1] = pack_bf16x2(s0v, s1v);
            }
            const int ch0 = half * (C / 2) + c * 64;   // first channel of the block
            // backward needs the gate (an MMA operand of two weight gradients) and ONE of its factors: with sigmoid kept,
            // tanh = gate / sigmoid in the gate-backward kernel.  tanh is stored only on request (save_th): without it
            // the two staging buffers alternate between passes and a pass no longer waits for the previous pass's stores.
            const bool keep_th = p.save_th != nullptr;
            const uint32_t sgoff = keep_th ? RB_ABYTES : (nchunk & 1u) * RB_ABYTES;
            if (issuer) {
              if (keep_th) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            epi_bar();
            uint8_t* blk = smem_gen + (act_base - smem_base) + (ch0 >> 6) * RB_ABYTES + row * 128;
            uint8_t* sth = smem_gen + (stg_base - smem_base) + row * 128;
            uint8_t* ssg = sth + sgoff;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int o = ((4 * h + j) ^ sw) << 4;
              *reinterpret_cast<uint4*>(blk + o) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
              if (keep_th)
                *reinterpret_cast<uint4*>(sth + o) = make_uint4(pt[4 * j], pt[4 * j + 1], pt[4 * j + 2], pt[4 * j + 3]);
              *reinterpret_cast<uint4*>(ssg + o) = make_uint4(ps[4 * j], ps[4 * j + 1], ps[4 * j + 2], ps[4 * j + 3]);
            }
            fence_proxy_async_smem();
            epi_bar();
            ++nchunk;
            if (issuer) {
              tma_store_3d(&map_act, act_base + (ch0 >> 6) * RB_ABYTES, ch0, t0, b);
              if (keep_th) tma_store_3d(&map_th, stg_base, ch0, t0, b);
              tma_store_3d(&map_sg, stg_base + sgoff, ch0, t0, b);
              bulk_commit();
            }
          }
        } else {
#pragma unroll 1
        for (int cc = 0; cc < C / 4; cc += 16) {
          const int col = h * (C / 4) + cc;
          float ac[16], gc[16];
          tmem_ld16(treg + col, ac);
          tmem_ld16(treg + C / 2 + col, gc);
          float bta[16], bsa[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 u = nb_t[j], v = nb_s[j];
            bta[4 * j] = u.x; bta[4 * j + 1] = u.y; bta[4 * j + 2] = u.z; bta[4 * j + 3] = u.w;
            bsa[4 * j] = v.x; bsa[4 * j + 1] = v.y; bsa[4 * j + 2] = v.z; bsa[4 * j + 3] = v.w;
          }
          tmem_wait_ld();
          if (cc + 16 < C / 4) {
            const float4* bt = reinterpret_cast<const float4*>(p.bias1 + half * C + col + 16);
            const float4* bs = reinterpret_cast<const float4*>(p.bias1 + half * C + C / 2 + col + 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) { nb_t[j] = __ldg(bt + j); nb_s[j] = __ldg(bs + j); }
          }
          uint32_t pk[8];
#ifdef WNB200_TIMELINE
          if (PREC && (p.xflags & 1)) {        // experiment: the approximate gate on the pre-scaled arguments
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float v0 = tanh_approx(fmaf(ac[i], -2.885390081777927f, bta[i]) * -0.34657359f) *
                               sigmoid_approx(fmaf(gc[i], -1.4426950408889634f, bsa[i]) * -0.69314718f);
              const float v1 = tanh_approx(fmaf(ac[i + 1], -2.885390081777927f, bta[i + 1]) * -0.34657359f) *
                               sigmoid_approx(fmaf(gc[i + 1], -1.4426950408889634f, bsa[i + 1]) * -0.69314718f);
              pk[i >> 1] = pack_f16x2(v0, v1);
            }
          } else
#endif
          if constexpr (PREC) {
            // bias1 arrives pre-scaled: tanh rows by -2 log2(e), sigmoid rows by -log2(e)
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float v0 = gate_precise(fmaf(ac[i], -2.885390081777927f, bta[i]), fmaf(gc[i], -1.4426950408889634f, bsa[i]));
              const float v1 = gate_precise(fmaf(ac[i + 1], -2.885390081777927f, bta[i + 1]),
                                            fmaf(gc[i + 1], -1.4426950408889634f, bsa[i + 1]));
              pk[i >> 1] = pack_f16x2(v0, v1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float v0 = tanh_approx(ac[i] + bta[i]) * sigmoid_approx(gc[i] + bsa[i]);
              const float v1 = tanh_approx(ac[i + 1] + bta[i + 1]) * sigmoid_approx(gc[i + 1] + bsa[i + 1]);
              pk[i >> 1] = pack_bf16x2(v0, v1);
            }
          }
          const int ch = half * (C / 2) + col;
          const int kb = ch >> 6, ci = (ch & 63) >> 3;
          uint8_t* blk = smem_gen + (act_base - smem_base) + kb * RB_ABYTES + row * 128;
          *reinterpret_cast<uint4*>(blk + ((ci ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(blk + (((ci + 1) ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive_cluster(half == 0 ? r_e1a : r_e1b);
      }

      bool drain = SAVE;        // E1's stores still read the staging buffers: drain them before the first reuse
      // ---- E2a: res ----
      float4 nbv[8];                                  // biases of the next chunk (see E1)
      if (p.write_res) {
        const float4* bp = reinterpret_cast<const float4*>(p.bias2 + h * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) nbv[j] = __ldg(bp + j);
      }
      mbar_wait(accA_full, 1u);
      tc_fence_after();
      if (issuer && rank == 0) RB_STAMP(10);
      {
      if (p.write_res) {
#pragma unroll 1
        for (int c = 0; c < C / 64; ++c, ++nchunk) {
          const int col = c * 64 + h * 32;
          float a[32];
          tmem_ld16(tmemA + lane_off + col, a);
          tmem_ld16(tmemA + lane_off + col + 16, a + 16);
          float bv[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 u = nbv[j];
            bv[4 * j] = u.x; bv[4 * j + 1] = u.y; bv[4 * j + 2] = u.z; bv[4 * j + 3] = u.w;
          }
          if (c + 1 < C / 64) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias2 + col + 64);
#pragma unroll
            for (int j = 0; j < 8; ++j) nbv[j] = __ldg(bp + j);
          }
          tmem_wait_ld();
          uint32_t pk[16];
          if constexpr (PREC) {
            // the stream leaves as an fp16 (hi, lo) pair: both staging buffers per 64-channel chunk
            uint32_t pl[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) split_f16x2(a[i] + bv[i], a[i + 1] + bv[i + 1], pk[i >> 1], pl[i >> 1]);
            if (issuer) bulk_wait_read0();
            epi_bar();
            uint8_t* srow = smem_gen + (stg_base - smem_base) + row * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int o = ((4 * h + j) ^ sw) << 4;
              *reinterpret_cast<uint4*>(srow + o) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
              *reinterpret_cast<uint4*>(srow + RB_ABYTES + o) =
                  make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
            }
            fence_proxy_async_smem();
            epi_bar();
            if (issuer) {
              tma_store_3d(&map_res, stg_base, c * 64, t0, b);
#ifdef WNB200_TIMELINE
              if (!(p.xflags & 2))
#endif
              tma_store_3d(&map_reslo, stg_base + RB_ABYTES, c * 64, t0, b);
              bulk_commit();
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) pk[i >> 1] = pack_bf16x2(a[i] + bv[i], a[i + 1] + bv[i + 1]);
            const uint32_t boff = (nchunk & 1u) * RB_ABYTES;
            if (issuer) {
              if (drain) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            drain = false;
            epi_bar();
            uint8_t* srow = smem_gen + (stg_base - smem_base) + boff + row * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(srow + (((4 * h + j) ^ sw) << 4)) =
                  make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            fence_proxy_async_smem();
            epi_bar();
            if (issuer) {
              tma_store_3d(&map_res, stg_base + boff, c * 64, t0, b);
              bulk_commit();
            }
          }
        }
        if (PREC) drain = true;      // one bulk group reads both buffers: the skip chunks start after it has drained
      }
      tc_fence_before();
      mbar_arrive_cluster(r_e2a);
      }

      // ---- E2b: skips ----
      float4 nb2[4];
      if (!p.final_act) {
        const float4* bp = reinterpret_cast<const float4*>(p.bias2 + C + h * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) nb2[j] = __ldg(bp + j);
      }
      mbar_wait(accB_full, 1u);
      tc_fence_after();
      if (issuer && rank == 0) RB_STAMP(11);
      if (p.final_act) {
        // Last layer of an inference stack (wavenet.py:100-103): nothing accumulates after this launch, so instead
        // of adding the tile into the running sum and leaving LeakyReLU + bf16 conversion to another pass over it,
        // the running sum's fp32 chunks come IN by TMA (KB at a time, into the idle gate blocks), meet the
        // accumulator in registers and leave as bf16 64-channel chunks through the staging buffers:
        // skips_act = LeakyReLU(sum + (acc + bias)) -- the same fp32 additions the L2 adder would have done.
#pragma unroll 1
        for (int w0 = 0; w0 < C / 32; w0 += K::KB) {
          if (w0 > 0) epi_bar();                 // every thread is done with the previous wave's landing buffers
          if (issuer) {
            // every store this thread issued has left shared memory: the previous wave's (staging buffers) and, when
            // the gate factors are saved, E1's stores of the gate blocks the loads below land in
            bulk_wait_read0();
            if (!p.skips_init) {
              mbar_expect_tx(ld_bar, K::KB * RB_ABYTES);
#pragma unroll
              for (int u = 0; u < K::KB; ++u)
                tma_load_3d(act_base + u * RB_ABYTES, &map_skips, ld_bar, (w0 + u) * 32, t0, b);
            }
          }
          if (!p.skips_init) {
            mbar_wait(ld_bar, ld_phase);
            ld_phase ^= 1u;
          }
          epi_bar();
#pragma unroll
          for (int u = 0; u < K::KB; ++u) {
            const int col = (w0 + u) * 32 + h * 16;
            float a[16];
            tmem_ld16(tmemB + lane_off + col, a);
            const float4* bp = reinterpret_cast<const float4*>(p.bias2 + C + col);
            float bv[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 v4 = __ldg(bp + j);
              bv[4 * j] = v4.x; bv[4 * j + 1] = v4.y; bv[4 * j + 2] = v4.z; bv[4 * j + 3] = v4.w;
            }
            float sm[16];
            const uint8_t* lrow = smem_gen + (act_base - smem_base) + u * RB_ABYTES + row * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
              if (!p.skips_init) q = *reinterpret_cast<const float4*>(lrow + (((4 * h + j) ^ sw) << 4));
              sm[4 * j] = q.x; sm[4 * j + 1] = q.y; sm[4 * j + 2] = q.z; sm[4 * j + 3] = q.w;
            }
            tmem_wait_ld();
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2)
              pk[i >> 1] = pack_act2<PREC>(leaky(sm[i] + (a[i] + bv[i])), leaky(sm[i + 1] + (a[i + 1] + bv[i + 1])));
            // bf16 row of 64 channels = 8 sixteen-byte pieces; this thread's 16 channels are pieces (u&1)*4 + 2h, +1
            uint8_t* orow = smem_gen + (stg_base - smem_base) + (u >> 1) * RB_ABYTES + row * 128;
            const int c16 = (u & 1) * 4 + 2 * h;
            *reinterpret_cast<uint4*>(orow + ((c16 ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(orow + (((c16 + 1) ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          fence_proxy_async_smem();
          epi_bar();
          if (issuer) {
#pragma unroll
            for (int g = 0; g < K::KB / 2; ++g)
              tma_store_3d(&map_res, stg_base + g * RB_ABYTES, w0 * 32 + g * 64, t0, b);   // (res is not written)
            bulk_commit();
          }
        }
      } else
#pragma unroll 1
      for (int c = 0; c < C / 32; ++c, ++nchunk) {
        const int col = c * 32 + h * 16;
        float a[16];
        tmem_ld16(tmemB + lane_off + col, a);
        float bv[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 u = nb2[j];
          bv[4 * j] = u.x; bv[4 * j + 1] = u.y; bv[4 * j + 2] = u.z; bv[4 * j + 3] = u.w;
        }
        if (c + 1 < C / 32) {
          const float4* bp = reinterpret_cast<const float4*>(p.bias2 + C + col + 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) nb2[j] = __ldg(bp + j);
        }
        tmem_wait_ld();
        const uint32_t boff = (nchunk & 1u) * RB_ABYTES;
        if (issuer) {
          if (drain) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        drain = false;
        epi_bar();
        uint8_t* srow = smem_gen + (stg_base - smem_base) + boff + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(srow + (((4 * h + j) ^ sw) << 4)) =
              make_float4(a[4 * j] + bv[4 * j], a[4 * j + 1] + bv[4 * j + 1], a[4 * j + 2] + bv[4 * j + 2],
                          a[4 * j + 3] + bv[4 * j + 3]);
        fence_proxy_async_smem();
        epi_bar();
        if (issuer) {
          if (p.skips_init) tma_store_3d(&map_skips, stg_base + boff, c * 32, t0, b);
          else tma_reduce_add_3d(&map_skips, stg_base + boff, c * 32, t0, b);
          bulk_commit();
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(r_e2b);
      if (issuer && rank == 0) RB_STAMP(12);
    }
    if (issuer) bulk_wait0();
  }

  tc_fence_before();
  cluster_sync_all();          // the peer's shared / tensor memory must outlive the leader's last MMA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * C);
  }
}

// ------------------------------------------------------------------------------------------ host side
template <int C>
static int launch_resblock(const CUtensorMap& mx, const CUtensorMap& mw1, const CUtensorMap& mw2,
                           const CUtensorMap& mres, const CUtensorMap& msk, const ResDev& p, cudaStream_t st) {
  using K = RCfg<C>;
  WNB_SET_SMEM_ATTR(K::SMEM, resblock_kernel<C>);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  resblock_kernel<C><<<grid, RB_THREADS, K::SMEM, st>>>(mx, mw1, mw2, mres, msk, p);
  WNB_LAUNCH_OK();
  return 0;
}

template <int C, bool SAVE, bool PREC = false>
static int launch_resblock2(const CUtensorMap& mx, const CUtensorMap& mw1, const CUtensorMap& mw2,
                            const CUtensorMap& mres, const CUtensorMap& msk, const CUtensorMap& mact,
                            const CUtensorMap& mth, const CUtensorMap& msg, const ResDev& p, cudaStream_t st) {
  using K = R2Cfg<C>;
  WNB_SET_SMEM_ATTR(K::SMEM, resblock2_kernel<C, SAVE, PREC>);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int pairs = sms / 2;
  if (p.num_tiles < pairs) pairs = p.num_tiles;
  resblock2_kernel<C, SAVE, PREC><<<2 * pairs, RB_THREADS, K::SMEM, st>>>(mx, mw1, mw2, mres, msk, mact, mth, msg, p);
  WNB_LAUNCH_OK();
  return 0;
}

__global__ void leaky_to_bf16_kernel(long long n4, const float4* x, uint2* y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    uint2 o;
    o.x = pack_bf16x2(leaky(v.x), leaky(v.y));
    o.y = pack_bf16x2(leaky(v.z), leaky(v.w));
    y[i] = o;
  }
}

}  // namespace wnb

using namespace wnb;

namespace wnb {
int resblock3_launch(const wnb200_resblock_t* a, void* stream);      // resblock3_tc.cu: deferred-skip variant
}

extern "C" int wnb200_resblock_fwd_tc(const wnb200_resblock_t* a, void* stream) {
  WNB_CHECK_ARG(a != nullptr, "resblock_fwd_tc: null argument");
  WNB_CHECK_STRUCT(a, wnb200_resblock_t, "resblock_fwd_tc");
  const int C = a->C;
  WNB_CHECK_ARG(C == 128 || C == 256, "resblock_fwd_tc: C=%d not in {128,256}", C);
  WNB_CHECK_ARG(a->ntaps >= 1 && a->ntaps <= 3, "resblock_fwd_tc: ntaps=%d not in 1..3", a->ntaps);
  if (a->B == 0 || a->T == 0) return 0;
  WNB_CHECK_ARG(a->act_fmt == WNB200_ACT_BF16 || a->act_fmt == WNB200_ACT_F16X2, "resblock_fwd_tc: bad act_fmt %d",
                a->act_fmt);
  if (a->gate_out) {       // deferred skip: the gate is stored, the skip sum is one stack-wide contraction afterwards
    WNB_CHECK_ARG(a->x && a->w1 && a->w2 && a->bias1 && a->bias2, "resblock_fwd_tc: null pointer");
    return resblock3_launch(a, stream);
  }
  WNB_CHECK_ARG(a->x && a->w1 && a->w2 && a->bias1 && a->bias2 && a->skips, "resblock_fwd_tc: null pointer");
  ResDev p;
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.T = a->T;
  p.tiles_per_seq = ceil_div(a->T, RB_TILE);
  p.num_tiles = p.tiles_per_seq * a->B;
  p.ntaps = a->ntaps;
  for (int j = 0; j < 3; ++j) p.t_off[j] = a->t_off[j];
  p.bias1 = a->bias1; p.bias2 = a->bias2;
  p.write_res = a->res != nullptr; p.skips_init = a->skips_init;
  p.final_act = a->skips_act != nullptr;
  WNB_CHECK_ARG(!p.final_act || ((a->variant & 3) != 1 && !a->res),
                "resblock_fwd_tc: skips_act needs the CTA-pair kernel and the last layer (res = NULL)");
  p.dbg = (long long*)a->dbg;
  p.save_act = (bf16*)a->save_act; p.save_th = (bf16*)a->save_th; p.save_sg = (bf16*)a->save_sg;
  WNB_CHECK_ARG(!a->save_act || (a->save_sg && (a->variant & 3) != 1),
                "resblock_fwd_tc: saving for backward needs save_act and save_sg (save_th optional) and the CTA-pair kernel");
  WNB_CHECK_ARG(a->save_act || (!a->save_th && !a->save_sg), "resblock_fwd_tc: save_th / save_sg without save_act");
  const bool pair = (a->variant & 3) != 1;     // 0 / 2: CTA-pair kernel (default); 1: single-CTA kernel
  p.xflags = (a->variant >> 2) & 3;
  const bool prec = a->act_fmt == WNB200_ACT_F16X2;
  WNB_CHECK_ARG(a->act_fmt == WNB200_ACT_BF16 || prec, "resblock_fwd_tc: bad act_fmt %d", a->act_fmt);
  WNB_CHECK_ARG(!prec || (pair && !a->save_act), "resblock_fwd_tc: the fp16 (hi, lo) format needs the CTA-pair kernel, inference only");
  WNB_CHECK_ARG(!prec || !a->res || a->res_lo, "resblock_fwd_tc: the fp16 (hi, lo) format writes res AND res_lo");
  WNB_CHECK_ARG(prec || (!a->x_lo && !a->res_lo), "resblock_fwd_tc: x_lo / res_lo belong to the fp16 (hi, lo) format");
  p.has_lo = prec && a->x_lo != nullptr;
  const int wbox = pair ? C / 2 : C;
  CUtensorMap mx, mw1, mw2, mres, msk;
  int rc;
  if ((rc = rb_map_nlc(&mx, a->x, a->B, a->T, C, 2))) return rc;
  if ((rc = rb_map_2d(&mw1, a->w1, 2 * C, a->ntaps * C, wbox))) return rc;
  if ((rc = rb_map_2d(&mw2, a->w2, 2 * C, 2 * C, wbox))) return rc;
  if ((rc = rb_map_nlc(&mres, a->res ? a->res : a->x, a->B, a->T, C, 2))) return rc;
  if ((rc = rb_map_nlc(&msk, a->skips, a->B, a->T, C, 4))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pair) {
    // last layer: the `res` map carries the bf16 head input instead
    if (p.final_act && (rc = rb_map_nlc(&mres, a->skips_act, a->B, a->T, C, 2))) return rc;
    p.tiles_per_seq = ceil_div(a->T, 2 * RB_TILE);
    p.num_tiles = p.tiles_per_seq * a->B;
    if (a->save_act) {
      CUtensorMap mact, mth, msg;
      if ((rc = rb_map_nlc(&mact, a->save_act, a->B, a->T, C, 2))) return rc;
      mth = mact;
      if (a->save_th && (rc = rb_map_nlc(&mth, a->save_th, a->B, a->T, C, 2))) return rc;
      if ((rc = rb_map_nlc(&msg, a->save_sg, a->B, a->T, C, 2))) return rc;
      return C == 256 ? launch_resblock2<256, true>(mx, mw1, mw2, mres, msk, mact, mth, msg, p, st)
                      : launch_resblock2<128, true>(mx, mw1, mw2, mres, msk, mact, mth, msg, p, st);
    }
    if (prec) {
      CUtensorMap mxlo = mx, mreslo = mx;
      if (a->x_lo && (rc = rb_map_nlc(&mxlo, a->x_lo, a->B, a->T, C, 2))) return rc;
      if (a->res && (rc = rb_map_nlc(&mreslo, a->res_lo, a->B, a->T, C, 2))) return rc;
      return C == 256 ? launch_resblock2<256, false, true>(mx, mw1, mw2, mres, msk, mxlo, mreslo, mx, p, st)
                      : launch_resblock2<128, false, true>(mx, mw1, mw2, mres, msk, mxlo, mreslo, mx, p, st);
    }
    return C == 256 ? launch_resblock2<256, false>(mx, mw1, mw2, mres, msk, mx, mx, mx, p, st)
                    : launch_resblock2<128, false>(mx, mw1, mw2, mres, msk, mx, mx, mx, p, st);
  }
  return C == 256 ? launch_resblock<256>(mx, mw1, mw2, mres, msk, p, st)
                  : launch_resblock<128>(mx, mw1, mw2, mres, msk, p, st);
}

extern "C" int wnb200_leaky_to_bf16(int64_t n, const float* x, void* y, void* stream) {
  WNB_CHECK_ARG(n % 4 == 0, "leaky_to_bf16: n must be a multiple of 4");
  if (n == 0) return 0;
  WNB_CHECK_ARG(x && y, "leaky_to_bf16: null pointer");
  const long long n4 = n / 4;
  long long g = (n4 + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  leaky_to_bf16_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(n4, (const float4*)x, (uint2*)y);
  WNB_LAUNCH_OK();
  return 0;
}
