// Weight gradient on tensor cores (CTA pair), NLC bf16 operands:
//     dW[m, n] += sum_{b,t} G[b, t, m0 + m] * X[b, t + off, n]         m < 256, n < N (N = 128 or 256)
// i.e. the contraction runs over TIME, so both operands are read "transposed": G[t][m] has m contiguous, which is an
// MN-major A operand for tcgen05.mma (and likewise X as B).  A TMA box {64 channels, 64 frames} of the NLC tensor
// lands as 64 rows of 128 B with the 128-byte swizzle -- exactly the MN-major SWIZZLE_128B canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: 8-frame groups 1024 B apart (SBO), 64-channel blocks one box
// (8 KB) apart (LBO).  Out-of-range frames (t + off outside [0,T), tail of T) are zero-filled by TMA = the conv's
// zero padding.
//
// Split over time: the (batch, 64-frame) blocks are divided evenly over the CTA pairs; each pair accumulates its
// 256 x N partial in TMEM (CTA r of the pair owns output rows [128r, 128r+128) and loads half of X's channels) and
// adds it to dW with TMA reduce-add (fp32; the order of the cross-pair additions is not fixed).
// Replaces autograd's weight gradients of conv_ops.py:43,78 / block.py:73-78 on the training step.
//
// One launch carries up to WG_MAXJOBS such products ("jobs") over the SAME frames -- all weight gradients of one residual
// block: [dtanh ; dsigmoid] x the taps of x (two 256-row jobs), dres x (gate, x), dskips x gate.  The CTA pairs are
// divided over the jobs in proportion to their MMA work, and a pair walks the (batch, 64-frame) blocks with the stride
// of its job's pair count, so every job sweeps the frames front to back at the same pace: x, the gate and the
// gradients are fetched from HBM once per launch and found in L2 by the other jobs (as four separate launches the
// block's weight gradients read 2.95 GB at the config-3 shape, 1.6 GB of them distinct).
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace wnb {
using namespace tc;

constexpr int WG_MAXJOBS = 6;
struct WgJob {
  CUtensorMap map_g, map_x, map_x2, map_dw;
  int m0, N, off;
  int nsrc, off2;        // optional second X tensor (same N): its dW block follows the first one's columns
  int pair_begin, npairs;
  int pad_[9];
};
struct WgDev {
  int T, kblocks_per_seq, total_kblocks, njobs;
  int pad_[12];
  WgJob job[WG_MAXJOBS];
};
static_assert(sizeof(WgJob) % 64 == 0 && sizeof(WgDev) <= 4000, "kernel parameter layout");

constexpr int WG_THREADS = 192;                  // TMA warp, MMA warp, 4 epilogue warps
constexpr int WG_BOX = 64 * 128;                 // one {64 ch, 64 frames} box = 8 KB
constexpr int WG_STAGE = 6 * WG_BOX;             // A: 2 boxes (128 rows of dW), B: up to 4 boxes (this CTA's share of X, X2)
constexpr int WG_NSTAGE = 4;
constexpr int WG_STAGING = RB_TILE * 128;
constexpr int WG_SMEM = WG_NSTAGE * WG_STAGE + WG_STAGING + 1024 + 256;

// MN-major, 128-byte swizzle: LBO = distance between 64-element blocks along M/N, SBO = between 8-row groups along K
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WG_THREADS, 1)
wgrad2_kernel(const __grid_constant__ WgDev P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stg_base = smem_base + WG_NSTAGE * WG_STAGE;
  const uint32_t bar_base = stg_base + WG_STAGING;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WG_NSTAGE + s); };
  const uint32_t acc_full = bar_base + 8u * (2 * WG_NSTAGE);
  const uint32_t tmem_slot = acc_full + 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  int ji = 0;
  for (int j = 1; j < P.njobs; ++j)
    if (pair >= P.job[j].pair_begin) ji = j;
  const WgJob& p = P.job[ji];
  const CUtensorMap& map_g = p.map_g;
  const CUtensorMap& map_x = p.map_x;
  const CUtensorMap& map_x2 = p.map_x2;
  const CUtensorMap& map_dw = p.map_dw;
  // this pair's (batch, 64-frame) blocks: kb_begin, kb_begin + npairs, ... (strided: all pairs of all jobs move through
  // the frames together)
  const int kb_begin = pair - p.pair_begin, kb_step = p.npairs;
  const int nkb = kb_begin < P.total_kblocks ? (P.total_kblocks - kb_begin + kb_step - 1) / kb_step : 0;
  const int nhalf = p.N / 2;                                  // X channels held by this CTA
  const int xboxes = nhalf / 64;                              // boxes per X tensor per CTA
  const uint32_t stage_bytes = (uint32_t)(2 * WG_BOX + p.nsrc * xboxes * WG_BOX);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_g);
    prefetch_tensormap(&map_x);
    if (p.nsrc > 1) prefetch_tensormap(&map_x2);
    prefetch_tensormap(&map_dw);
    for (int s = 0; s < WG_NSTAGE; ++s) {
      mbar_init(full_bar(s), 2);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < P.total_kblocks; kb += kb_step) {
        const int b = kb / P.kblocks_per_seq;
        const int t0 = (kb - b * P.kblocks_per_seq) * 64;
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sa = smem_base + stage * WG_STAGE;
        if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * stage_bytes);
        const uint32_t lfull = mapa_shared(full_bar(stage), 0);
        // A: this CTA's 128 output rows = 2 blocks of 64 channels of G
        for (int i = 0; i < 2; ++i)
          tma_load_3d_2sm(sa + i * WG_BOX, &map_g, lfull, p.m0 + (int)rank * 128 + i * 64, t0, b);
        // B: this CTA's half of X's channels
        for (int i = 0; i < xboxes; ++i)
          tma_load_3d_2sm(sa + (2 + i) * WG_BOX, &map_x, lfull, (int)rank * nhalf + i * 64, t0 + p.off, b);
        if (p.nsrc > 1)
          for (int i = 0; i < xboxes; ++i)
            tma_load_3d_2sm(sa + (2 + xboxes + i) * WG_BOX, &map_x2, lfull, (int)rank * nhalf + i * 64, t0 + p.off2, b);
        if (rank != 0) mbar_arrive_cluster(lfull);
        if (++stage == WG_NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0 && nkb > 0) {
      int stage = 0;
      uint32_t phase = 0;
      // M = 256 (pair), N, both operands MN-major (bits 15 and 16)
      const uint32_t idesc = make_idesc_bf16(2 * RB_TILE, p.N) | (1u << 15) | (1u << 16);
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * WG_STAGE;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {   // 64 frames = 4 K-steps of 16 frames = 2 eight-row groups each
          const uint64_t da = make_smem_desc_mn_sw128(sa + k4 * 2048, WG_BOX);
          const uint32_t acc = (i == 0 && k4 == 0) ? 0u : 1u;
          umma_bf16_2sm(tmem_base, da, make_smem_desc_mn_sw128(sa + 2 * WG_BOX + k4 * 2048, WG_BOX), idesc, acc);
          if (p.nsrc > 1)
            umma_bf16_2sm(tmem_base + (uint32_t)p.N, da,
                          make_smem_desc_mn_sw128(sa + (2 + xboxes) * WG_BOX + k4 * 2048, WG_BOX), idesc, acc);
        }
        umma_commit_2sm(empty_bar(stage));
        if (++stage == WG_NSTAGE) { stage = 0; phase ^= 1; }
      }
      umma_commit_2sm(acc_full);
    }
  } else if (nkb > 0) {
    // epilogue: 4 warps, thread = one row of dW (128 rows per CTA), 32 columns per TMA reduce-add
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int sw = row & 7;
    const bool issuer = (threadIdx.x == 64);
    mbar_wait(acc_full, 0u);
    tc_fence_after();
    // All MMAs have completed (acc_full), so the operand ring is idle: its 12 x 16 KB slots (4 stages x 48 KB) serve
    // as staging buffers, one per 32-column chunk -- the TMA reduce-adds of up to 12 chunks are in flight while the
    // next chunks are drained from TMEM (a single staging buffer serialised every chunk behind the previous TMA read:
    // ~24 us per launch for the 256 x 512 tile).
    constexpr int NSLOT = WG_NSTAGE * WG_STAGE / WG_STAGING;      // 12
    for (int c = 0; c < p.nsrc * p.N / 32; ++c) {
      float a[32];
      tmem_ld16(tmem_base + lane_off + c * 32, a);
      tmem_ld16(tmem_base + lane_off + c * 32 + 16, a + 16);
      tmem_wait_ld();
      const uint32_t slot = smem_base + (uint32_t)(c % NSLOT) * WG_STAGING;
      if (c >= NSLOT) {                                           // the slot's previous reduce-add must have read it
        if (issuer) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NSLOT - 1) : "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      uint8_t* srow = smem_gen + (slot - smem_base) + row * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(srow + ((j ^ sw) << 4)) = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (issuer) {
        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&map_dw)),
                     "r"(slot), "r"(c * 32), "r"((int)rank * 128)
                     : "memory");
        bulk_commit();
      }
    }
    if (issuer) bulk_wait0();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// NLC [B][T][C] bf16 -> boxes [1][64 frames][64 channels], 128B swizzle
static int wg_map_nlc64(CUtensorMap* m, const void* ptr, int B, int T, int C) {
  EncodeTiledFn enc = rb_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return 5; }
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2};
  cuuint32_t box[3] = {64, 64, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(nlc64) failed: %d", (int)r); return 5; }
  return 0;
}

// dW fp32 [rows][N] row-major -> boxes [128 rows][32 columns], 128B swizzle
static int wg_map_dw(CUtensorMap* m, const void* ptr, int rows, int N) {
  EncodeTiledFn enc = rb_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return 5; }
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)N * 4};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(dw) failed: %d", (int)r); return 5; }
  return 0;
}

}  // namespace wnb

using namespace wnb;

static int wg_launch(int B, int T_, int njobs, const wnb200_wgrad_job_t* jobs, void* stream) {
  WNB_CHECK_ARG(njobs >= 1 && njobs <= WG_MAXJOBS, "wgrad_jobs_tc: %d jobs (1..%d per launch)", njobs, WG_MAXJOBS);
  WNB_CHECK_ARG(jobs != nullptr, "wgrad_jobs_tc: null job list");
  if (B == 0 || T_ == 0) return 0;
  WgDev P;
  memset(&P, 0, sizeof(P));
  P.T = T_;
  P.kblocks_per_seq = ceil_div(T_, 64);
  P.total_kblocks = P.kblocks_per_seq * B;
  P.njobs = njobs;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int pairs = sms / 2;
  if ((long long)P.total_kblocks * njobs < pairs) pairs = P.total_kblocks * njobs;
  if (pairs < njobs) pairs = njobs;
  int wsum = 0;
  for (int j = 0; j < njobs; ++j) {
    const wnb200_wgrad_job_t& a = jobs[j];
    WNB_CHECK_ARG(a.N == 128 || a.N == 256, "wgrad_tc: N=%d must be 128 or 256", a.N);
    WNB_CHECK_ARG(a.nsrc == 1 || a.nsrc == 2, "wgrad_tc: nsrc=%d must be 1 or 2", a.nsrc);
    WNB_CHECK_ARG(a.m0 >= 0 && a.m0 < a.Cg && a.Cg % 8 == 0, "wgrad_tc: row offset %d outside the %d channels of g", a.m0, a.Cg);
    WNB_CHECK_ARG(a.g && a.x && a.dw && (a.nsrc == 1 || a.x2), "wgrad_tc: null pointer");
    wsum += a.nsrc;
  }
  // pairs per job in proportion to the MMA work (one 256 x N product per source and 64 frames); the remainder goes to the
  // first jobs
  int given = 0;
  for (int j = 0; j < njobs; ++j) {
    int n = pairs * jobs[j].nsrc / wsum;
    if (n < 1) n = 1;
    P.job[j].npairs = n;
    given += n;
  }
  for (int j = 0; given < pairs; j = (j + 1) % njobs) { ++P.job[j].npairs; ++given; }
  for (int j = njobs - 1; given > pairs; j = (j + njobs - 1) % njobs)
    if (P.job[j].npairs > 1) { --P.job[j].npairs; --given; }
  int begin = 0;
  for (int j = 0; j < njobs; ++j) {
    const wnb200_wgrad_job_t& a = jobs[j];
    WgJob& q = P.job[j];
    q.pair_begin = begin;
    begin += q.npairs;
    q.m0 = a.m0; q.N = a.N; q.off = a.off[0];
    q.nsrc = a.nsrc; q.off2 = a.nsrc > 1 ? a.off[1] : 0;
    int rc;
    // rows beyond Cg are zero-filled by TMA (their dW rows receive zeros)
    if ((rc = wg_map_nlc64(&q.map_g, a.g, B, T_, a.Cg))) return rc;
    if ((rc = wg_map_nlc64(&q.map_x, a.x, B, T_, a.N))) return rc;
    q.map_x2 = q.map_x;
    if (a.nsrc > 1 && (rc = wg_map_nlc64(&q.map_x2, a.x2, B, T_, a.N))) return rc;
    if ((rc = wg_map_dw(&q.map_dw, a.dw, 256, a.nsrc * a.N))) return rc;
  }
  WNB_SET_SMEM_ATTR(WG_SMEM, wgrad2_kernel);
  wgrad2_kernel<<<2 * begin, WG_THREADS, WG_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(P);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_wgrad_jobs_tc(int B, int T_, int njobs, const wnb200_wgrad_job_t* jobs, void* stream) {
  return wg_launch(B, T_, njobs, jobs, stream);
}

extern "C" int wnb200_wgrad2_tc(int B, int T_, int Cg, int m0, int N, int nsrc, const int32_t* off /*host*/,
                                const void* g_nlc, const void* x_nlc, const void* x2_nlc, float* dw, void* stream) {
  WNB_CHECK_ARG(off != nullptr, "wgrad_tc: null pointer");
  wnb200_wgrad_job_t a;
  memset(&a, 0, sizeof(a));
  a.g = g_nlc; a.Cg = Cg; a.m0 = m0; a.x = x_nlc; a.x2 = x2_nlc; a.N = N; a.nsrc = nsrc;
  a.off[0] = off[0]; a.off[1] = nsrc > 1 ? off[1] : 0;
  a.dw = dw;
  return wg_launch(B, T_, 1, &a, stream);
}

extern "C" int wnb200_wgrad_tc(int B, int T_, int Cg, int m0, int N, int off, const void* g_nlc, const void* x_nlc,
                               float* dw, void* stream) {
  const int32_t offs[2] = {off, 0};
  return wnb200_wgrad2_tc(B, T_, Cg, m0, N, 1, offs, g_nlc, x_nlc, nullptr, dw, stream);
}
