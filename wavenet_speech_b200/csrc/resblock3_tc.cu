// Fused residual block, "deferred skip" variant (inference; C = 128 or 256; CTA pair, cta_group::2).
//
// resblock2_kernel (resblock_tc.cu) adds every layer's bottleneck output into a running fp32 skip sum in HBM: 1.07 GB
// of its 1.55 GB (bf16 format) / 2.09 GB (fp16 (hi, lo) format) of DRAM traffic per launch, a fourth contraction (G2s)
// and a fourth epilogue (E2b) per tile.  The skip sum of a stack is linear in the gates,
//
//     skips = sum_l (Wbn_l Wskip_l) gate_l + sum_l (Wbn_l bskip_l + bbn_l)                    (wavenet.py:97-100),
//
// so this kernel only STORES its gate (NLC, straight from the shared tile that feeds the residual contraction -- one
// 2-byte tensor instead of the fp32 read-modify-write) into layer l's slot of a gate stack, and after the last layer ONE
// contraction over K = layers x channels (wnb200_dense_fwd_tc with `nlayers`) accumulates the whole sum in tensor
// memory: the running skip sum never exists in HBM -- the north star's "the skip sum accumulates in TMEM across layers".
//
// Per 128-frame tile a CTA now runs three contractions through two TMEM regions that swap roles every tile:
//     G1a  x taps * W1[half 0]        -> R1        E1a: gate channels [0, C/2)  -> gate tile (smem)
//     G1b  x taps * W1[half 1]        -> R2        E1b: gate channels [C/2, C)  -> gate tile, then TMA store of the gate
//     G2r  x(t)*Wproj (+ x_lo*Wproj) + gate*Wres -> R1      E2a: res (+ bias) -> staging -> TMA store (hi [, lo])
// with R1(tile i+1) = R2(tile i): the next tile's G1a starts in the region E1b has just drained, and E2a -- which held
// the next G1a back for 3.6k (bf16) / 4.9k (fp16 pair) cycles per tile in resblock2_kernel -- has the whole of G1a to
// finish before G1b needs its region.  The last layer of a stack has no residual output: G2r / E2a are skipped.
//
// Roles as in resblock2_kernel: warp 0 TMA producer, warp 1 MMA issue (leader CTA), warps 2..9 epilogue.
// Reference semantics: modules/block.py:54-82 (the skip branch of :74 is the stored gate; wavenet.py:100 happens in the
// stack-wide contraction).
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace wnb {
using namespace tc;
typedef __nv_bfloat16 bf16;

struct Res3Dev {
  int B, T, tiles_per_seq, num_tiles;
  int ntaps, t_off[3];
  const float* bias1;   // [2C] packed like W1 rows (pre-scaled in the fp16 format)
  const float* bias2;   // [>= C] bres + bproj
  int write_res, has_lo, write_lo;   // write_lo: the fp16 format's output carries its lo half (else hi only)
  int* sat_flag;        // optional: set to 1 when a stream value left the fp16 range (the hi half saturated)
  bf16* sg_out;         // SAVE (training, bf16 format): NLC [B,T,C] that receives sigmoid(.) -- backward's second factor
};

constexpr int R3_THREADS = 320;
constexpr int R3_EPI_THREADS = 256;
constexpr int R3_ABYTES = RB_TILE * 128;

template <int C>
struct R3Cfg {
  static constexpr int KB = C / 64;
  static constexpr int BHBYTES = (C / 2) * 128;
  static constexpr int STAGE = R3_ABYTES + BHBYTES;
  static constexpr int ACT = KB * R3_ABYTES;
  static constexpr int STAGING = 2 * R3_ABYTES;
  static constexpr int NSTAGE = (C == 256) ? 4 : 6;
  static constexpr bool RETAIN = (NSTAGE == KB);
  static constexpr int SMEM = NSTAGE * STAGE + ACT + STAGING + 1024 + 256;
};

template <int C, bool F16>
__device__ __forceinline__ void mma3_kblock(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, bool first) {
  constexpr uint32_t idesc = make_idesc(F16, 2 * RB_TILE, C);
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4)
    umma_bf16_2sm(tmem_d, make_smem_desc_sw128(a_addr + k4 * 32), make_smem_desc_sw128(b_addr + k4 * 32), idesc,
                  (first && k4 == 0) ? 0u : 1u);
}

template <int C, bool PREC, bool SAVE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(R3_THREADS, 1)
resblock3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_res,
                 const __grid_constant__ CUtensorMap map_gate, const __grid_constant__ CUtensorMap map_xlo,
                 const __grid_constant__ CUtensorMap map_reslo, const Res3Dev p) {
  using K = R3Cfg<C>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t act_base = smem_base + K::NSTAGE * K::STAGE;
  const uint32_t stg_base = act_base + K::ACT;
  const uint32_t bar_base = stg_base + K::STAGING;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (K::NSTAGE + s); };
  const uint32_t bb = bar_base + 8u * (2 * K::NSTAGE);
  const uint32_t acc1_full = bb, acc2_full = bb + 8, e1a_done = bb + 16, e1b_done = bb + 24, e2a_done = bb + 32,
                 tmem_slot = bb + 48;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_w1);
    prefetch_tensormap(&map_w2);
    prefetch_tensormap(&map_gate);
    if (p.write_res) prefetch_tensormap(&map_res);
    if (PREC) {
      prefetch_tensormap(&map_xlo);
      prefetch_tensormap(&map_reslo);
    }
    for (int s = 0; s < K::NSTAGE; ++s) {
      mbar_init(full_bar(s), 2);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(acc2_full, 1);
    mbar_init(e1a_done, 2 * R3_EPI_THREADS);
    mbar_init(e1b_done, 2 * R3_EPI_THREADS);
    mbar_init(e2a_done, 2 * R3_EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, 2 * C);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto next = [&]() {
        if (++stage == K::NSTAGE) { stage = 0; phase ^= 1; }
      };
      const int wrow = (int)rank * (C / 2);
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
        // G1a walks the taps as ord[0..ntaps), G1b backwards; with a ring exactly KB stages deep the activation blocks of
        // the tap a phase starts with are still in the stages it is about to use: only the weight half is reloaded
        for (int half = 0; half < 2; ++half) {
          for (int jj = 0; jj < p.ntaps; ++jj) {
            const int tap = half == 0 ? ord[jj] : ord[p.ntaps - 1 - jj];
            const bool keep = K::RETAIN && half == 1 && jj == 0;
            for (int cb = 0; cb < K::KB; ++cb) {
              const uint32_t sa = smem_base + stage * K::STAGE;
              const uint32_t lfull = begin_stage(keep ? K::BHBYTES : K::STAGE);
              if (!keep) tma_load_3d_2sm(sa, &map_x, lfull, cb * 64, t0 + p.t_off[tap], b);
              tma_load_2d_2sm(sa + R3_ABYTES, &map_w1, lfull, (tap * K::KB + cb) * 64, half * C + wrow);
              end_stage(lfull);
            }
          }
        }
        if (!p.write_res) continue;               // last layer of a stack: no residual output
        const bool keep_x = K::RETAIN && p.t_off[ord[0]] == 0;
        for (int kb = 0; kb < K::KB; ++kb) {      // x(t) * Wproj
          const uint32_t sa = smem_base + stage * K::STAGE;
          const uint32_t lfull = begin_stage(keep_x ? K::BHBYTES : K::STAGE);
          if (!keep_x) tma_load_3d_2sm(sa, &map_x, lfull, kb * 64, t0, b);
          tma_load_2d_2sm(sa + R3_ABYTES, &map_w2, lfull, C + kb * 64, wrow);
          end_stage(lfull);
        }
        if (PREC && p.has_lo) {                   // the stream's low half against the retained Wproj blocks
          for (int kb = 0; kb < K::KB; ++kb) {
            const uint32_t sa = smem_base + stage * K::STAGE;
            const uint32_t lfull = begin_stage(K::RETAIN ? R3_ABYTES : K::STAGE);
            tma_load_3d_2sm(sa, &map_xlo, lfull, kb * 64, t0, b);
            if (!K::RETAIN) tma_load_2d_2sm(sa + R3_ABYTES, &map_w2, lfull, C + kb * 64, wrow);
            end_stage(lfull);
          }
        }
        for (int kb = 0; kb < K::KB; ++kb) {      // gate * Wres: weights only
          const uint32_t sa = smem_base + stage * K::STAGE;
          const uint32_t lfull = begin_stage(K::BHBYTES);
          tma_load_2d_2sm(sa + R3_ABYTES, &map_w2, lfull, kb * 64, wrow);
          end_stage(lfull);
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
      uint32_t n_e1a = 0, n_e1b = 0, n_e2a = 0;   // epilogue phases consumed so far (parity = count & 1)
      int it = 0;
      for (int pt = pair; pt < p.num_tiles; pt += npairs, ++it) {
        const uint32_t r1 = tmem_base + ((it & 1) ? C : 0), r2 = tmem_base + ((it & 1) ? 0 : C);
        // ---- G1a -> R1 = the previous tile's R2: drained by its E1b
        if (it > 0 && n_e1b < (uint32_t)it) {
          mbar_wait(e1b_done, n_e1b & 1u);
          ++n_e1b;
          tc_fence_after();
        }
        for (int kb = 0; kb < p.ntaps * K::KB; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * K::STAGE;
          mma3_kblock<C, PREC>(r1, sa, sa + R3_ABYTES, kb == 0);
          umma_commit_2sm(empty_bar(stage));
          next();
        }
        umma_commit_2sm(acc1_full);
        // ---- G1b -> R2 = the previous tile's R1: drained by its E2a (or by its E1a when nothing followed G1a)
        if (it > 0) {
          if (p.write_res) {
            mbar_wait(e2a_done, n_e2a & 1u);
            ++n_e2a;
          } else {
            mbar_wait(e1a_done, n_e1a & 1u);
            ++n_e1a;
          }
          tc_fence_after();
        }
        for (int kb = 0; kb < p.ntaps * K::KB; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * K::STAGE;
          mma3_kblock<C, PREC>(r2, sa, sa + R3_ABYTES, kb == 0);
          umma_commit_2sm(empty_bar(stage));
          next();
        }
        umma_commit_2sm(acc2_full);
        if (!p.write_res) continue;
        // ---- G2r -> R1 (E1a has drained it): x(t) * Wproj [+ x_lo * Wproj] + gate * Wres
        mbar_wait(e1a_done, n_e1a & 1u);
        ++n_e1a;
        tc_fence_after();
        const int nproj = K::KB * ((PREC && p.has_lo) ? 2 : 1);
        for (int kb = 0; kb < nproj; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * K::STAGE;
          mma3_kblock<C, PREC>(r1, sa, sa + R3_ABYTES, kb == 0);
          umma_commit_2sm(empty_bar(stage));
          next();
        }
        for (int kb = 0; kb < K::KB; ++kb) {
          if (kb == K::KB / 2) {                  // second half of the gate: after E1b
            mbar_wait(e1b_done, n_e1b & 1u);
            ++n_e1b;
            tc_fence_after();
          }
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * K::STAGE;
          mma3_kblock<C, PREC>(r1, act_base + kb * R3_ABYTES, sa + R3_ABYTES, false);
          umma_commit_2sm(empty_bar(stage));
          next();
        }
        umma_commit_2sm(acc1_full);
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
    const uint32_t r_e1a = mapa_shared(e1a_done, 0), r_e1b = mapa_shared(e1b_done, 0), r_e2a = mapa_shared(e2a_done, 0);
    uint32_t n_acc1 = 0, n_acc2 = 0, nchunk = 0;
    int it = 0;
    for (int pt = pair; pt < p.num_tiles; pt += npairs, ++it) {
      const int b = pt / p.tiles_per_seq;
      const int t0 = (pt - b * p.tiles_per_seq) * (2 * RB_TILE) + (int)rank * RB_TILE;
      const uint32_t r1 = tmem_base + ((it & 1) ? C : 0), r2 = tmem_base + ((it & 1) ? 0 : C);

      if (!p.write_res && it > 0) {               // (with a residual output E2a's own waits cover this)
        if (issuer) bulk_wait_read0();            // the previous tile's gate store has left the gate tile
        epi_bar();
      }
      // ---------------- E1a / E1b: gate -> gate tile (2-byte, swizzled K-major blocks) ----------------
      for (int half = 0; half < 2; ++half) {
        float4 nb_t[4], nb_s[4];                  // the next step's biases are fetched while this one is worked on
        {
          const float4* bt = reinterpret_cast<const float4*>(p.bias1 + half * C + h * (C / 4));
          const float4* bs = reinterpret_cast<const float4*>(p.bias1 + half * C + C / 2 + h * (C / 4));
#pragma unroll
          for (int j = 0; j < 4; ++j) { nb_t[j] = __ldg(bt + j); nb_s[j] = __ldg(bs + j); }
        }
        if (half == 0) { mbar_wait(acc1_full, n_acc1 & 1u); ++n_acc1; }
        else { mbar_wait(acc2_full, n_acc2 & 1u); ++n_acc2; }
        tc_fence_after();
        const uint32_t treg = (half == 0 ? r1 : r2) + lane_off;
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
          if constexpr (PREC) {                   // bias1 pre-scaled: tanh rows by -2 log2(e), sigmoid rows by -log2(e)
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float v0 = gate_precise(fmaf(ac[i], -2.885390081777927f, bta[i]), fmaf(gc[i], -1.4426950408889634f, bsa[i]));
              const float v1 = gate_precise(fmaf(ac[i + 1], -2.885390081777927f, bta[i + 1]),
                                            fmaf(gc[i + 1], -1.4426950408889634f, bsa[i + 1]));
              pk[i >> 1] = pack_f16x2(v0, v1);
            }
          } else if constexpr (SAVE) {
            // training: the sigmoid factor goes straight from registers to HBM (32 contiguous bytes per thread and step;
            // no shared staging is left in this kernel, and a register copy-out measured the same as a TMA store)
            uint32_t ps[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float s0 = sigmoid_approx(gc[i] + bsa[i]), s1 = sigmoid_approx(gc[i + 1] + bsa[i + 1]);
              pk[i >> 1] = pack_bf16x2(tanh_approx(ac[i] + bta[i]) * s0, tanh_approx(ac[i + 1] + bta[i + 1]) * s1);
              ps[i >> 1] = pack_bf16x2(s0, s1);
            }
            if (t0 + row < p.T) {
              uint4* dst = reinterpret_cast<uint4*>(p.sg_out + ((size_t)b * p.T + t0 + row) * C + half * (C / 2) + col);
              dst[0] = make_uint4(ps[0], ps[1], ps[2], ps[3]);
              dst[1] = make_uint4(ps[4], ps[5], ps[6], ps[7]);
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
          uint8_t* blk = smem_gen + (act_base - smem_base) + kb * R3_ABYTES + row * 128;
          *reinterpret_cast<uint4*>(blk + ((ci ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(blk + (((ci + 1) ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive_cluster(half == 0 ? r_e1a : r_e1b);
      }
      // ---------------- the gate leaves for the stack-wide skip contraction: TMA store from the gate tile ----------------
      epi_bar();                                  // every warp's part of the tile is written (and fenced)
      if (issuer) {
#pragma unroll
        for (int kb = 0; kb < K::KB; ++kb) tma_store_3d(&map_gate, act_base + kb * R3_ABYTES, kb * 64, t0, b);
        bulk_commit();
      }
      if (!p.write_res) continue;

      // ---------------- E2a: res = acc + bias -> staging -> TMA store (fp16 format: hi and lo) ----------------
      float4 nbv[8];
      {
        const float4* bp = reinterpret_cast<const float4*>(p.bias2 + h * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) nbv[j] = __ldg(bp + j);
      }
      mbar_wait(acc1_full, n_acc1 & 1u);
      ++n_acc1;
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < C / 64; ++c, ++nchunk) {
        const int col = c * 64 + h * 32;
        float a[32];
        tmem_ld16(r1 + lane_off + col, a);
        tmem_ld16(r1 + lane_off + col + 16, a + 16);
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
        if (c == C / 64 - 1) {                    // the region has been read: the next tile's G1b may overwrite it
          tc_fence_before();
          mbar_arrive_cluster(r_e2a);
        }
        uint32_t pk[16];
        bool pair_out = false;
        if constexpr (PREC) pair_out = p.write_lo != 0;
        if (pair_out) {
          uint32_t pl[16];
          uint32_t sat = 0;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            split_f16x2(a[i] + bv[i], a[i + 1] + bv[i + 1], pk[i >> 1], pl[i >> 1]);
            // |hi| = 65504 (0x7bff): the value was at or beyond the fp16 range (cvt.satfinite clamps instead of inf)
            const uint32_t m = pk[i >> 1] & 0x7fff7fffu;
            sat |= (uint32_t)((m & 0xffffu) == 0x7bffu) | (uint32_t)((m >> 16) == 0x7bffu);
          }
          if (sat && p.sat_flag) atomicOr(p.sat_flag, 1);
          if (issuer) bulk_wait_read0();
          epi_bar();
          uint8_t* srow = smem_gen + (stg_base - smem_base) + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int o = ((4 * h + j) ^ sw) << 4;
            *reinterpret_cast<uint4*>(srow + o) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            *reinterpret_cast<uint4*>(srow + R3_ABYTES + o) =
                make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
          }
          fence_proxy_async_smem();
          epi_bar();
          if (issuer) {
            tma_store_3d(&map_res, stg_base, c * 64, t0, b);
            tma_store_3d(&map_reslo, stg_base + R3_ABYTES, c * 64, t0, b);
            bulk_commit();
          }
        } else {
          if constexpr (PREC) {                   // fp16 format, hi half only (the tail of a deep stack, fastpath.py)
            uint32_t sat = 0;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              pk[i >> 1] = pack_f16x2_sat(a[i] + bv[i], a[i + 1] + bv[i + 1]);
              const uint32_t m = pk[i >> 1] & 0x7fff7fffu;
              sat |= (uint32_t)((m & 0xffffu) == 0x7bffu) | (uint32_t)((m >> 16) == 0x7bffu);
            }
            if (sat && p.sat_flag) atomicOr(p.sat_flag, 1);
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) pk[i >> 1] = pack_bf16x2(a[i] + bv[i], a[i + 1] + bv[i + 1]);
          }
          const uint32_t boff = (nchunk & 1u) * R3_ABYTES;
          // at most the latest group may still be reading: that is either the other buffer's chunk or the gate store
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
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
    }
    if (issuer) bulk_wait0();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * C);
  }
}

template <int C, bool PREC, bool SAVE>
static int launch_resblock3(const CUtensorMap& mx, const CUtensorMap& mw1, const CUtensorMap& mw2,
                            const CUtensorMap& mres, const CUtensorMap& mgate, const CUtensorMap& mxlo,
                            const CUtensorMap& mreslo, const Res3Dev& p, cudaStream_t st) {
  using K = R3Cfg<C>;
  WNB_SET_SMEM_ATTR(K::SMEM, (resblock3_kernel<C, PREC, SAVE>));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int pairs = sms / 2;
  if (p.num_tiles < pairs) pairs = p.num_tiles;
  resblock3_kernel<C, PREC, SAVE><<<2 * pairs, R3_THREADS, K::SMEM, st>>>(mx, mw1, mw2, mres, mgate, mxlo, mreslo, p);
  WNB_LAUNCH_OK();
  return 0;
}

// Called by wnb200_resblock_fwd_tc (resblock_tc.cu) when the caller passes `gate_out`.
int resblock3_launch(const wnb200_resblock_t* a, void* stream) {
  const int C = a->C;
  const bool prec = a->act_fmt == WNB200_ACT_F16X2;
  WNB_CHECK_ARG(!a->save_act && !a->save_th && !a->skips_act,
                "resblock_fwd_tc: with gate_out the gate IS the saved activation; only save_sg may accompany it");
  WNB_CHECK_ARG(!a->save_sg || !prec, "resblock_fwd_tc: save_sg (training) belongs to the bf16 format");
  WNB_CHECK_ARG(prec || (!a->x_lo && !a->res_lo), "resblock_fwd_tc: x_lo / res_lo belong to the fp16 (hi, lo) format");
  // (fp16 format: res without res_lo = the output stream carries its hi half only)
  Res3Dev p;
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.T = a->T;
  p.tiles_per_seq = ceil_div(a->T, 2 * RB_TILE);
  p.num_tiles = p.tiles_per_seq * a->B;
  p.ntaps = a->ntaps;
  for (int j = 0; j < 3; ++j) p.t_off[j] = a->t_off[j];
  p.bias1 = a->bias1; p.bias2 = a->bias2;
  p.write_res = a->res != nullptr;
  p.has_lo = prec && a->x_lo != nullptr;
  p.write_lo = prec && a->res != nullptr && a->res_lo != nullptr;
  p.sat_flag = prec ? reinterpret_cast<int*>(a->sat_flag) : nullptr;
  p.sg_out = reinterpret_cast<bf16*>(a->save_sg);
  CUtensorMap mx, mw1, mw2, mres, mgate, mxlo, mreslo;
  int rc;
  if ((rc = rb_map_nlc(&mx, a->x, a->B, a->T, C, 2))) return rc;
  if ((rc = rb_map_2d(&mw1, a->w1, 2 * C, a->ntaps * C, C / 2))) return rc;
  if ((rc = rb_map_2d(&mw2, a->w2, 2 * C, 2 * C, C / 2))) return rc;
  if ((rc = rb_map_nlc(&mgate, a->gate_out, a->B, a->T, C, 2))) return rc;
  mres = mx; mxlo = mx; mreslo = mx;
  if (a->res && (rc = rb_map_nlc(&mres, a->res, a->B, a->T, C, 2))) return rc;
  if (p.has_lo && (rc = rb_map_nlc(&mxlo, a->x_lo, a->B, a->T, C, 2))) return rc;
  if (p.write_lo && (rc = rb_map_nlc(&mreslo, a->res_lo, a->B, a->T, C, 2))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (prec)
    return C == 256 ? launch_resblock3<256, true, false>(mx, mw1, mw2, mres, mgate, mxlo, mreslo, p, st)
                    : launch_resblock3<128, true, false>(mx, mw1, mw2, mres, mgate, mxlo, mreslo, p, st);
  if (p.sg_out)
    return C == 256 ? launch_resblock3<256, false, true>(mx, mw1, mw2, mres, mgate, mxlo, mreslo, p, st)
                    : launch_resblock3<128, false, true>(mx, mw1, mw2, mres, mgate, mxlo, mreslo, p, st);
  return C == 256 ? launch_resblock3<256, false, false>(mx, mw1, mw2, mres, mgate, mxlo, mreslo, p, st)
                  : launch_resblock3<128, false, false>(mx, mw1, mw2, mres, mgate, mxlo, mreslo, p, st);
}

}  // namespace wnb
