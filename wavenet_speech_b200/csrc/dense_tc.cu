// Dense channel contraction on tensor cores, CTA pair, NLC bf16 input:
//     y[b, t, :] = epi( W · [x[b, t+off_0, :] ; x[b, t+off_1, :] ; ...] + bias )          N <= 256 outputs
// Used for the WaveNet entry conv (wavenet.py:54,93), both 1x1 convs of the output heads (wavenet.py:67-71,
// raw_ctcnet.py:84-88, classifier.py:70-74) and RawCTCNet's feature 1x1 (raw_ctcnet.py:60).
//
// Same machinery as resblock2_kernel (cta_group::2, weights split across the pair, TMA-fed ring, multicast commits),
// but one accumulator per tile: TMEM holds two N-column regions that alternate between tiles, so the epilogue of
// tile i overlaps the MMAs of tile i+1.  Epilogues:
//   NLC : (+bias, optional LeakyReLU) -> bf16 -> swizzled staging -> TMA store into an NLC tensor
//   HEAD: (+bias, optional channel softmax, thread-local over TMEM columns + one cross-warp combine) ->
//         transposed staging [channel][frame] -> TMA store into the NCL output (or direct stores when the NCL
//         row pitch is not 16-byte aligned).
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace wnb {
using namespace tc;
typedef __nv_bfloat16 bf16;

// one fp32 value -> the 16-bit word of the activation format (bf16, or fp16 bits carried in a bf16-typed slot)
__device__ __forceinline__ bf16 act16(bool f16, float v) {
  if (f16) {
    const __half hv = __float2half_rn(v);
    return *reinterpret_cast<const bf16*>(&hv);
  }
  return __float2bfloat16_rn(v);
}

struct DenseDev {
  int B, T, tiles_per_seq, num_tiles;
  int ntaps, t_off[3];
  int kb_per_tap;   // Cin / 64
  int ntaps2, t_off2[3], kb_per_tap2;   // optional second source tensor (its taps follow the first one's in K)
  int N;            // accumulator columns (multiple of 16, <= 256)
  int mode;         // 0 NLC store, 1 HEAD
  int leaky;        // NLC: LeakyReLU(0.01) after bias
  int n_out, softmax, out_f32, tma_out;
  const float* bias;
  void* out;        // HEAD direct-store fallback: NCL [B, n_out, T]
  float* colsum;    // NLC mode, optional: fp32 [N] += column sums of the (bf16-rounded) output over all valid frames
  int f16;          // operands (and NLC output) are fp16 instead of bf16 (WNB200_ACT_F16X2)
  int split;        // NLC mode, fp16: the output leaves as an fp16 (hi, lo) pair (map_y, map_ylo)
  int nlayers;      // > 0: x is a stack [nlayers][B][T][Cin] and K runs over (layer, channel): y = sum_l W_l x_l (+ bias)
  const float* pos_w;    // NLC mode, optional: y += hardtanh(pos_w[c] * (pos_t0 + t) + pos_b[c]) after bias / LeakyReLU
  const float* pos_b;    // (RawCTCNet position mixing, raw_ctcnet.py:131-135)
  int pos_t0;
  const bf16* gb_gate;   // gate-backward epilogue (training, NLC mode, bf16): acc is d(gate); with the forward's gate and
  const bf16* gb_sg;     // sigmoid (NLC [B,T,N]) y[.., 0:N] = d(tanh pre-act), y[.., N:2N] = d(sigmoid pre-act); colsum [2N]
};

constexpr int DN_THREADS = 320;
constexpr int DN_ABYTES = RB_TILE * 128;
constexpr int DN_STAGE = 2 * DN_ABYTES;        // A block + up to 128 weight rows
constexpr int DN_NSTAGE = 5;
constexpr int DN_STAGING = 2 * DN_ABYTES;
constexpr int DN_SCRATCH = 2 * RB_TILE * 8;    // softmax partials (max, sum) per column half
constexpr int DN_BIAS = 256 * 4;                // head: bias (x log2 e under softmax) staged once per CTA
constexpr int DN_CSACC = 8 * 512 * 4;           // gate-backward epilogue: per-warp column sums of the 2N outputs
constexpr int DN_SMEM = DN_NSTAGE * DN_STAGE + DN_STAGING + DN_SCRATCH + DN_BIAS + DN_CSACC + 1024 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DN_THREADS, 1)
dense2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_x2,
              const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_y,
              const __grid_constant__ CUtensorMap map_ylo, const DenseDev p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stg_base = smem_base + DN_NSTAGE * DN_STAGE;
  const uint32_t scr_base = stg_base + DN_STAGING;
  const uint32_t sbias_base = scr_base + DN_SCRATCH;
  const uint32_t csacc_base = sbias_base + DN_BIAS;
  const uint32_t bar_base = csacc_base + DN_CSACC;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (DN_NSTAGE + s); };
  const uint32_t bb = bar_base + 8u * (2 * DN_NSTAGE);
  auto acc_full = [&](int r) { return bb + 8u * r; };
  auto acc_empty = [&](int r) { return bb + 16u + 8u * r; };
  const uint32_t tmem_slot = bb + 32;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nkb1 = (p.nlayers > 0 ? p.nlayers : p.ntaps) * p.kb_per_tap;
  const int nkb = nkb1 + p.ntaps2 * p.kb_per_tap2;
  const uint32_t bhalf_bytes = (uint32_t)(p.N / 2) * 128u;
  // the gate-backward epilogue stages two tiles per chunk and wants them double-buffered: it runs a 4-deep ring and
  // uses the fifth stage's 32 KB as its second staging pair (K = 2C there: eight K-blocks per tile)
  const int NST = p.gb_gate ? DN_NSTAGE - 1 : DN_NSTAGE;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_x);
    if (p.ntaps2 > 0) prefetch_tensormap(&map_x2);
    prefetch_tensormap(&map_w);
    prefetch_tensormap(&map_y);
    if (p.split) prefetch_tensormap(&map_ylo);
    for (int s = 0; s < DN_NSTAGE; ++s) {
      mbar_init(full_bar(s), 2);
      mbar_init(empty_bar(s), 1);
    }
    for (int r = 0; r < 2; ++r) {
      mbar_init(acc_full(r), 1);
      mbar_init(acc_empty(r), 2 * 256);
    }
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
      for (int pt = pair; pt < p.num_tiles; pt += npairs) {
        const int b = pt / p.tiles_per_seq;
        const int t0 = (pt - b * p.tiles_per_seq) * (2 * RB_TILE) + (int)rank * RB_TILE;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * DN_STAGE;
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * (DN_ABYTES + bhalf_bytes));
          const uint32_t lfull = mapa_shared(full_bar(stage), 0);
          if (kb < nkb1) {
            const int tap = kb / p.kb_per_tap, cb = kb - tap * p.kb_per_tap;
            if (p.nlayers > 0) tma_load_3d_2sm(sa, &map_x, lfull, cb * 64, t0, tap * p.B + b);     // tap = layer index
            else tma_load_3d_2sm(sa, &map_x, lfull, cb * 64, t0 + p.t_off[tap], b);
          } else {
            const int k2 = kb - nkb1;
            const int tap = k2 / p.kb_per_tap2, cb = k2 - tap * p.kb_per_tap2;
            tma_load_3d_2sm(sa, &map_x2, lfull, cb * 64, t0 + p.t_off2[tap], b);
          }
          tma_load_2d_2sm(sa + DN_ABYTES, &map_w, lfull, kb * 64, (int)rank * (p.N / 2));
          if (rank != 0) mbar_arrive_cluster(lfull);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc = p.f16 ? make_idesc_f16(2 * RB_TILE, p.N) : make_idesc_bf16(2 * RB_TILE, p.N);
      int it = 0;
      for (int pt = pair; pt < p.num_tiles; pt += npairs, ++it) {
        const int r = it & 1, use = it >> 1;
        if (use > 0) {
          mbar_wait(acc_empty(r), (uint32_t)((use - 1) & 1));
          tc_fence_after();
        }
        const uint32_t tacc = tmem_base + (uint32_t)r * 256u;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * DN_STAGE;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_bf16_2sm(tacc, make_smem_desc_sw128(sa + k4 * 32), make_smem_desc_sw128(sa + DN_ABYTES + k4 * 32),
                          idesc, (kb == 0 && k4 == 0) ? 0u : 1u);
          umma_commit_2sm(empty_bar(stage));
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(acc_full(r));
      }
    }
  } else {
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const bool issuer = (threadIdx.x == 64);
    const int sw = row & 7;
    float* scratch = reinterpret_cast<float*>(smem_gen + (scr_base - smem_base));
    float* sbias = reinterpret_cast<float*>(smem_gen + (sbias_base - smem_base));
    float* csacc = reinterpret_cast<float*>(smem_gen + (csacc_base - smem_base));
    {
      // bias staged once per CTA (head: pre-scaled by log2 e under softmax, zero past n_out); the gate-backward
      // epilogue also keeps its 2N column sums here (scratch) until the last tile
      const int t256 = threadIdx.x - 64;
      if (t256 < p.N) {
        if (p.mode == 1) sbias[t256] = (t256 < p.n_out) ? __ldg(p.bias + t256) * (p.softmax ? 1.4426950408889634f : 1.f) : 0.f;
        else sbias[t256] = __ldg(p.bias + t256);
      }
      if (p.gb_gate)
        for (int i = t256; i < DN_CSACC / 4; i += 256) csacc[i] = 0.f;
      epi_bar();
    }
    uint32_t nchunk = 0;
    float csum[4] = {0.f, 0.f, 0.f, 0.f};        // column sums: thread -> column (tid & 63) of each 64-channel chunk,
    const int cs_col = (threadIdx.x - 64) & 63;  // rows [32 rq, 32 rq + 32) of the tile
    const int cs_rq = (threadIdx.x - 64) >> 6;
    int it = 0;
    for (int pt = pair; pt < p.num_tiles; pt += npairs, ++it) {
      const int b = pt / p.tiles_per_seq;
      const int t0 = (pt - b * p.tiles_per_seq) * (2 * RB_TILE) + (int)rank * RB_TILE;
      const int r = it & 1, use = it >> 1;
      const uint32_t tacc = tmem_base + (uint32_t)r * 256u + lane_off;
      // gate-backward epilogue: this thread's 32 gate / sigmoid values of the coming chunk, fetched one chunk ahead
      // (64 contiguous bytes per tensor and row; the first chunk's loads are in flight while the contraction finishes)
      uint4 ng[4], ns[4];
      const bool gb_row = p.gb_gate != nullptr && t0 + row < p.T;
      const size_t gb_base = ((size_t)b * p.T + t0 + row) * (size_t)p.N + h * 32;
      auto gb_fetch = [&](int c) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ng[j] = gb_row ? __ldg(reinterpret_cast<const uint4*>(p.gb_gate + gb_base + c * 64) + j) : make_uint4(0, 0, 0, 0);
          ns[j] = gb_row ? __ldg(reinterpret_cast<const uint4*>(p.gb_sg + gb_base + c * 64) + j) : make_uint4(0, 0, 0, 0);
        }
      };
      if (p.gb_gate) gb_fetch(0);
      mbar_wait(acc_full(r), (uint32_t)(use & 1));
      tc_fence_after();

      const float pos_t = (float)(p.pos_t0 + t0 + row);       // this thread's frame, as the reference's float arange
      auto pos_term = [&](int ch) { return fminf(1.f, fmaxf(-1.f, fmaf(__ldg(p.pos_w + ch), pos_t, __ldg(p.pos_b + ch)))); };
      if (p.mode == 0) {
        // -------- (+bias, LeakyReLU [, + position term]) -> bf16 NLC, 64 channels per TMA store --------
        for (int c = 0; c < p.N / 64; ++c, ++nchunk) {
          const int col = c * 64 + h * 32;
          float a[32];
          tmem_ld16(tacc + col, a);
          tmem_ld16(tacc + col + 16, a + 16);
          const float4* bp = reinterpret_cast<const float4*>(sbias + col);
          float bv[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 u = bp[j];
            bv[4 * j] = u.x; bv[4 * j + 1] = u.y; bv[4 * j + 2] = u.z; bv[4 * j + 3] = u.w;
          }
          tmem_wait_ld();
          uint32_t pk[16];
          if (p.gb_gate) {
            // ---- gate backward (block.py:66-71): acc = d(gate); the forward kept gate = tanh * sigmoid and sigmoid
            uint32_t pl[16], wg[16], ws[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              wg[4 * j] = ng[j].x; wg[4 * j + 1] = ng[j].y; wg[4 * j + 2] = ng[j].z; wg[4 * j + 3] = ng[j].w;
              ws[4 * j] = ns[j].x; ws[4 * j + 1] = ns[j].y; ws[4 * j + 2] = ns[j].z; ws[4 * j + 3] = ns[j].w;
            }
            if (c + 1 < p.N / 64) gb_fetch(c + 1);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float gx = __uint_as_float(wg[i >> 1] << 16), gy = __uint_as_float(wg[i >> 1] & 0xffff0000u);
              const float sx = __uint_as_float(ws[i >> 1] << 16), sy = __uint_as_float(ws[i >> 1] & 0xffff0000u);
              // tanh = gate / sigmoid (0 / 0 -> 0); 1 - tanh^2 held at >= 0 against the rounding of the two bf16 factors;
              // tanh * sigmoid is the stored gate itself
              const float tx = gx * rcp_approx(fmaxf(sx, 1e-30f)), ty = gy * rcp_approx(fmaxf(sy, 1e-30f));
              const float ux = fmaxf(fmaf(-tx, tx, 1.f), 0.f), uy = fmaxf(fmaf(-ty, ty, 1.f), 0.f);
              const float d0 = a[i] + bv[i], d1 = a[i + 1] + bv[i + 1];
              pk[i >> 1] = pack_bf16x2(d0 * sx * ux, d1 * sy * uy);
              pl[i >> 1] = pack_bf16x2(d0 * gx * (1.f - sx), d1 * gy * (1.f - sy));
            }
            const uint32_t pbase = (nchunk & 1u) ? smem_base + (DN_NSTAGE - 1) * DN_STAGE : stg_base;
            if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            epi_bar();
            uint8_t* srow = smem_gen + (pbase - smem_base) + row * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int o = ((4 * h + j) ^ sw) << 4;
              *reinterpret_cast<uint4*>(srow + o) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
              *reinterpret_cast<uint4*>(srow + DN_ABYTES + o) =
                  make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
            }
            fence_proxy_async_smem();
            epi_bar();
            if (issuer) {
              tma_store_3d(&map_y, pbase, c * 64, t0, b);
              tma_store_3d(&map_y, pbase + DN_ABYTES, p.N + c * 64, t0, b);
              bulk_commit();
            }
            if (p.colsum) {
              // bias gradients = column sums of both staged tiles (rows past T hold zeros: their gate / sigmoid were
              // fetched as 0).  A thread sums one 16-byte unit (8 channels) over 4 rows, the 4 lanes that share a unit
              // meet by shuffle, and 8 lanes per warp add into that WARP's 2N shared accumulators (no atomics: a shared
              // fp32 atomicAdd is a compare-and-swap loop).
              const int t256 = threadIdx.x - 64, ci = t256 & 7, r0 = t256 >> 3;
              float s1[8], s2[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
              const uint8_t* rp = smem_gen + (pbase - smem_base) + r0 * 128 + ((ci ^ (r0 & 7)) << 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint4 u = *reinterpret_cast<const uint4*>(rp + k * 32 * 128);
                const uint4 v = *reinterpret_cast<const uint4*>(rp + DN_ABYTES + k * 32 * 128);
                const uint32_t uw[4] = {u.x, u.y, u.z, u.w}, vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  s1[2 * j] += __uint_as_float(uw[j] << 16); s1[2 * j + 1] += __uint_as_float(uw[j] & 0xffff0000u);
                  s2[2 * j] += __uint_as_float(vw[j] << 16); s2[2 * j + 1] += __uint_as_float(vw[j] & 0xffff0000u);
                }
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 8);  s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 8);  s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
              }
              if (lane < 8) {
                float4* q1 = reinterpret_cast<float4*>(csacc + (warp - 2) * 512 + c * 64 + ci * 8);
                float4* q2 = q1 + 64;              // + 256 floats: the second half of the output columns
                float4 u0 = q1[0], u1 = q1[1], v0 = q2[0], v1 = q2[1];
                u0.x += s1[0]; u0.y += s1[1]; u0.z += s1[2]; u0.w += s1[3];
                u1.x += s1[4]; u1.y += s1[5]; u1.z += s1[6]; u1.w += s1[7];
                v0.x += s2[0]; v0.y += s2[1]; v0.z += s2[2]; v0.w += s2[3];
                v1.x += s2[4]; v1.y += s2[5]; v1.z += s2[6]; v1.w += s2[7];
                q1[0] = u0; q1[1] = u1; q2[0] = v0; q2[1] = v1;
              }
            }
            continue;
          }
          if (p.split) {
            // fp16 (hi, lo) pair: both staging buffers per chunk, one bulk group
            uint32_t pl[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float v0 = a[i] + bv[i], v1 = a[i + 1] + bv[i + 1];
              if (p.leaky) { v0 = leaky(v0); v1 = leaky(v1); }
              if (p.pos_w) { v0 += pos_term(col + i); v1 += pos_term(col + i + 1); }
              split_f16x2(v0, v1, pk[i >> 1], pl[i >> 1]);
            }
            if (issuer) bulk_wait_read0();
            epi_bar();
            uint8_t* srow = smem_gen + (stg_base - smem_base) + row * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int o = ((4 * h + j) ^ sw) << 4;
              *reinterpret_cast<uint4*>(srow + o) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
              *reinterpret_cast<uint4*>(srow + DN_ABYTES + o) =
                  make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
            }
            fence_proxy_async_smem();
            epi_bar();
            if (issuer) {
              tma_store_3d(&map_y, stg_base, c * 64, t0, b);
              tma_store_3d(&map_ylo, stg_base + DN_ABYTES, c * 64, t0, b);
              bulk_commit();
            }
            continue;
          }
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float v0 = a[i] + bv[i], v1 = a[i + 1] + bv[i + 1];
            if (p.leaky) { v0 = leaky(v0); v1 = leaky(v1); }
            if (p.pos_w) { v0 += pos_term(col + i); v1 += pos_term(col + i + 1); }
            pk[i >> 1] = pack_act2(p.f16 != 0, v0, v1);
          }
          const uint32_t boff = (nchunk & 1u) * DN_ABYTES;
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
            tma_store_3d(&map_y, stg_base + boff, c * 64, t0, b);
            bulk_commit();
          }
          if (p.colsum) {
            // bias gradient of the NEXT contraction for free: the staged tile is summed over its valid frames while
            // the TMA store reads it (the buffer is rewritten two chunks and two barriers from now)
            const int nvalid = min(RB_TILE, p.T - t0) - cs_rq * 32;
            const uint8_t* sb = smem_gen + (stg_base - smem_base) + boff + (cs_rq * 32) * 128 + (cs_col & 7) * 2;
            const int ch16 = cs_col >> 3;
            float acc = 0.f;
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr)
              if (rr < nvalid)
                acc += __bfloat162float(*reinterpret_cast<const bf16*>(sb + rr * 128 + ((ch16 ^ (rr & 7)) << 4)));
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q == c) csum[q] += acc;
          }
        }
      } else {
        // -------- HEAD: optional channel softmax, output NCL --------
        // log2 domain: z = acc * log2(e) + bias * log2(e) (bias pre-scaled in shared memory), softmax = 2^(z - max) / sum.
        // Two cheap sweeps over the TMEM row (max, then sum) instead of an online update with a branch per element.
        constexpr float L2E = 1.4426950408889634f;
        const float sc = p.softmax ? L2E : 1.f;
        float mx = 0.f, inv = 1.f;
        auto ld_bias16 = [&](int c, float (&bb)[16]) {         // 16 consecutive (pre-scaled) biases: four 16-byte loads
          const float4* q4 = reinterpret_cast<const float4*>(sbias + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 u = q4[j];
            bb[4 * j] = u.x; bb[4 * j + 1] = u.y; bb[4 * j + 2] = u.z; bb[4 * j + 3] = u.w;
          }
        };
        if (p.softmax) {
          const int cbeg = h * (p.N / 2);
          const int cend = (cbeg + p.N / 2) < p.n_out ? (cbeg + p.N / 2) : p.n_out;   // this warp's valid columns
          // (the epilogue issues ~25 instructions per output element when every element drags its own shared-memory
          // bias load and bounds predicate along: whole 16-column groups take the straight-line path below)
          float m = -INFINITY;
          for (int c0 = cbeg; c0 < cend; c0 += 16) {
            float a[16], bb[16];
            tmem_ld16(tacc + c0, a);
            ld_bias16(c0, bb);
            tmem_wait_ld();
            if (c0 + 16 <= cend) {
#pragma unroll
              for (int i = 0; i < 16; ++i) m = fmaxf(m, fmaf(a[i], L2E, bb[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c0 + i < cend) m = fmaxf(m, fmaf(a[i], L2E, bb[i]));
            }
          }
          scratch[(h * RB_TILE + row) * 2] = m;
          epi_bar();
          mx = fmaxf(m, scratch[((h ^ 1) * RB_TILE + row) * 2]);
          float sum = 0.f;
          for (int c0 = cbeg; c0 < cend; c0 += 16) {
            float a[16], bb[16];
            tmem_ld16(tacc + c0, a);
            ld_bias16(c0, bb);
            tmem_wait_ld();
            if (c0 + 16 <= cend) {
#pragma unroll
              for (int i = 0; i < 16; ++i) sum += fast_ex2(fmaf(a[i], L2E, bb[i]) - mx);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c0 + i < cend) sum += fast_ex2(fmaf(a[i], L2E, bb[i]) - mx);
            }
          }
          scratch[(h * RB_TILE + row) * 2 + 1] = sum;
          epi_bar();
          inv = 1.f / (sum + scratch[((h ^ 1) * RB_TILE + row) * 2 + 1]);
          epi_bar();      // scratch may be rewritten by the next tile only after everyone has read it
        }
        const int esize = p.out_f32 ? 4 : 2;
        for (int c32 = 0; c32 * 32 < p.n_out; ++c32, ++nchunk) {
          const int col = c32 * 32 + h * 16;
          float a[16], bb[16];
          tmem_ld16(tacc + col, a);
          ld_bias16(col, bb);
          tmem_wait_ld();
          float v[16];                                           // columns >= n_out hold zero weights and bias
          if (p.softmax) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fast_ex2(fmaf(a[i], L2E, bb[i]) - mx) * inv;
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(a[i], sc, bb[i]);
          }
          if (p.tma_out) {
            const uint32_t boff = (nchunk & 1u) * DN_ABYTES;
            if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            epi_bar();
            uint8_t* sbase = smem_gen + (stg_base - smem_base) + boff;     // [32 channels][128 frames]
            if (p.out_f32) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                *reinterpret_cast<float*>(sbase + ((h * 16 + i) * RB_TILE + row) * 4) = v[i];
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                *reinterpret_cast<bf16*>(sbase + ((h * 16 + i) * RB_TILE + row) * 2) = __float2bfloat16_rn(v[i]);
            }
            fence_proxy_async_smem();
            epi_bar();
            if (issuer) {
              tma_store_3d(&map_y, stg_base + boff, t0, c32 * 32, b);
              bulk_commit();
            }
          } else {
            const int t = t0 + row;
            if (t < p.T) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int c = col + i;
                if (c < p.n_out) {
                  const long long o = ((long long)b * p.n_out + c) * p.T + t;
                  if (p.out_f32) reinterpret_cast<float*>(p.out)[o] = v[i];
                  else reinterpret_cast<bf16*>(p.out)[o] = __float2bfloat16_rn(v[i]);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(mapa_shared(acc_empty(r), 0));
    }
    if (p.colsum && p.mode == 0)
    {
      if (p.gb_gate) {                           // the CTA's shared accumulators, one global add per column
        epi_bar();
        const int t256 = threadIdx.x - 64;
        if (t256 < p.N) {
          float a1 = 0.f, a2 = 0.f;
#pragma unroll
          for (int wq = 0; wq < 8; ++wq) { a1 += csacc[wq * 512 + t256]; a2 += csacc[wq * 512 + 256 + t256]; }
          atomicAdd(p.colsum + t256, a1);
          atomicAdd(p.colsum + p.N + t256, a2);
        }
      } else {
        for (int q = 0; q < p.N / 64; ++q) atomicAdd(p.colsum + q * 64 + cs_col, csum[q]);
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

// NCL -> NLC bf16: 64 x 64 (channels x frames) tiles through shared memory, 16/32-byte accesses on both sides
template <typename T>
__global__ void __launch_bounds__(256) ncl_to_nlc_v2_kernel(int C, int Tn, long long sb, long long sc, bool vec, bool f16,
                                                            const T* x, bf16* y) {
  // x[b, c, t] at x + b * sb + c * sc + t: a (B, C, T) tensor or a time slice of a longer one (train.py:30 feeds
  // sig[:, :, 0:-1]) is read in place; `vec`: every row starts 16-byte aligned
  __shared__ __align__(16) bf16 tile[64][72];     // [frame][channel], 144-byte rows
  const int b = blockIdx.z, c0 = blockIdx.y * 64, t0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
  {   // load: thread -> (channel = tid % 64, 16 consecutive frames)
    const int c = tid & 63, tq = (tid >> 6) * 16;
    const T* src = x + (long long)b * sb + (long long)(c0 + c) * sc + t0 + tq;
    bf16 v[16];
    const bool cok = (c0 + c) < C;
    if (cok && t0 + tq + 16 <= Tn && vec && sizeof(T) == 2) {
      *reinterpret_cast<uint4*>(&v[0]) = __ldg(reinterpret_cast<const uint4*>(src));
      *reinterpret_cast<uint4*>(&v[8]) = __ldg(reinterpret_cast<const uint4*>(src) + 1);
      if (f16) {       // bf16 -> fp16 bit patterns (exact for |v| in [2^-14, 65504]; the tile holds raw 16-bit words)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const __half hv = __float2half_rn(__bfloat162float(v[i]));
          v[i] = *reinterpret_cast<const bf16*>(&hv);
        }
      }
    } else if (f16) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const __half hv = __float2half_rn((cok && t0 + tq + i < Tn) ? to_f32<T>(src[i]) : 0.f);
        v[i] = *reinterpret_cast<const bf16*>(&hv);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        v[i] = (cok && t0 + tq + i < Tn) ? __float2bfloat16_rn(to_f32<T>(src[i])) : __float2bfloat16_rn(0.f);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) tile[tq + i][c] = v[i];
  }
  __syncthreads();
  {   // store: thread -> (frame = tid / 4, 16 consecutive channels)
    const int t = tid >> 2, cq = (tid & 3) * 16;
    if (t0 + t < Tn) {
      bf16* dst = y + ((long long)b * Tn + t0 + t) * C + c0 + cq;
      if (c0 + cq + 16 <= C && (C % 8 == 0)) {
        reinterpret_cast<uint4*>(dst)[0] = *reinterpret_cast<const uint4*>(&tile[t][cq]);
        reinterpret_cast<uint4*>(dst)[1] = *reinterpret_cast<const uint4*>(&tile[t][cq + 8]);
      } else {
        for (int i = 0; i < 16; ++i)
          if (c0 + cq + i < C) dst[i] = tile[t][cq + i];
      }
    }
  }
}


// NCL bf16 -> NLC (bf16 / fp16), rows 16-byte aligned, C % 64 == 0: a thread reads 8 frames (16 bytes) of TWO adjacent
// channels -- eight lanes cover 128 contiguous bytes of a row, the v2 kernel's lanes read 32 bytes from 32 different rows --
// interleaves them into eight (c, c+1) words, the words cross a [64 frames][33] shared tile (odd pitch: conflict-free
// writes by frame rows, 2-way reads), and leave as 16-byte vectors of 8 channels.  0.70 -> 0.9 of the copy peak.
__global__ void __launch_bounds__(256) ncl_to_nlc_v3_kernel(int C, int Tn, long long sb, long long sc, bool f16,
                                                            const bf16* __restrict__ x, bf16* __restrict__ y) {
  __shared__ uint32_t tile[64][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 64, t0 = blockIdx.x * 64;
  const int cp = threadIdx.x >> 3, fv = (threadIdx.x & 7) * 8;
  const bf16* r0 = x + (long long)b * sb + (long long)(c0 + 2 * cp) * sc + t0 + fv;
  const bf16* r1 = r0 + sc;
  uint32_t a[4] = {0, 0, 0, 0}, d[4] = {0, 0, 0, 0};
  if (t0 + fv + 8 <= Tn) {
    const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(r0)), q1 = __ldg(reinterpret_cast<const uint4*>(r1));
    a[0] = q0.x; a[1] = q0.y; a[2] = q0.z; a[3] = q0.w;
    d[0] = q1.x; d[1] = q1.y; d[2] = q1.z; d[3] = q1.w;
  } else {
    uint16_t* ae = reinterpret_cast<uint16_t*>(a);
    uint16_t* de = reinterpret_cast<uint16_t*>(d);
    for (int q = 0; q < 8; ++q)
      if (t0 + fv + q < Tn) {
        ae[q] = *reinterpret_cast<const uint16_t*>(r0 + q);
        de[q] = *reinterpret_cast<const uint16_t*>(r1 + q);
      }
  }
  if (f16) {           // bf16 -> fp16 bit patterns (the same conversion as the v2 kernel)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 ha = __floats2half2_rn(__uint_as_float(a[j] << 16), __uint_as_float(a[j] & 0xffff0000u));
      const __half2 hd = __floats2half2_rn(__uint_as_float(d[j] << 16), __uint_as_float(d[j] & 0xffff0000u));
      a[j] = *reinterpret_cast<const uint32_t*>(&ha);
      d[j] = *reinterpret_cast<const uint32_t*>(&hd);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    tile[fv + 2 * j][cp] = __byte_perm(a[j], d[j], 0x5410);          // (frame 2j:   channel c | channel c+1)
    tile[fv + 2 * j + 1][cp] = __byte_perm(a[j], d[j], 0x7632);      // (frame 2j+1: channel c | channel c+1)
  }
  __syncthreads();
#pragma unroll
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {
    const int f = i >> 3, v = i & 7;
    if (t0 + f < Tn) {
      const uint4 o = make_uint4(tile[f][4 * v], tile[f][4 * v + 1], tile[f][4 * v + 2], tile[f][4 * v + 3]);
      *reinterpret_cast<uint4*>(y + ((long long)b * Tn + t0 + f) * C + c0 + 8 * v) = o;
    }
  }
}

// RawCTCNet featuriser, first layer (raw_ctcnet.py:57-59): Conv1d(1, F, fk, padding=fk-1) + LeakyReLU on the raw
// 1-channel signal, written as NLC bf16 [B, T+fk-1, F].  Bandwidth kernel: 4 B read, 2F B written per frame.
// Block = 32 frames x F channels; thread = 8 consecutive channels of one frame -> 16-byte coalesced stores.
// A thread owns 8 consecutive channels for the whole block (its fk x 8 weights and 8 biases live in registers when
// fk <= 4), the F/8 threads of a frame sit side by side (a warp writes whole 512-byte rows at F = 256), and a block
// walks FEAT_FRAMES frames so that the per-block set-up (weights, signal window) is amortised.
constexpr int FEAT_FRAMES = 512;
template <typename T, int FKR>      // FKR = fk when the weights fit in registers (1..4), 0 = weights read from smem
__global__ void __launch_bounds__(256) featurize_nlc_kernel(int Tn, int F, int fk, bool f16, const T* x, const float* w,
                                                            const float* bias, bf16* y) {
  extern __shared__ float fsm[];            // [fk][F] weights (tap-major), [FEAT_FRAMES + fk] signal window
  float* ws = fsm;
  float* xs = fsm + fk * F;
  const int b = blockIdx.y, t0 = blockIdx.x * FEAT_FRAMES, To = Tn + fk - 1;
  if (FKR == 0)
    for (int i = threadIdx.x; i < fk * F; i += blockDim.x) {
      const int j = i / F, f = i - j * F;
      ws[i] = w[f * fk + j];
    }
  for (int i = threadIdx.x; i < FEAT_FRAMES + fk - 1; i += blockDim.x) {
    const int t = t0 + i - (fk - 1);
    xs[i] = (t >= 0 && t < Tn) ? to_f32<T>(x[(long long)b * Tn + t]) : 0.f;
  }
  const int groups = F / 8, rows = 256 / groups;          // rows of frames per pass (groups <= 256)
  const int gq = threadIdx.x % groups, rl = threadIdx.x / groups, f0 = gq * 8;
  float wr[(FKR > 0 ? FKR : 1) * 8], br[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) br[k] = bias[f0 + k];
  if (FKR > 0) {
#pragma unroll
    for (int j = 0; j < FKR; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) wr[j * 8 + k] = w[(f0 + k) * FKR + j];
  }
  __syncthreads();
  if (rl >= rows) return;
  for (int tl = rl; tl < FEAT_FRAMES; tl += rows) {
    const int t = t0 + tl;
    if (t >= To) break;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = br[k];
    if (FKR > 0) {
#pragma unroll
      for (int j = 0; j < FKR; ++j) {
        const float xv = xs[tl + j];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(wr[j * 8 + k], xv, acc[k]);
      }
    } else {
      for (int j = 0; j < fk; ++j) {
        const float xv = xs[tl + j];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(ws[j * F + f0 + k], xv, acc[k]);
      }
    }
    uint4 o;
    o.x = pack_act2(f16, leaky(acc[0]), leaky(acc[1]));
    o.y = pack_act2(f16, leaky(acc[2]), leaky(acc[3]));
    o.z = pack_act2(f16, leaky(acc[4]), leaky(acc[5]));
    o.w = pack_act2(f16, leaky(acc[6]), leaky(acc[7]));
    *reinterpret_cast<uint4*>(y + ((long long)b * To + t) * F + f0) = o;
  }
}

// AvgPool1d(pool) on an NCL tensor fused with the NCL -> NLC bf16 layout change (classifier.py:53,102):
// y[b, to, c] = mean_i x[b, c, to*pool + i].  32 x 32 (channels x pooled frames) tiles through shared memory.
template <typename T>
__global__ void __launch_bounds__(256) avgpool_ncl_to_nlc_kernel(int C, int Tn, int To, int pool, bool f16, const T* x, bf16* y) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  const float inv = 1.f / (float)pool;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, to = t0 + tx;
    float s = 0.f;
    if (c < C && to < To) {
      const T* src = x + ((long long)b * C + c) * Tn + (long long)to * pool;
      for (int k = 0; k < pool; ++k) s += to_f32<T>(src[k]);
    }
    tile[i][tx] = s * inv;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int to = t0 + i, c = c0 + tx;
    if (to < To && c < C) y[((long long)b * To + to) * C + c] = act16(f16, tile[tx][i]);
  }
}

// v2: rows are brought in with ALIGNED 16-byte loads whatever the row alignment (T - 1 = 16383 frames in the train
// step makes every row start on a different 2-byte phase): the vectors covering a row's segment land in shared memory
// at their natural position, the row's lead (0..15 bytes) is added when the pooled frames are read back.  A CTA owns
// 64 channels x TF pooled frames; a warp pools 8 channels with its lanes over frames (conflict-free shared reads),
// the bf16 results cross a padded [TF][66] tile and leave as 128-byte rows (one pooled frame x 64 channels).
template <typename T>
__global__ void __launch_bounds__(256)
avgpool_ncl_to_nlc_v2_kernel(int C, int Tn, int To, int pool, int TF, int RS, long long total_bytes, bool f16,
                             const T* x, bf16* y) {
  extern __shared__ __align__(16) uint8_t pool_smem[];
  constexpr int ES = (int)sizeof(T);
  uint8_t* in_s = pool_smem;                                   // [64][RS]
  bf16* out_s = reinterpret_cast<bf16*>(pool_smem + 64 * RS);  // [TF][66]
  __shared__ int lead_s[64];
  const int b = blockIdx.z, c0 = blockIdx.y * 64, t0 = blockIdx.x * TF;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = (TF * pool * ES + 15) / 16 + 1;
  const uint8_t* xb = reinterpret_cast<const uint8_t*>(x);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = warp + 8 * i, c = c0 + row;
    const long long start = (((long long)b * C + c) * Tn + (long long)t0 * pool) * ES;
    const long long a0 = start & ~15LL;
    if (lane == 0) lead_s[row] = (int)(start - a0);
    for (int j = lane; j < nvec; j += 32) {
      const long long a = a0 + 16LL * j;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (c < C) {
        if (a + 16 <= total_bytes) {
          v = __ldg(reinterpret_cast<const uint4*>(xb + a));
        } else {                                               // the tensor's last bytes: element by element
          T* e = reinterpret_cast<T*>(&v);
          for (int q = 0; q < 16 / ES; ++q)
            if (a + (q + 1) * ES <= total_bytes) e[q] = *reinterpret_cast<const T*>(xb + a + q * ES);
        }
      }
      *reinterpret_cast<uint4*>(in_s + row * RS + 16 * j) = v;
    }
  }
  __syncthreads();
  const float inv = 1.f / (float)pool;
  for (int f = lane; f < TF; f += 32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int row = warp * 8 + j;
      const T* src = reinterpret_cast<const T*>(in_s + row * RS + lead_s[row]) + f * pool;
      float acc = 0.f;
      for (int k = 0; k < pool; ++k) acc += to_f32<T>(src[k]);
      out_s[f * 66 + row] = act16(f16, acc * inv);
    }
  }
  __syncthreads();
  const uint32_t* ow = reinterpret_cast<const uint32_t*>(out_s);
  for (int idx = threadIdx.x; idx < TF * 32; idx += 256) {
    const int f = idx >> 5, wc = idx & 31;
    if (t0 + f < To && c0 + 2 * wc + 1 < C)
      *reinterpret_cast<uint32_t*>(y + ((long long)b * To + t0 + f) * C + c0 + 2 * wc) = ow[f * 33 + wc];
  }
}

// Entry conv of a WaveNet fed with quantised LEVELS instead of their one-hot encoding (wavenet.py:93 applied to
// fns.py:6-15 / pore_model.py:88-96): a convolution over a one-hot signal is a gather,
//     y[b, t, :] = bias + sum_j Wemb[j][lev[b, t + off_j]][:]          (taps outside [0, T) contribute nothing),
// with Wemb[j][l][c] = W[c, l, j] in bf16.  One warp per frame, a lane owns 8 channels per 256 (16-byte loads from the
// L1-resident table, one 16-byte store).  Same products, same fp32 accumulation order as the dense kernel on the one-hot
// tensor, without the 2 x 256-fold larger input.
__global__ void __launch_bounds__(256)
entry_embed_nlc_kernel(long long frames, int Tn, int C, int in_dim, int ntaps, int o0, int o1, int o2,
                       const int* __restrict__ lev, const bf16* __restrict__ wemb, const float* __restrict__ bias,
                       bf16* __restrict__ y, bf16* __restrict__ y_lo, bool f16) {
  const int lane = threadIdx.x & 31;
  const long long f = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (f >= frames) return;
  const int t = (int)(f % Tn);
  const int offs[3] = {o0, o1, o2};
  int l[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    l[j] = -1;
    if (j < ntaps) {
      const int tt = t + offs[j];
      if (tt >= 0 && tt < Tn) {
        const int v = lev[f + offs[j]];
        l[j] = v < 0 ? 0 : (v >= in_dim ? in_dim - 1 : v);
      }
    }
  }
  for (int c0 = lane * 8; c0 < C; c0 += 256) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (l[j] < 0) continue;
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(wemb + ((long long)j * in_dim + l[j]) * C + c0));
      const __nv_bfloat162* w2 = reinterpret_cast<const __nv_bfloat162*>(&w);
      const __half2* h2 = reinterpret_cast<const __half2*>(&w);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 v = f16 ? __half22float2(h2[q]) : __bfloat1622float2(w2[q]);
        acc[2 * q] += v.x;
        acc[2 * q + 1] += v.y;
      }
    }
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
    uint4 o;
    if (y_lo) {      // fp16 (hi, lo) pair
      uint4 l;
      split_f16x2(acc[0] + b0.x, acc[1] + b0.y, o.x, l.x);
      split_f16x2(acc[2] + b0.z, acc[3] + b0.w, o.y, l.y);
      split_f16x2(acc[4] + b1.x, acc[5] + b1.y, o.z, l.z);
      split_f16x2(acc[6] + b1.z, acc[7] + b1.w, o.w, l.w);
      *reinterpret_cast<uint4*>(y_lo + f * C + c0) = l;
    } else {
      o.x = pack_act2(f16, acc[0] + b0.x, acc[1] + b0.y);
      o.y = pack_act2(f16, acc[2] + b0.z, acc[3] + b0.w);
      o.z = pack_act2(f16, acc[4] + b1.x, acc[5] + b1.y);
      o.w = pack_act2(f16, acc[6] + b1.z, acc[7] + b1.w);
    }
    *reinterpret_cast<uint4*>(y + f * C + c0) = o;
  }
}

// Backward of the mean-pool + layout change above, fused the other way round: the classifier's input gradient
// dh [B, To, C] (NLC bf16) -> dx [B, C, T] (NCL): dx[b, c, t] = dh[b, t / pool, c] / pool for t < To * pool, 0 after
// (autograd of nn.AvgPool1d, classifier.py:102).  [64 pooled frames x 64 channels] tiles through shared memory; a
// warp owns 8 channels, its lanes walk pooled frames and write the `pool` copies -- no division in the loop.
template <typename TO>
__global__ void __launch_bounds__(256)
avgpool_bwd_nlc_to_ncl_kernel(int C, int Tn, int To, int pool, const bf16* __restrict__ dh, TO* __restrict__ dx) {
  __shared__ __align__(16) bf16 tile[64][72];      // [pooled frame][channel]
  const int b = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {
    const int pf = i >> 3, v = i & 7;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (p0 + pf < To && c0 + v * 8 + 8 <= C)
      val = __ldg(reinterpret_cast<const uint4*>(dh + ((long long)b * To + p0 + pf) * C + c0 + v * 8));
    *reinterpret_cast<uint4*>(&tile[pf][v * 8]) = val;
  }
  __syncthreads();
  const float fpool = (float)pool;                 // a true division, like autograd's grad / kernel_size
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + warp * 8 + j;
    if (c >= C) break;
    TO* row = dx + ((long long)b * C + c) * Tn;
    for (int pf = lane; pf < 64; pf += 32) {
      if (p0 + pf >= To) break;
      const TO val = from_f32<TO>(__fdiv_rn(__bfloat162float(tile[pf][warp * 8 + j]), fpool));
      TO* dst = row + (long long)(p0 + pf) * pool;
      for (int k = 0; k < pool; ++k) dst[k] = val;
    }
    if (p0 + 64 >= To)                             // this tile ends the pooled range: zero the tail [To * pool, T)
      for (int t = To * pool + lane; t < Tn; t += 32) row[t] = from_f32<TO>(0.f);
  }
}

}  // namespace wnb

using namespace wnb;

extern "C" int wnb200_avgpool_bwd_nlc_to_ncl(int dtype_out, int B, int C, int T_, int pool, const void* dh, void* dx,
                                             void* stream) {
  WNB_CHECK_ARG(dtype_out == WNB200_F32 || dtype_out == WNB200_BF16, "avgpool_bwd_nlc_to_ncl: bad dtype");
  WNB_CHECK_ARG(pool >= 1 && C % 8 == 0, "avgpool_bwd_nlc_to_ncl: pool >= 1 and C %% 8 == 0 required");
  if (B == 0 || C == 0 || T_ == 0) return 0;
  WNB_CHECK_ARG(dx != nullptr, "avgpool_bwd_nlc_to_ncl: null pointer");
  const int To = T_ / pool;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = dtype_out == WNB200_F32 ? 4 : 2;
  if (To == 0) {
    WNB_CUDA_OK(cudaMemsetAsync(dx, 0, (size_t)B * C * T_ * es, st));
    return 0;
  }
  WNB_CHECK_ARG(dh != nullptr, "avgpool_bwd_nlc_to_ncl: null pointer");
  WNB_CHECK_ARG(B <= 65535 && ceil_div(C, 64) <= 65535, "avgpool_bwd_nlc_to_ncl: shape too large");
  dim3 grid(ceil_div(To, 64), ceil_div(C, 64), B);
  if (dtype_out == WNB200_F32)
    avgpool_bwd_nlc_to_ncl_kernel<float><<<grid, 256, 0, st>>>(C, T_, To, pool, (const bf16*)dh, (float*)dx);
  else
    avgpool_bwd_nlc_to_ncl_kernel<bf16><<<grid, 256, 0, st>>>(C, T_, To, pool, (const bf16*)dh, (bf16*)dx);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_entry_embed_nlc(int B, int T_, int C, int in_dim, int ntaps, const int32_t* t_off,
                                      const int32_t* levels, const void* wemb, const float* bias, int act_fmt, void* y,
                                      void* y_lo, void* stream) {
  WNB_CHECK_ARG(act_fmt == WNB200_ACT_BF16 || act_fmt == WNB200_ACT_F16X2, "entry_embed_nlc: bad act_fmt");
  WNB_CHECK_ARG(!y_lo || act_fmt == WNB200_ACT_F16X2, "entry_embed_nlc: y_lo needs the fp16 (hi, lo) format");
  WNB_CHECK_ARG(C >= 8 && C % 8 == 0 && in_dim >= 1 && ntaps >= 1 && ntaps <= 3, "entry_embed_nlc: bad sizes");
  if (B == 0 || T_ == 0) return 0;
  WNB_CHECK_ARG(t_off && levels && wemb && bias && y, "entry_embed_nlc: null pointer");
  const long long frames = (long long)B * T_;
  WNB_CHECK_ARG(frames / 8 + 1 < (1ll << 31), "entry_embed_nlc: too many frames");
  entry_embed_nlc_kernel<<<(unsigned)ceil_div64(frames, 8), 256, 0, (cudaStream_t)stream>>>(
      frames, T_, C, in_dim, ntaps, t_off[0], ntaps > 1 ? t_off[1] : 0, ntaps > 2 ? t_off[2] : 0, levels,
      (const bf16*)wemb, bias, (bf16*)y, (bf16*)y_lo, act_fmt == WNB200_ACT_F16X2);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_dense_fwd_tc(const wnb200_dense_t* a, void* stream) {
  WNB_CHECK_ARG(a != nullptr, "dense_fwd_tc: null argument");
  WNB_CHECK_STRUCT(a, wnb200_dense_t, "dense_fwd_tc");
  if (a->B == 0 || a->T == 0) return 0;
  WNB_CHECK_ARG(a->Cin >= 64 && a->Cin % 64 == 0, "dense_fwd_tc: Cin=%d must be a multiple of 64", a->Cin);
  WNB_CHECK_ARG(a->ntaps >= 1 && a->ntaps <= 3, "dense_fwd_tc: ntaps=%d not in 1..3", a->ntaps);
  WNB_CHECK_ARG(a->N >= 16 && a->N <= 256 && a->N % 16 == 0, "dense_fwd_tc: N=%d must be a multiple of 16 <= 256", a->N);
  WNB_CHECK_ARG(a->x && a->w && a->bias && a->y, "dense_fwd_tc: null pointer");
  WNB_CHECK_ARG(a->mode == 0 || a->mode == 1, "dense_fwd_tc: bad mode");
  WNB_CHECK_ARG(a->mode != 0 || a->N % 64 == 0, "dense_fwd_tc: NLC output needs N %% 64 == 0");
  WNB_CHECK_ARG(a->mode != 1 || (a->n_out >= 1 && a->n_out <= a->N), "dense_fwd_tc: bad n_out");
  DenseDev p;
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.T = a->T;
  p.tiles_per_seq = ceil_div(a->T, 2 * RB_TILE);
  p.num_tiles = p.tiles_per_seq * a->B;
  p.ntaps = a->ntaps;
  for (int j = 0; j < 3; ++j) p.t_off[j] = a->t_off[j];
  p.kb_per_tap = a->Cin / 64;
  p.ntaps2 = a->x2 ? a->ntaps2 : 0;
  for (int j = 0; j < 3; ++j) p.t_off2[j] = a->t_off2[j];
  p.kb_per_tap2 = a->Cin2 / 64;
  if (p.ntaps2 > 0)
    WNB_CHECK_ARG(a->Cin2 >= 64 && a->Cin2 % 64 == 0 && a->ntaps2 <= 3, "dense_fwd_tc: bad second source");
  p.N = a->N; p.mode = a->mode; p.leaky = a->leaky;
  p.n_out = a->n_out; p.softmax = a->softmax; p.out_f32 = a->out_f32;
  p.bias = a->bias; p.out = a->y;
  p.colsum = a->mode == 0 ? a->colsum : nullptr;
  WNB_CHECK_ARG(a->act_fmt == WNB200_ACT_BF16 || a->act_fmt == WNB200_ACT_F16X2, "dense_fwd_tc: bad act_fmt %d", a->act_fmt);
  p.f16 = a->act_fmt == WNB200_ACT_F16X2;
  p.split = p.f16 && a->mode == 0 && a->y_lo != nullptr;
  p.nlayers = a->nlayers;
  WNB_CHECK_ARG(a->nlayers >= 0 && (a->nlayers == 0 || (a->ntaps == 1 && a->t_off[0] == 0 && !a->x2)),
                "dense_fwd_tc: a layer stack is contracted with one zero-offset tap and no second source");
  WNB_CHECK_ARG(!a->y_lo || p.split, "dense_fwd_tc: y_lo needs act_fmt = WNB200_ACT_F16X2 and the NLC mode");
  WNB_CHECK_ARG(!p.f16 || !p.colsum, "dense_fwd_tc: colsum is a bf16 (training) feature");
  WNB_CHECK_ARG((a->gb_gate == nullptr) == (a->gb_sg == nullptr), "dense_fwd_tc: gb_gate and gb_sg come together");
  WNB_CHECK_ARG(!a->gb_gate || (a->mode == 0 && !p.f16 && !a->leaky),
                "dense_fwd_tc: the gate-backward epilogue is an NLC, bf16, linear one");
  WNB_CHECK_ARG(!a->gb_gate || ((reinterpret_cast<uintptr_t>(a->gb_gate) | reinterpret_cast<uintptr_t>(a->gb_sg)) & 15) == 0,
                "dense_fwd_tc: gb_gate / gb_sg must be 16-byte aligned");
  WNB_CHECK_ARG((a->pos_w == nullptr) == (a->pos_b == nullptr), "dense_fwd_tc: pos_w and pos_b come together");
  WNB_CHECK_ARG(!a->pos_w || (a->mode == 0 && !a->gb_gate), "dense_fwd_tc: the position term belongs to the NLC mode");
  p.pos_w = a->pos_w; p.pos_b = a->pos_b; p.pos_t0 = a->pos_t0;
  p.gb_gate = reinterpret_cast<const bf16*>(a->gb_gate);
  p.gb_sg = reinterpret_cast<const bf16*>(a->gb_sg);
  const int esize = a->out_f32 ? 4 : 2;
  p.tma_out = (a->mode == 1 && ((long long)a->T * esize) % 16 == 0) ? 1 : 0;
  CUtensorMap mx, mx2, mw, my, mylo;
  int rc;
  if ((rc = rb_map_nlc(&mx, a->x, a->B * (a->nlayers > 0 ? a->nlayers : 1), a->T, a->Cin, 2))) return rc;
  mx2 = mx;
  if (p.ntaps2 > 0 && (rc = rb_map_nlc(&mx2, a->x2, a->B, a->T, a->Cin2, 2))) return rc;
  if ((rc = rb_map_2d(&mw, a->w, a->N, (a->nlayers > 0 ? a->nlayers : a->ntaps) * a->Cin + p.ntaps2 * a->Cin2,
                      a->N / 2))) return rc;
  if (a->mode == 0) {
    if ((rc = rb_map_nlc(&my, a->y, a->B, a->T, a->gb_gate ? 2 * a->N : a->N, 2))) return rc;
  } else if (p.tma_out) {
    if ((rc = rb_map_ncl(&my, a->y, a->B, a->n_out, a->T, esize, RB_TILE, a->n_out < 32 ? a->n_out : 32))) return rc;
  } else {
    my = mx;
  }
  mylo = my;
  if (p.split && (rc = rb_map_nlc(&mylo, a->y_lo, a->B, a->T, a->N, 2))) return rc;
  WNB_SET_SMEM_ATTR(DN_SMEM, dense2_kernel);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int pairs = sms / 2;
  if (p.num_tiles < pairs) pairs = p.num_tiles;
  dense2_kernel<<<2 * pairs, DN_THREADS, DN_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(mx, mx2, mw, my, mylo, p);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_ncl_to_nlc_act(int dtype, int act_fmt, int B, int C, int T_, int64_t sb, int64_t sc, const void* x,
                                     void* y, void* stream) {
  WNB_CHECK_ARG(dtype == WNB200_F32 || dtype == WNB200_BF16, "ncl_to_nlc_bf16: bad dtype");
  WNB_CHECK_ARG(act_fmt == WNB200_ACT_BF16 || act_fmt == WNB200_ACT_F16X2, "ncl_to_nlc_act: bad act_fmt");
  const bool f16 = act_fmt == WNB200_ACT_F16X2;
  if (B == 0 || C == 0 || T_ == 0) return 0;
  WNB_CHECK_ARG(x && y, "ncl_to_nlc_bf16: null pointer");
  WNB_CHECK_ARG(B <= 65535 && ceil_div(C, 64) <= 65535, "ncl_to_nlc_bf16: shape too large");
  WNB_CHECK_ARG(sc >= T_ && sb >= 0, "ncl_to_nlc_bf16: bad strides");
  dim3 grid(ceil_div(T_, 64), ceil_div(C, 64), B);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = dtype == WNB200_BF16 && sb % 8 == 0 && sc % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (dtype == WNB200_F32)
    ncl_to_nlc_v2_kernel<float><<<grid, 256, 0, st>>>(C, T_, sb, sc, vec, f16, (const float*)x, (bf16*)y);
  else if (vec && C % 64 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0)
    ncl_to_nlc_v3_kernel<<<grid, 256, 0, st>>>(C, T_, sb, sc, f16, (const bf16*)x, (bf16*)y);
  else
    ncl_to_nlc_v2_kernel<bf16><<<grid, 256, 0, st>>>(C, T_, sb, sc, vec, f16, (const bf16*)x, (bf16*)y);
  WNB_LAUNCH_OK();
  return 0;
}

extern "C" int wnb200_ncl_to_nlc_bf16_strided(int dtype, int B, int C, int T_, int64_t sb, int64_t sc, const void* x,
                                              void* y, void* stream) {
  return wnb200_ncl_to_nlc_act(dtype, WNB200_ACT_BF16, B, C, T_, sb, sc, x, y, stream);
}

extern "C" int wnb200_ncl_to_nlc_bf16(int dtype, int B, int C, int T_, const void* x, void* y, void* stream) {
  return wnb200_ncl_to_nlc_bf16_strided(dtype, B, C, T_, (int64_t)C * T_, (int64_t)T_, x, y, stream);
}

extern "C" int wnb200_featurize_nlc(int dtype, int B, int T_, int F, int fk, const void* x, const float* w,
                                    const float* bias, int act_fmt, void* y, void* stream) {
  WNB_CHECK_ARG(dtype == WNB200_F32 || dtype == WNB200_BF16, "featurize_nlc: bad dtype");
  WNB_CHECK_ARG(act_fmt == WNB200_ACT_BF16 || act_fmt == WNB200_ACT_F16X2, "featurize_nlc: bad act_fmt");
  const bool f16 = act_fmt == WNB200_ACT_F16X2;
  WNB_CHECK_ARG(F >= 8 && F % 8 == 0 && fk >= 1 && fk <= 64, "featurize_nlc: F=%d fk=%d unsupported", F, fk);
  if (B == 0 || T_ == 0) return 0;
  WNB_CHECK_ARG(x && w && bias && y, "featurize_nlc: null pointer");
  WNB_CHECK_ARG(B <= 65535, "featurize_nlc: batch too large");
  const int To = T_ + fk - 1;
  WNB_CHECK_ARG(F <= 2048, "featurize_nlc: F=%d too large", F);
  const size_t smem = sizeof(float) * ((size_t)fk * F + FEAT_FRAMES + fk);
  WNB_CHECK_ARG(smem <= 48 * 1024, "featurize_nlc: fk*F too large for the weight cache");
  dim3 grid(ceil_div(To, FEAT_FRAMES), B);
  cudaStream_t st = (cudaStream_t)stream;
#define FEAT_LAUNCH(TT, FKR) featurize_nlc_kernel<TT, FKR><<<grid, 256, smem, st>>>(T_, F, fk, f16, (const TT*)x, w, bias, (bf16*)y)
#define FEAT_DISPATCH(TT)                         \
  switch (fk) {                                   \
    case 1: FEAT_LAUNCH(TT, 1); break;            \
    case 2: FEAT_LAUNCH(TT, 2); break;            \
    case 3: FEAT_LAUNCH(TT, 3); break;            \
    case 4: FEAT_LAUNCH(TT, 4); break;            \
    default: FEAT_LAUNCH(TT, 0); break;           \
  }
  if (dtype == WNB200_F32) { FEAT_DISPATCH(float) } else { FEAT_DISPATCH(bf16) }
#undef FEAT_DISPATCH
#undef FEAT_LAUNCH
  WNB_LAUNCH_OK();
  return 0;
}

namespace wnb { int avgpool_ncl_to_nlc_v3_launch(int B, int C, int T_, int pool, bool f16, const void* x, void* y, cudaStream_t st); }

extern "C" int wnb200_avgpool_ncl_to_nlc(int dtype, int B, int C, int T_, int pool, const void* x, int act_fmt, void* y,
                                         void* stream) {
  WNB_CHECK_ARG(dtype == WNB200_F32 || dtype == WNB200_BF16, "avgpool_ncl_to_nlc: bad dtype");
  WNB_CHECK_ARG(act_fmt == WNB200_ACT_BF16 || act_fmt == WNB200_ACT_F16X2, "avgpool_ncl_to_nlc: bad act_fmt");
  const bool f16 = act_fmt == WNB200_ACT_F16X2;
  WNB_CHECK_ARG(pool >= 1, "avgpool_ncl_to_nlc: bad pool");
  const int To = T_ / pool;
  if (B == 0 || C == 0 || To == 0) return 0;
  WNB_CHECK_ARG(x && y, "avgpool_ncl_to_nlc: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int es = dtype == WNB200_F32 ? 4 : 2;
  // v3 (bytenet.cu): bf16 rows that start 16-byte aligned, whole 64-channel groups, pool <= 4
  if (dtype == WNB200_BF16 && pool <= 4 && C % 64 == 0 && T_ % 8 == 0 && B <= 65535 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    if (avgpool_ncl_to_nlc_v3_launch(B, C, T_, pool, f16, x, y, st) == 0) {
      WNB_LAUNCH_OK();
      return 0;
    }
  }
  // v2 (aligned vector loads, any row alignment): C even, base pointers 16-byte aligned, a [64 x TF*pool] tile in 40 KB
  int TF = 64;
  while (TF > 8 && TF * pool * es > 576) TF >>= 1;
  if (C % 2 == 0 && TF * pool * es <= 576 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(y) & 3) == 0) {
    const int RS = ((TF * pool * es + 15) / 16 + 1) * 16 + 16;
    const size_t smem = (size_t)64 * RS + (size_t)TF * 66 * 2;
    const long long total = (long long)B * C * T_ * es;
    dim3 grid2(ceil_div(To, TF), ceil_div(C, 64), B);
    if (dtype == WNB200_F32)
      avgpool_ncl_to_nlc_v2_kernel<float><<<grid2, 256, smem, st>>>(C, T_, To, pool, TF, RS, total, f16,
                                                                    (const float*)x, (bf16*)y);
    else
      avgpool_ncl_to_nlc_v2_kernel<bf16><<<grid2, 256, smem, st>>>(C, T_, To, pool, TF, RS, total, f16,
                                                                   (const bf16*)x, (bf16*)y);
    WNB_LAUNCH_OK();
    return 0;
  }
  dim3 grid(ceil_div(To, 32), ceil_div(C, 32), B);
  if (dtype == WNB200_F32)
    avgpool_ncl_to_nlc_kernel<float><<<grid, 256, 0, st>>>(C, T_, To, pool, f16, (const float*)x, (bf16*)y);
  else
    avgpool_ncl_to_nlc_kernel<bf16><<<grid, 256, 0, st>>>(C, T_, To, pool, f16, (const bf16*)x, (bf16*)y);
  WNB_LAUNCH_OK();
  return 0;
}
