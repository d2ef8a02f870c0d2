// All-pairs correlation volume + 4-level average-pooled pyramid for sm_100a in one kernel
// (GMA.whl!/GMA/core/corr.py:16-30 CorrBlock.__init__ and :55-63 CorrBlock.corr).
//
// The kernel is bound by the fp32 pyramid WRITE (4 N^2 * 1.33 bytes per pair; the MMA work is 13% of that in
// time), so it is organised around the store path:
//   * one CTA owns 128 query pixels (fmap1 rows, resident in shared memory) and streams the 8 x 32 target-pixel
//     tiles of fmap2 through a TMA ring; the 512 TMEM columns hold two 128 x 256 accumulators so the MMA warp
//     fills one while the epilogue drains the other;
//   * epilogue warps (two groups of four, one group per accumulator) own one query per thread, scale the
//     accumulator row, stage [32 queries x 32 targets] fp32 boxes in swizzled shared memory and hand them to
//     TMA stores: every box row is a contiguous 128-byte run of one query's level-0 map, written by the TMA
//     engine as full sectors (the previous per-thread float4 stores used 16 of every 32-byte sector and ran
//     at 24% of the HBM roofline);
//   * levels 1..3 are the hierarchical 2x2 means of corr.py:28-30 (floor semantics), accumulated in registers
//     from the same accumulator rows and stored the same way (boxes of 16 / 8 / 4 floats per query);
//   * TMA clipping against the true extents (W_l, H_l, N, batch) replaces every edge predicate.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_fp16.h>

#include "common.h"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace atdn {

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

constexpr int kCorrThreads = 384;
constexpr int kCorrBRingBytes = 3 * 256 * 128;     // ring of 64-channel chunks of the target tile: 3 x 32 KiB, or 6 x 16 KiB per CTA of a pair
constexpr int kCorrABytes = 4 * 128 * 128;         // 128 queries x 256 channels
constexpr int kCorrSlotBytes = 32 * 128;           // one staging slot: 32 queries x 32 fp32
constexpr int kCorrSmem = kCorrABytes + kCorrBRingBytes + 8 * 2 * kCorrSlotBytes + 1024;
constexpr int kCorrMaxBStages = 6;

struct alignas(64) CorrParams {
  CUtensorMap tmA, tmB, tmL[4];
  int h, w, tiles_w, tiles;
  float alpha;
  int dbg;                // experiment switches (ATDN_CORR_DBG): 1 = no level 1..3 stores, 2 = no stores, 4 = LSU level-0 stores (fp32)
  float* l0;
  int n, pitch0;
  __half* l3;             // HALF: level 3 is stored by the threads directly
  int pitch3;
  int tiles_w3, tiles3;   // HALF: level-3 tile columns rounded up to even (16-byte rows for the lookup's cp.async), tiles per query
};

// PAIR: a cluster of two CTAs (the two SMs of a TPC) owns 256 queries and executes cta_group::2 MMAs of M = 256: each CTA
// keeps its own 128 query rows resident and stages only HALF of every target tile (4 of its 8 rows), so the L2 -> SM
// operand stream per SM halves and the same ring bytes look ahead twice as far.  ncu on the single-CTA kernel (27 pairs,
// profiles/r02a_ncu_corr_pyramid_lookup_*): the MMA warp spent 60% of its time waiting for target chunks and the producer
// 74% waiting for free stages -- a latency-bound 3-stage ring (96 KiB in flight per SM against ~160 KiB needed once the
// pyramid stores load the L2), tensor pipe 37% busy, with the epilogue warps idle 64% of the time.
template <bool HALF, bool PAIR>
__global__ void __launch_bounds__(kCorrThreads, 1) corr_pyramid_kernel(const __grid_constant__ CorrParams p) {
  constexpr int CL = PAIR ? 2 : 1;
  constexpr int kCorrBStageBytes = (256 / CL) * 128;
  constexpr int kCorrBStages = kCorrBRingBytes / kCorrBStageBytes;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full, b_full[kCorrMaxBStages], b_empty[kCorrMaxBStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kCorrABytes;
  uint8_t* smem_st = smem_b + kCorrBStages * kCorrBStageBytes;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const int m0 = blockIdx.x * 128;                   // cluster c = CTAs 2c, 2c + 1: rank r owns queries (2c + r) * 128 ..
  const int batch = blockIdx.y;
  const int T = p.tiles;

  if (threadIdx.x == 0) {
    mbar_init(&a_full, 1);
    for (int s = 0; s < kCorrBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4 * CL); }
    fence_barrier_init();
  }
  if (warp == 2 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int l = 0; l < (HALF ? 3 : 4); ++l) tma_prefetch_desc(&p.tmL[l]);
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair(&tmem_base_smem, 512);
    else tmem_alloc(&tmem_base_smem, 512);
  }
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();            // the peer's barriers exist before any remote arrive / complete_tx
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===== producer (both CTAs of a pair: own query rows, own half of every target tile; bytes land on the leader's barriers) =====
    if (elect_one_sync()) {
      if (rank == 0) mbar_arrive_expect_tx(&a_full, CL * kCorrABytes);
      for (int c = 0; c < 4; ++c) {
        if constexpr (PAIR) tma_load_4d_pair(smem_a + c * 128 * 128, &p.tmA, &a_full, c * 64, m0, 0, batch);
        else tma_load_4d(smem_a + c * 128 * 128, &p.tmA, &a_full, c * 64, m0, 0, batch);
      }
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < T; ++t) {
      const int bh0 = (t / p.tiles_w) * 8 + rank * 4, bw0 = (t % p.tiles_w) * 32;
      for (int c = 0; c < 4; ++c) {
        mbar_wait(&b_empty[stage], phase ^ 1u);
        if (elect_one_sync()) {
          if (rank == 0) mbar_arrive_expect_tx(&b_full[stage], CL * kCorrBStageBytes);
          if constexpr (HALF) {
            // strip order: B row (= accumulator column) strip * 64 + row * 8 + col; boxes of 8 cols x 8 rows (8 KiB); a CTA
            // of a pair stages strips 2 * rank, 2 * rank + 1
            constexpr int kStrips = 4 / CL;
#pragma unroll
            for (int sidx = 0; sidx < kStrips; ++sidx) {
              uint8_t* dst = smem_b + stage * kCorrBStageBytes + sidx * 8192;
              const int x0 = (t % p.tiles_w) * 32 + (rank * kStrips + sidx) * 8, y0 = (t / p.tiles_w) * 8;
              if constexpr (PAIR) tma_load_4d_pair(dst, &p.tmB, &b_full[stage], c * 64, x0, y0, batch);
              else tma_load_4d(dst, &p.tmB, &b_full[stage], c * 64, x0, y0, batch);
            }
          } else {
            if constexpr (PAIR) tma_load_4d_pair(smem_b + stage * kCorrBStageBytes, &p.tmB, &b_full[stage], c * 64, bw0, bh0, batch);
            else tma_load_4d(smem_b + stage * kCorrBStageBytes, &p.tmB, &b_full[stage], c * 64, bw0, bh0, batch);
          }
        }
        __syncwarp();
        if (++stage == kCorrBStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader CTA of a pair) =====
    constexpr uint32_t kIdesc = make_idesc_f16(128 * CL, 256);
    const uint32_t a_u32 = smem_u32(smem_a), b_u32 = smem_u32(smem_b);
    int stage = 0;
    uint32_t phase = 0, pe0 = 0, pe1 = 0;
    mbar_wait(&a_full, 0);
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      const uint32_t pe = buf ? pe1 : pe0;
      mbar_wait(&acc_empty[buf], pe ^ 1u);
      if (buf) pe1 ^= 1u; else pe0 ^= 1u;
      tcgen05_fence_after();
      const uint32_t d = tmem_base + static_cast<uint32_t>(buf * 256);
      for (int c = 0; c < 4; ++c) {
        mbar_wait(&b_full[stage], phase);
        tcgen05_fence_after();
        const uint64_t a_desc = make_smem_desc_sw128(a_u32 + c * 128 * 128);
        const uint64_t b_desc = make_smem_desc_sw128(b_u32 + stage * kCorrBStageBytes);
        if (elect_one_sync()) {
          if constexpr (PAIR) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_pair(d, a_desc + 2u * k, b_desc + 2u * k, kIdesc, (c | k) ? 1u : 0u);
            umma_commit_pair(&b_empty[stage]);
            if (c == 3) umma_commit_pair(&acc_full[buf]);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d, a_desc + 2u * k, b_desc + 2u * k, kIdesc, (c | k) ? 1u : 0u);
            umma_commit(&b_empty[stage]);
            if (c == 3) umma_commit(&acc_full[buf]);
          }
        }
        __syncwarp();
        if (++stage == kCorrBStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    // group g (4 warps) drains every second tile (accumulator buffer g).  (Letting both groups drain every tile, two strips
    // each, was measured: no gain on level 0 and slower pooled-level stores -- the kernel is bound by the store path, not by
    // the epilogue arithmetic: profiles/r02k_*.)
    const int q = warp & 3, g = (warp - 4) >> 2;
    const uint32_t trow_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* st_ptr = smem_st + (warp - 4) * 2 * kCorrSlotBytes;
    const uint32_t st_u32 = smem_u32(st_ptr);
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    const int H1 = p.h / 2, H2 = H1 / 2, H3 = H2 / 2;
    const int qrow = m0 + q * 32;
    uint32_t pf = 0;
    int nstore = 0;

    // staging slot ring: a slot is reused once the bulk store issued two stores ago has read it
    auto acquire = [&]() -> int {
      // (a ring of four 2-KiB slots was measured: no faster -- slot reuse is not what the epilogue waits for)
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
      return (nstore & 1) * kCorrSlotBytes;
    };
    auto commit = [&](const CUtensorMap* m, int off, int c0, int c1) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(m, st_ptr + off, c0, c1, qrow, batch);
        bulk_commit();
      }
      ++nstore;
    };

    for (int t = g; t < T; t += 2) {
      const int bh0 = (t / p.tiles_w) * 8, bw0 = (t % p.tiles_w) * 32;
      const int buf = g;
      const uint32_t trow = trow_base + static_cast<uint32_t>(buf * 256);
      mbar_wait(&acc_full[buf], pf);
      pf ^= 1u;
      tcgen05_fence_after();
      if constexpr (HALF) {
        // fp16 pyramid, STRIP layout (include/atdn_b200.h): the target tile is loaded as four 8-column strips, so
        // accumulator column = strip * 64 + row * 8 + col and everything a strip contributes to one query's level-0 map is
        // one contiguous 128-byte run [8 rows][8 cols] -- a 64-byte DRAM fetch granule of the lookup is a 4 x 8 texel
        // block instead of a 1 x 32 texel tile row (6.9 instead of 14 granules per 10 x 10 window).
        // The strip is drained as two units of 32 columns (4 rows x 8 cols); the tcgen05.ld of unit u + 1 is in flight
        // while unit u is scaled, packed and pooled.  Level 1 (4 x 4 per strip), level 2 (2 x 2) and level 3 (1) are
        // the hierarchical 2 x 2 means of corr.py:28-30, kept in registers until the end of the tile; pooled texels
        // outside the level's map are written as zeros (the lookup stages windows without masking).
        uint32_t va[32], vb[32];
        uint32_t w0[16];                   // packed level-0 words of the strip's first unit (rows 0..3)
        uint32_t l1w[32], l2w[8];          // level-1 chunk [pair][row 4][col 8], level-2 chunk [row 2][col 8] of this tile
        float l3v[4];
        float l2a[2];                      // level-2 row 0 of the current strip (2 cols), waiting for level 3
        // valid pooled texels of this tile (floor semantics of avg_pool2d: maps are (h >> l) x (w >> l))
        const int nx1 = (p.w >> 1) - (bw0 >> 1), ny1 = (p.h >> 1) - (bh0 >> 1);
        const int nx2 = (p.w >> 2) - (bw0 >> 2), ny2 = (p.h >> 2) - (bh0 >> 2);
        const int nx3 = (p.w >> 3) - (bw0 >> 3), ny3 = (p.h >> 3) - (bh0 >> 3);
        const bool unit_scale = p.alpha == 1.0f;     // the flow net folds 1/sqrt(C) = 2^-4 into its feature maps (exact)
        // (Measured and dropped: sending every second box through coalesced 16-byte LSU stores instead of a TMA store -- 8 lanes
        // x 16 B per query row -- to use both store engines side by side: 2.09 vs 1.93 ms per 54 pairs, profiles/r02l_*.)
        const long long q_first = static_cast<long long>(batch) * p.n + qrow;
        tmem_ld_32x32(trow, va);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int strip = u >> 1, hh = u & 1;     // unit = rows 4 * hh .. 4 * hh + 3 of the strip
          uint32_t (&cur)[32] = (u & 1) ? vb : va;
          tmem_ld_wait();
          if (u < 7) tmem_ld_32x32(trow + (u + 1) * 32, (u & 1) ? va : vb);
          float c[32];
          if (unit_scale) {
#pragma unroll
            for (int j = 0; j < 32; ++j) c[j] = __uint_as_float(cur[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) c[j] = p.alpha * __uint_as_float(cur[j]);
          }
          // level 1: rows 2 * hh, 2 * hh + 1 of the strip's 4 x 4 block: ((a + b) + (c + d)) * 0.25
          float l1r[8];
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int x = 0; x < 4; ++x)
              l1r[r * 4 + x] = ((c[(2 * r) * 8 + 2 * x] + c[(2 * r) * 8 + 2 * x + 1]) + (c[(2 * r + 1) * 8 + 2 * x] + c[(2 * r + 1) * 8 + 2 * x + 1])) * 0.25f;
          // level 2: row hh of the strip's 2 x 2 block
          float l2r[2];
#pragma unroll
          for (int x = 0; x < 2; ++x) l2r[x] = ((l1r[2 * x] + l1r[2 * x + 1]) + (l1r[4 + 2 * x] + l1r[4 + 2 * x + 1])) * 0.25f;
          // zero the pooled texels outside the maps, then pack: level-1 word index = pair * 16 + row * 4 + (strip & 1) * 2 + x / 2
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const bool yok = 2 * hh + r < ny1;
            const float a0 = (yok && strip * 4 + 0 < nx1) ? l1r[r * 4 + 0] : 0.0f, a1 = (yok && strip * 4 + 1 < nx1) ? l1r[r * 4 + 1] : 0.0f;
            const float a2 = (yok && strip * 4 + 2 < nx1) ? l1r[r * 4 + 2] : 0.0f, a3 = (yok && strip * 4 + 3 < nx1) ? l1r[r * 4 + 3] : 0.0f;
            l1w[(strip >> 1) * 16 + (2 * hh + r) * 4 + (strip & 1) * 2] = pack_half2(a0, a1);
            l1w[(strip >> 1) * 16 + (2 * hh + r) * 4 + (strip & 1) * 2 + 1] = pack_half2(a2, a3);
          }
          {
            const bool yok = hh < ny2;
            l2w[hh * 4 + strip] = pack_half2((yok && strip * 2 < nx2) ? l2r[0] : 0.0f, (yok && strip * 2 + 1 < nx2) ? l2r[1] : 0.0f);
          }
          if (hh == 0) {
            l2a[0] = l2r[0];
            l2a[1] = l2r[1];
#pragma unroll
            for (int j = 0; j < 16; ++j) w0[j] = pack_half2(c[2 * j], c[2 * j + 1]);
          } else {
            const float l3 = ((l2a[0] + l2a[1]) + (l2r[0] + l2r[1])) * 0.25f;
            l3v[strip] = (0 < ny3 && strip < nx3) ? l3 : 0.0f;
            if (!(p.dbg & 2)) {
              // level-0 strip: 128 bytes per query = [8 rows][8 cols] fp16
              // 16-byte chunk j = row j, 128B swizzle, one TMA store of [32 queries][128 B].  (Measured and dropped: four 256-bit
              // stores per thread straight from the registers for every / every second strip -- 2.23 / 2.04 ms against 1.92 ms
              // for TMA only, profiles/r02r_exp_corr_direct_256bit_stores.txt: the store paths do not add up.)
              const int off = acquire();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                st_shared_v4(st_u32 + off + lane * 128 + ((static_cast<uint32_t>(j) ^ sw) << 4),
                             make_uint4(w0[4 * j], w0[4 * j + 1], w0[4 * j + 2], w0[4 * j + 3]));
                st_shared_v4(st_u32 + off + lane * 128 + ((static_cast<uint32_t>(j + 4) ^ sw) << 4),
                             make_uint4(pack_half2(c[8 * j], c[8 * j + 1]), pack_half2(c[8 * j + 2], c[8 * j + 3]),
                                        pack_half2(c[8 * j + 4], c[8 * j + 5]), pack_half2(c[8 * j + 6], c[8 * j + 7])));
              }
              commit(&p.tmL[0], off, strip * 64, t);
            }
          }
        }
        if (!(p.dbg & 3)) {
          int off = acquire();               // level 1: [pair 2][row 4][col 8] = 128 bytes per query, 128B swizzle
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(st_u32 + off + lane * 128 + ((static_cast<uint32_t>(j) ^ sw) << 4),
                         make_uint4(l1w[4 * j], l1w[4 * j + 1], l1w[4 * j + 2], l1w[4 * j + 3]));
          commit(&p.tmL[1], off, 0, t);
          off = acquire();                   // level 2: [row 2][col 8] = 32 bytes per query, 32B swizzle
#pragma unroll
          for (int r = 0; r < 2; ++r)
            st_shared_v4(st_u32 + off + lane * 32 + ((static_cast<uint32_t>(r) ^ (static_cast<uint32_t>(lane >> 2) & 1u)) << 4),
                         make_uint4(l2w[4 * r], l2w[4 * r + 1], l2w[4 * r + 2], l2w[4 * r + 3]));
          commit(&p.tmL[2], off, 0, t);
          // level 3: 4 fp16 = 8 bytes per query and tile, below the 16-byte TMA granularity: stored by the thread
          if (qrow + lane < p.n)
            *reinterpret_cast<uint2*>(p.l3 + ((q_first + lane) * p.tiles3 + (t / p.tiles_w) * p.tiles_w3 + t % p.tiles_w) * 4) =
                make_uint2(pack_half2(l3v[0], l3v[1]), pack_half2(l3v[2], l3v[3]));
        }
      } else {
      float hs1[16];   // horizontal pair sums of the previous (even) level-0 row
      float hs2[8];    // ... of the previous (even) level-1 row
      float hs3[4];    // ... of the previous (even) level-2 row
#pragma unroll
      for (int hl = 0; hl < 8; ++hl) {
        const int h = bh0 + hl;
        if (h >= p.h) break;                       // warp-uniform; rows past the grid only feed non-existent pooled rows
        uint32_t v[32];
        tmem_ld_32x32(trow + hl * 32, v);
        tmem_ld_wait();
        float c[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) c[j] = p.alpha * __uint_as_float(v[j]);
        if (!HALF && (p.dbg & 4)) {
          // LSU path: transpose through the swizzled slot, then 8 lanes write one 128-byte row segment
          const int off = (nstore & 1) * kCorrSlotBytes;
          ++nstore;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            st_shared_v4(st_u32 + off + lane * 128 + ((static_cast<uint32_t>(ch) ^ sw) << 4),
                         make_uint4(__float_as_uint(c[4 * ch]), __float_as_uint(c[4 * ch + 1]), __float_as_uint(c[4 * ch + 2]),
                                    __float_as_uint(c[4 * ch + 3])));
          __syncwarp();
          const int chn = lane & 7, r0 = lane >> 3;
          const bool colok = bw0 + chn * 4 < p.pitch0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = r0 + 4 * i;
            uint4 val;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                         : "r"(st_u32 + off + r * 128 + ((static_cast<uint32_t>(chn) ^ static_cast<uint32_t>(r & 7)) << 4)));
            if (colok && qrow + r < p.n)
              *reinterpret_cast<uint4*>(p.l0 + ((static_cast<long long>(batch) * p.n + qrow + r) * p.h + h) * p.pitch0 + bw0 + chn * 4) = val;
          }
        } else if (!(p.dbg & 2)) {
          const int off = acquire();
          if constexpr (HALF) {
            // 64 bytes per query, 64B swizzle: 16-byte chunk index ^= (row / 2) & 3 (conflict-free across 8 lanes)
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              st_shared_v4(st_u32 + off + lane * 64 + ((static_cast<uint32_t>(ch) ^ (static_cast<uint32_t>(lane >> 1) & 3u)) << 4),
                           make_uint4(pack_half2(c[8 * ch], c[8 * ch + 1]), pack_half2(c[8 * ch + 2], c[8 * ch + 3]),
                                      pack_half2(c[8 * ch + 4], c[8 * ch + 5]), pack_half2(c[8 * ch + 6], c[8 * ch + 7])));
          } else {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
              st_shared_v4(st_u32 + off + lane * 128 + ((static_cast<uint32_t>(ch) ^ sw) << 4),
                           make_uint4(__float_as_uint(c[4 * ch]), __float_as_uint(c[4 * ch + 1]), __float_as_uint(c[4 * ch + 2]),
                                      __float_as_uint(c[4 * ch + 3])));
          }
          commit(&p.tmL[0], off, bw0, h);
        }
        if ((hl & 1) == 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) hs1[j] = c[2 * j] + c[2 * j + 1];
        } else {
          float l1[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) l1[j] = (hs1[j] + (c[2 * j] + c[2 * j + 1])) * 0.25f;
          const int r1 = h >> 1;
          if (r1 < H1 && !(p.dbg & 3)) {
            const int off = acquire();
            if constexpr (HALF) {
              // 32 bytes per query, 32B swizzle: chunk index ^= (row / 4) & 1
#pragma unroll
              for (int ch = 0; ch < 2; ++ch)
                st_shared_v4(st_u32 + off + lane * 32 + ((static_cast<uint32_t>(ch) ^ (static_cast<uint32_t>(lane >> 2) & 1u)) << 4),
                             make_uint4(pack_half2(l1[8 * ch], l1[8 * ch + 1]), pack_half2(l1[8 * ch + 2], l1[8 * ch + 3]),
                                        pack_half2(l1[8 * ch + 4], l1[8 * ch + 5]), pack_half2(l1[8 * ch + 6], l1[8 * ch + 7])));
            } else {
#pragma unroll
              for (int ch = 0; ch < 4; ++ch)
                st_shared_v4(st_u32 + off + lane * 64 + ch * 16,
                             make_uint4(__float_as_uint(l1[4 * ch]), __float_as_uint(l1[4 * ch + 1]), __float_as_uint(l1[4 * ch + 2]),
                                        __float_as_uint(l1[4 * ch + 3])));
            }
            commit(&p.tmL[1], off, bw0 >> 1, r1);
          }
          if ((hl & 3) == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) hs2[j] = l1[2 * j] + l1[2 * j + 1];
          } else {
            float l2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) l2[j] = (hs2[j] + (l1[2 * j] + l1[2 * j + 1])) * 0.25f;
            const int r2 = h >> 2;
            if (r2 < H2 && !(p.dbg & 3)) {
              const int off = acquire();
              if constexpr (HALF) {
                st_shared_v4(st_u32 + off + lane * 16,
                             make_uint4(pack_half2(l2[0], l2[1]), pack_half2(l2[2], l2[3]), pack_half2(l2[4], l2[5]), pack_half2(l2[6], l2[7])));
              } else {
#pragma unroll
                for (int ch = 0; ch < 2; ++ch)
                  st_shared_v4(st_u32 + off + lane * 32 + ch * 16,
                               make_uint4(__float_as_uint(l2[4 * ch]), __float_as_uint(l2[4 * ch + 1]), __float_as_uint(l2[4 * ch + 2]),
                                          __float_as_uint(l2[4 * ch + 3])));
              }
              commit(&p.tmL[2], off, bw0 >> 2, r2);
            }
            if (hl == 3) {
#pragma unroll
              for (int j = 0; j < 4; ++j) hs3[j] = l2[2 * j] + l2[2 * j + 1];
            } else {
              const int r3 = h >> 3;
              const float l3a = (hs3[0] + (l2[0] + l2[1])) * 0.25f, l3b = (hs3[1] + (l2[2] + l2[3])) * 0.25f;
              const float l3c = (hs3[2] + (l2[4] + l2[5])) * 0.25f, l3d = (hs3[3] + (l2[6] + l2[7])) * 0.25f;
              if constexpr (HALF) {
                // 4 fp16 = 8 bytes per query: below the 16-byte TMA box granularity, so every thread stores its own
                // (1% of the pyramid bytes; neighbouring tiles complete the 32-byte sectors while they sit in L2)
                if (r3 < H3 && qrow + lane < p.n && !(p.dbg & 3))
                  *reinterpret_cast<uint2*>(p.l3 + ((static_cast<long long>(batch) * p.n + qrow + lane) * H3 + r3) * p.pitch3 + (bw0 >> 3)) =
                      make_uint2(pack_half2(l3a, l3b), pack_half2(l3c, l3d));
              } else if (r3 < H3 && !(p.dbg & 3)) {
                const int off = acquire();
                st_shared_v4(st_u32 + off + lane * 16,
                             make_uint4(__float_as_uint(l3a), __float_as_uint(l3b), __float_as_uint(l3c), __float_as_uint(l3d)));
                commit(&p.tmL[3], off, bw0 >> 3, r3);
              }
            }
          }
        }
      }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(&acc_empty[buf], 0);   // the leader's MMA warp owns the accumulators of both CTAs
        else mbar_arrive(&acc_empty[buf]);
      }
    }
    if (lane == 0) bulk_wait_read<0>();             // shared memory stays valid until the last store has read it
    __syncwarp();
  }

  tcgen05_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();            // neither CTA frees TMEM / exits while the pair's MMAs or arrives are in flight
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace atdn

using namespace atdn;

extern "C" int atdn_corr_pyramid(const void* fmap1, const void* fmap2, int64_t fmap_pitch, int32_t channels,
                                 void* const lvl[4], const int32_t lvl_pitch[4], int32_t half_levels, int32_t batch,
                                 int32_t h8, int32_t w8, float alpha, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(fmap1 && fmap2 && lvl && lvl_pitch && batch > 0, ATDN_ERR_ARG, "atdn_corr_pyramid: null / empty argument");
  ATDN_REQUIRE(half_levels == 0 || half_levels == 4, ATDN_ERR_UNSUP, "atdn_corr_pyramid: half_levels must be 0 (fp32 pyramid) or 4 (fp16 pyramid), got %d", half_levels);
  ATDN_REQUIRE(channels == 256, ATDN_ERR_UNSUP, "atdn_corr_pyramid: %d feature channels (the GMA feature net has 256)", channels);
  ATDN_REQUIRE(h8 >= 16 && w8 >= 16, ATDN_ERR_ARG, "atdn_corr_pyramid: grid %dx%d is smaller than 16x16 (level 3 would be < 2x2)", h8, w8);
  ATDN_REQUIRE(fmap_pitch >= channels && fmap_pitch % 8 == 0, ATDN_ERR_ALIGN, "atdn_corr_pyramid: fmap_pitch %lld", (long long)fmap_pitch);
  CorrParams p;
  memset(&p, 0, sizeof(p));
  const int n = h8 * w8;
  const bool pair = !env_switches().corr_no_pair;
  p.h = h8;
  p.w = w8;
  p.tiles_w = ceil_div(w8, 32);
  p.tiles = p.tiles_w * ceil_div(h8, 8);
  p.tiles_w3 = (p.tiles_w + 1) & ~1;
  p.tiles3 = p.tiles_w3 * ceil_div(h8, 8);
  p.alpha = alpha;
  p.dbg = env_switches().corr_dbg;
  p.l0 = static_cast<float*>(lvl[0]);
  p.l3 = static_cast<__half*>(lvl[3]);
  p.pitch3 = lvl_pitch[3];
  p.n = n;
  p.pitch0 = lvl_pitch[0];
  const uint32_t ones[4] = {1, 1, 1, 1};
  {
    const int64_t dims[4] = {channels, n, 1, batch};
    const int64_t str[3] = {fmap_pitch, (int64_t)n * fmap_pitch, (int64_t)n * fmap_pitch};
    const uint32_t box[4] = {64, 128, 1, 1};
    if (int e = make_map_f16(&p.tmA, fmap1, dims, str, box, ones, "fmap1")) return e;
  }
  {
    const int64_t dims[4] = {channels, w8, h8, batch};
    const int64_t str[3] = {fmap_pitch, (int64_t)w8 * fmap_pitch, (int64_t)n * fmap_pitch};
    // fp32 pyramid: whole 8 x 32 tile (a CTA of a pair stages 4 of its 8 rows); fp16 pyramid: 8 x 8 strips (see the kernel)
    const uint32_t box[4] = {64, half_levels ? 8u : 32u, half_levels ? 8u : (pair ? 4u : 8u), 1};
    if (int e = make_map_f16(&p.tmB, fmap2, dims, str, box, ones, "fmap2")) return e;
  }
  if (half_levels) {
    // strip fp16 layout (include/atdn_b200.h): level l = [batch * n, tiles, (8 >> l) * (32 >> l)], tiles = ceil(h8 / 8) * ceil(w8 / 32);
    // level 3 = [batch * n, ceil(h8 / 8) * tiles_w3, 4] with tiles_w3 = tiles_w rounded up to even (pad tiles stay zero)
    for (int l = 0; l < 4; ++l) {
      const int el = 256 >> (2 * l);
      ATDN_REQUIRE(lvl[l] != nullptr && aligned16(lvl[l]) && lvl_pitch[l] == el, ATDN_ERR_ALIGN,
                   "atdn_corr_pyramid: tiled level %d needs a 16-byte aligned buffer with %d elements per tile (got %d)", l, el, lvl_pitch[l]);
      if (l == 3) continue;                          // 8-byte tiles: stored without TMA
      const int64_t dims[4] = {el, p.tiles, n, batch};
      const int64_t str[3] = {el, (int64_t)p.tiles * el, (int64_t)n * p.tiles * el};
      const uint32_t box[4] = {l == 2 ? 16u : 64u, 1, 32, 1};
      if (int e = make_map(&p.tmL[l], 2, l == 2 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B, lvl[l], dims, str, box, ones,
                           "pyramid level")) return e;
    }
  } else {
    int hl = h8, wl = w8;
    for (int l = 0; l < 4; ++l) {
      ATDN_REQUIRE(lvl[l] != nullptr && lvl_pitch[l] % 4 == 0 && lvl_pitch[l] >= wl, ATDN_ERR_ALIGN, "atdn_corr_pyramid: level %d pitch %d", l, lvl_pitch[l]);
      const int64_t dims[4] = {wl, hl, n, batch};
      const int64_t str[3] = {lvl_pitch[l], (int64_t)hl * lvl_pitch[l], (int64_t)n * hl * lvl_pitch[l]};
      const uint32_t box[4] = {32u >> l, 1, 32, 1};
      // fp32 staging of the epilogue: level 0 = 128-byte rows (128B swizzle), the pooled levels are stored unswizzled
      if (int e = make_map(&p.tmL[l], 4, l == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, lvl[l], dims, str, box, ones,
                           "pyramid level")) return e;
      hl /= 2;
      wl /= 2;
    }
  }
  static DeviceOnce configured;
  if (configured.pending()) {
    ATDN_CUDA(cudaFuncSetAttribute(corr_pyramid_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCorrSmem));
    ATDN_CUDA(cudaFuncSetAttribute(corr_pyramid_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCorrSmem));
    ATDN_CUDA(cudaFuncSetAttribute(corr_pyramid_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCorrSmem));
    ATDN_CUDA(cudaFuncSetAttribute(corr_pyramid_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCorrSmem));
    configured.done();
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = pair ? dim3(2 * ceil_div(n, 256), batch) : dim3(ceil_div(n, 128), batch);
  cfg.blockDim = dim3(kCorrThreads);
  cfg.dynamicSmemBytes = kCorrSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (half_levels) {
    if (pair) ATDN_CUDA(cudaLaunchKernelEx(&cfg, corr_pyramid_kernel<true, true>, p));
    else ATDN_CUDA(cudaLaunchKernelEx(&cfg, corr_pyramid_kernel<true, false>, p));
  } else {
    if (pair) ATDN_CUDA(cudaLaunchKernelEx(&cfg, corr_pyramid_kernel<false, true>, p));
    else ATDN_CUDA(cudaLaunchKernelEx(&cfg, corr_pyramid_kernel<false, false>, p));
  }
  return 0;
}
