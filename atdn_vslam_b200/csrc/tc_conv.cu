// Persistent halo-reuse implicit-GEMM convolution for sm_100a (stride 1, NHWC fp16, K-major fp16 weights).
//
// Why a second kernel: the one-tile-per-CTA kernel in tc_gemm.cu re-fetches the 128-pixel A tile for every
// filter tap and the weight tile for every 128 pixels; at 64 FLOP per operand byte it runs at the L2->SM
// throughput cap, its single-thread MMA loop hands over only 256 tensor cycles per barrier round trip, and
// its epilogue is not overlapped (profiles/r02_*).  Here
//   * one CTA owns MT sub-tiles of 16 x 8 output pixels (M = 128 each) that sit side by side; for every
//     64-channel chunk ONE TMA box of (16 + kh - 1) x (8*MT + kw - 1) input pixels is staged and every filter
//     tap is a shifted view of it: the UMMA shared-memory descriptor starts (dy*WB + dx) pixel rows into the
//     box and strides 8-row groups by the box row pitch WB*128 B.  The 128B swizzle is a function of the
//     absolute shared-memory address, so row-shifted starts and group pitches that are not multiples of
//     1024 B read exactly what TMA wrote (verified by tools/umma_shift_test.cu on B200);
//   * each weight stage (BN x 64 fp16) feeds MT MMAs chains, so weight bytes per pixel drop MT-fold;
//   * the kernel is persistent (grid = #SMs, one contiguous tile range per CTA) with the 512 TMEM columns split into two
//     accumulator buffers: 8 epilogue warps drain tile i while the MMA warp accumulates tile i+1;
//   * A and B have independent producer threads and mbarrier rings (A advances per chunk, B per tap).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "tc_epilogue.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace atdn {

constexpr int kConvThreads = 384;   // warp 0: A producer, 1: B producer, 2: TMEM alloc + MMA, 3: idle, 4-11: epilogue
constexpr int kMaxAStages = 8;
constexpr int kMaxBStages = 12;
constexpr int kSubH = 16, kSubW = 8;   // one M = 128 sub-tile: 16 x 8 output pixels
constexpr int kMaxDynSmem = 232448 - 1024;   // 227 KiB per CTA minus the static shared memory (barriers)

struct alignas(64) ConvParams {
  CUtensorMap tmA, tmA2, tmB, tmOut;
  EpiParams e;
  int tiles_w, tiles_h, n_tiles, total_tiles;
  int out_h, out_w;
  int taps_h, taps_w, pad_h, pad_w;
  int chunks_a, chunks_a2, c_a, c_a2;
  int box_w;                         // halo box width in pixels
  int a_stage_bytes, a_tx_bytes, a_stages, b_stages;
  int stats_parts;                   // ATDN_F_STATS: partial-sum slots per image (0 = off)
  int out_tma;                       // STORE16: fp16 output boxes staged in shared memory and written by TMA stores
  int b_resident;                    // all (chunk, tap) weight tiles fit shared memory: loaded once per CTA, never recycled
  long long* stamps;                 // optional clock64 stamps of CTA 0 (timing experiments), or null
};

// SWIZZLE_128B K-major descriptor with an explicit 8-row-group pitch (stride byte offset)
__device__ __forceinline__ uint64_t make_smem_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

struct TileCoord {
  int h0, w0, n0, batch;
};
// t enumerates (batch, tile row, tile column group, N tile) with the N tile fastest; a CTA pair splits a
// column group of two tiles between its ranks.
template <int CL>
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int t, int rank, int tile_w_px, int bn) {
  TileCoord c;
  const int nt = t % p.n_tiles;
  int m = t / p.n_tiles;
  const int tw = m % p.tiles_w;
  m /= p.tiles_w;
  const int th = m % p.tiles_h;
  c.batch = m / p.tiles_h;
  c.h0 = th * kSubH;
  c.w0 = (tw * CL + rank) * tile_w_px;
  c.n0 = nt * bn;
  return c;
}

// PAIR = false: one CTA per tile, tcgen05.mma.cta_group::1 (M = 128 per sub-tile).
// PAIR = true : a cluster of two CTAs (the two SMs of a TPC) computes two horizontally adjacent tiles with
//               cta_group::2 MMAs of M = 256: each CTA stages its own halo box and HALF of the weight rows, so
//               the per-SM shared-memory operand traffic of one MMA drops from 4 KiB + 32 N to 4 KiB + 16 N
//               bytes -- the UMMA operand fetch sustains ~64 B/cycle/SM (measured), which caps cta_group::1 at
//               N/(N+128) of peak.  The leader CTA issues all MMAs; its commits are multicast to both CTAs.
template <int MT, int BN, int EPI, bool PAIR>
__global__ void __launch_bounds__(kConvThreads, 1) tc_conv_kernel(const __grid_constant__ ConvParams p) {
  static_assert(MT * BN <= 256, "two accumulator buffers must fit the 512 TMEM columns");
  static_assert(BN % 32 == 0, "epilogue works on 32-column chunks");
  constexpr int CL = PAIR ? 2 : 1;
  constexpr int kBBytes = (BN / CL) * 128;          // weight bytes staged per CTA and stage
  constexpr uint32_t kIdesc = make_idesc_f16(128 * CL, BN);
  constexpr int kChunksPerTile = MT * (BN / 32);

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
  __shared__ __align__(8) uint64_t b_full[kMaxBStages], b_empty[kMaxBStages];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + p.a_stages * p.a_stage_bytes;
  uint8_t* smem_out = smem_b + p.b_stages * kBBytes;   // out_tma: 8 epilogue warps x 2 slots x 2 KiB (1 KiB aligned)

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler as well
  const int lane = threadIdx.x & 31;
  const int rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
  // contiguous tile range per CTA (cluster): neighbouring halo boxes and the weight tiles stay hot in L2
  const int t_begin = static_cast<int>(static_cast<long long>(cid) * p.total_tiles / ncl);
  const int t_end = static_cast<int>(static_cast<long long>(cid + 1) * p.total_tiles / ncl);

  pdl_launch_dependents();     // the next kernel may set itself up while this grid drains (no-op without ATDN_PDL)
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8 * CL); }
    fence_barrier_init();
  }
  if (warp == 3 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if (p.chunks_a2 > 0) tma_prefetch_desc(&p.tmA2);
    if (p.out_tma) tma_prefetch_desc(&p.tmOut);
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_pair(&tmem_base_smem, 512);
    else tmem_alloc(&tmem_base_smem, 512);
  }
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers exist before any remote arrive / complete_tx
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int chunks = p.chunks_a + p.chunks_a2;
  const bool stamp = p.stamps != nullptr && blockIdx.x == 0;
  if (stamp && threadIdx.x == 0) p.stamps[0] = clock64();

  // Everything above, and the weight loads of warp 1 (constants), may overlap the tail of the preceding grid; the
  // activations, the recurrent state and every store of this kernel must not.
  if (warp != 1 && warp != 2) pdl_wait();

  if (warp == 0) {
    // ===== A producer (whole warp loops, one elected lane issues): one halo box per (tile, 64-channel chunk) =====
    int sa = 0;
    uint32_t pa = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const TileCoord tc = decode_tile<CL>(p, t, rank, kSubW * MT, BN);
      const int cw = tc.w0 - p.pad_w, ch = tc.h0 - p.pad_h;
      for (int c = 0; c < chunks; ++c) {
        mbar_wait(&a_empty[sa], pa ^ 1u);
        if (elect_one_sync()) {
          if (rank == 0) mbar_arrive_expect_tx(&a_full[sa], CL * p.a_tx_bytes);   // both CTAs' bytes land on the leader
          uint8_t* dst = smem + sa * p.a_stage_bytes;
          const CUtensorMap* tm = c < p.chunks_a ? &p.tmA : &p.tmA2;
          const int c0 = (c < p.chunks_a ? c : c - p.chunks_a) * 64;
          if constexpr (PAIR) tma_load_4d_pair(dst, tm, &a_full[sa], c0, cw, ch, tc.batch);
          else tma_load_4d(dst, tm, &a_full[sa], c0, cw, ch, tc.batch);
        }
        __syncwarp();
        if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== B producer: one (BN / CL) x 64 weight tile per (tile, chunk, tap) =====
    // b_resident: the whole filter of this N tile stays in shared memory (stage = chunk * taps + tap), so the weights
    // cross L2 -> SM once per CTA instead of once per tile (a 1x1 conv moves 2 weight bytes per activation byte otherwise)
    int sb = 0;
    uint32_t pb = 0;
    const int taps = p.taps_h * p.taps_w;
    for (int t = t_begin; t < t_end; ++t) {
      const int n0 = (t % p.n_tiles) * BN + rank * (BN / CL);
      for (int c = 0; c < chunks; ++c) {
        int k0 = c * 64;
        for (int tap = 0; tap < taps; ++tap, k0 += chunks * 64) {
          if (!p.b_resident) mbar_wait(&b_empty[sb], pb ^ 1u);
          if (elect_one_sync()) {
            if (rank == 0) mbar_arrive_expect_tx(&b_full[sb], CL * kBBytes);
            if constexpr (PAIR) tma_load_4d_pair(smem_b + sb * kBBytes, &p.tmB, &b_full[sb], k0, n0, 0, 0);
            else tma_load_4d(smem_b + sb * kBBytes, &p.tmB, &b_full[sb], k0, n0, 0, 0);
          }
          __syncwarp();
          if (++sb == p.b_stages) { sb = 0; pb ^= 1u; }
        }
      }
      if (p.b_resident) break;
    }
  } else if (warp == 2) {
    // ===== MMA issuer (leader CTA only in PAIR mode); warp-uniform loop, one elected lane issues =====
    if (rank == 0) {
      int sa = 0, sb = 0, buf = 0;
      uint32_t pa = 0, pb = 0, pacc0 = 0, pacc1 = 0;
      const uint32_t sbo = static_cast<uint32_t>(p.box_w) * 128u;
      const uint32_t smem_a_u32 = smem_u32(smem), smem_b_u32 = smem_u32(smem_b);
      int n_tile_iter = 0;
      for (int t = t_begin; t < t_end; ++t, ++n_tile_iter) {
        const uint32_t pacc = buf ? pacc1 : pacc0;
        mbar_wait(&acc_empty[buf], pacc ^ 1u);
        tcgen05_fence_after();
        if (buf) pacc1 ^= 1u; else pacc0 ^= 1u;
        const uint32_t d_base = tmem_base + static_cast<uint32_t>(buf * 256);
        uint32_t accumulate = 0;
        for (int c = 0; c < chunks; ++c) {
          const int rem = c < p.chunks_a ? p.c_a - c * 64 : p.c_a2 - (c - p.chunks_a) * 64;
          const int ksteps = rem >= 64 ? 4 : (rem + 15) >> 4;
          mbar_wait(&a_full[sa], pa);
          const uint64_t a_desc_stage = make_smem_desc_sw128_sbo(smem_a_u32 + sa * p.a_stage_bytes, sbo);
          uint32_t row_off = 0;   // (dy * WB + dx) * 128 B >> 4
          for (int dy = 0; dy < p.taps_h; ++dy, row_off += static_cast<uint32_t>(p.box_w - p.taps_w) * 8u) {
            for (int dx = 0; dx < p.taps_w; ++dx, row_off += 8u) {
              mbar_wait(&b_full[sb], p.b_resident ? 0u : pb);   // resident tiles complete phase 0 once and stay
              tcgen05_fence_after();
              const uint64_t a_desc = a_desc_stage + row_off;
              const uint64_t b_desc = make_smem_desc_sw128(smem_b_u32 + sb * kBBytes);
              if (elect_one_sync()) {
                // sub-tile s starts 8 pixel rows (8 * 128 B = 64 x 16 B) further; +32 B per 16-element K step
                if (ksteps == 4) {
#pragma unroll
                  for (int s = 0; s < MT; ++s) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      const uint32_t acc = (k == 0) ? accumulate : 1u;
                      if constexpr (PAIR) umma_f16_pair(d_base + s * BN, a_desc + static_cast<uint32_t>(s * 64 + 2 * k), b_desc + 2u * k, kIdesc, acc);
                      else umma_f16(d_base + s * BN, a_desc + static_cast<uint32_t>(s * 64 + 2 * k), b_desc + 2u * k, kIdesc, acc);
                    }
                  }
                } else {
#pragma unroll
                  for (int s = 0; s < MT; ++s) {
                    for (int k = 0; k < ksteps; ++k) {
                      const uint32_t acc = (k == 0) ? accumulate : 1u;
                      if constexpr (PAIR) umma_f16_pair(d_base + s * BN, a_desc + static_cast<uint32_t>(s * 64 + 2 * k), b_desc + 2u * k, kIdesc, acc);
                      else umma_f16(d_base + s * BN, a_desc + static_cast<uint32_t>(s * 64 + 2 * k), b_desc + 2u * k, kIdesc, acc);
                    }
                  }
                }
                if (!p.b_resident) {
                  if constexpr (PAIR) umma_commit_pair(&b_empty[sb]); else umma_commit(&b_empty[sb]);
                }
                if (dy == p.taps_h - 1 && dx == p.taps_w - 1) {
                  if constexpr (PAIR) umma_commit_pair(&a_empty[sa]); else umma_commit(&a_empty[sa]);
                  if (c == chunks - 1) {
                    if constexpr (PAIR) umma_commit_pair(&acc_full[buf]); else umma_commit(&acc_full[buf]);
                  }
                }
              }
              __syncwarp();
              accumulate = 1;
              if (++sb == p.b_stages) { sb = 0; pb ^= 1u; }
            }
          }
          if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
        }
        if (stamp && n_tile_iter < 8 && lane == 0) p.stamps[8 + n_tile_iter] = clock64();
        buf ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: warp w drains TMEM lanes 32*(w&3).. of every second 32-column chunk =====
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const int rh = r >> 3, rw = r & 7;
    int buf = 0;
    uint32_t pf0 = 0, pf1 = 0;
    int n_tile_iter = 0;
    int n_out = 0;                                            // out_tma: staging slots used so far by this warp
    for (int t = t_begin; t < t_end; ++t, ++n_tile_iter) {
      const TileCoord tc = decode_tile<CL>(p, t, rank, kSubW * MT, BN);
      const uint32_t pf = buf ? pf1 : pf0;
      mbar_wait(&acc_full[buf], pf);
      tcgen05_fence_after();
      if (buf) pf1 ^= 1u; else pf0 ^= 1u;
      if (stamp && threadIdx.x == 128 && n_tile_iter < 8) p.stamps[16 + n_tile_iter] = clock64();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * 256);
      const int h = tc.h0 + rh;
      constexpr bool kCanStats = EPI == ATDN_EPI_STORE16 && MT == 4 && BN == 64;
      float st_sum[kCanStats ? 32 : 1], st_sq[kCanStats ? 32 : 1];
      if constexpr (kCanStats) {
#pragma unroll
        for (int i = 0; i < 32; ++i) { st_sum[i] = 0.0f; st_sq[i] = 0.0f; }
      }
#pragma unroll 1
      for (int j = half; j < kChunksPerTile; j += 2) {
        const int s = j / (BN / 32), cc = j % (BN / 32);
        const int w = tc.w0 + s * kSubW + rw;
        const bool valid = (h < p.out_h) && (w < p.out_w);
        const long long pix = (static_cast<long long>(tc.batch) * p.out_h + h) * p.out_w + w;
        long long sidx = 0;
        if constexpr (EPI == ATDN_EPI_STORE16 || EPI == ATDN_EPI_STORE32 || EPI == ATDN_EPI_GRU_ZR || EPI == ATDN_EPI_GRU_Q)
          sidx = state_index(tc.batch, h, w, p.out_h, p.out_w);
        uint32_t v[32];
        tmem_ld_32x32(trow + s * BN + cc * 32, v);
        tmem_ld_wait();
        if constexpr (EPI == ATDN_EPI_STORE16 && MT == 4 && BN == 64) {
          // ATDN_F_STATS: this warp sees channel chunk cc == half in all four sub-tiles: running sums stay in registers
          if (p.stats_parts > 0 && valid) {
            const float4* b4 = reinterpret_cast<const float4*>(p.e.bias + tc.n0 + cc * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bb = __ldg(b4 + i);
              const float y0 = __uint_as_float(v[4 * i]) + bb.x, y1 = __uint_as_float(v[4 * i + 1]) + bb.y;
              const float y2 = __uint_as_float(v[4 * i + 2]) + bb.z, y3 = __uint_as_float(v[4 * i + 3]) + bb.w;
              st_sum[4 * i] += y0; st_sum[4 * i + 1] += y1; st_sum[4 * i + 2] += y2; st_sum[4 * i + 3] += y3;
              st_sq[4 * i] = fmaf(y0, y0, st_sq[4 * i]); st_sq[4 * i + 1] = fmaf(y1, y1, st_sq[4 * i + 1]);
              st_sq[4 * i + 2] = fmaf(y2, y2, st_sq[4 * i + 2]); st_sq[4 * i + 3] = fmaf(y3, y3, st_sq[4 * i + 3]);
            }
          }
        }
        if constexpr (EPI == ATDN_EPI_STORE16 || EPI == ATDN_EPI_GRU_Q || EPI == ATDN_EPI_GRU_ZR) {
          // fp16 outputs: out16 (STORE16, GRU_Q: the new hidden state) or r*h (GRU_ZR, accumulator columns 128..255)
          if (p.out_tma && (EPI != ATDN_EPI_GRU_ZR || tc.n0 + cc * 32 >= 128)) {
            // The warp's 32 accumulator rows are a 4 x 8 pixel patch: its [32 pixels][32 channels] fp16 chunk is one TMA
            // box {32, 8, 4, 1}.  Thread-per-row global stores touched 32 different rows per instruction (16 bytes each:
            // half sectors), which made every 1x1 convolution epilogue-bound (9K cycles per 128 x 256 tile vs 1-2.7K of MMAs).
            const int n = tc.n0 + cc * 32;
            if (n < p.e.n_valid) {                           // warp-uniform
              uint8_t* slot = smem_out + ((warp - 4) * 2 + (n_out & 1)) * 2048;
              if (lane == 0) bulk_wait_read<1>();            // the store issued two chunks ago has read this slot
              __syncwarp();
              epilogue_chunk<EPI>(p.e, valid, pix, n, v, sidx, smem_u32(slot) + lane * 64, static_cast<uint32_t>(lane >> 1) & 3u);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&p.tmOut, slot, EPI == ATDN_EPI_GRU_ZR ? n - 128 : n, tc.w0 + s * kSubW, tc.h0 + q * 4, tc.batch);
                bulk_commit();
              }
              ++n_out;
            }
            continue;
          }
        }
        epilogue_chunk<EPI>(p.e, valid, pix, tc.n0 + cc * 32, v, sidx);
      }
      if constexpr (kCanStats) {
        if (p.stats_parts > 0) {
          // transposing butterfly: after the step with stride d, lane bit d selects which half of the channels a lane keeps,
          // so lane l ends with the sums of channel l over the warp's 32 rows (x 4 sub-tiles)
#pragma unroll
          for (int d = 16, n = 16; d >= 1; d >>= 1, n >>= 1) {
            const bool up = (lane & d) != 0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
              const float send0 = up ? st_sum[i] : st_sum[i + n], keep0 = up ? st_sum[i + n] : st_sum[i];
              const float send1 = up ? st_sq[i] : st_sq[i + n], keep1 = up ? st_sq[i + n] : st_sq[i];
              st_sum[i] = keep0 + __shfl_xor_sync(0xffffffffu, send0, d);
              st_sq[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, d);
            }
          }
          const int tile_in_img = (t / p.n_tiles) % (p.tiles_w * p.tiles_h);
          const int part = (tile_in_img * CL + rank) * 4 + q;
          float2* dst = reinterpret_cast<float2*>(const_cast<float*>(p.e.aux32)) +
                        (static_cast<long long>(tc.batch) * p.stats_parts + part) * 64 + half * 32 + lane;
          *dst = make_float2(st_sum[0], st_sq[0]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(&acc_empty[buf], 0);   // the leader's MMA warp owns the accumulators of both CTAs
        else mbar_arrive(&acc_empty[buf]);
      }
      if (stamp && threadIdx.x == 128 && n_tile_iter < 8) p.stamps[24 + n_tile_iter] = clock64();
      buf ^= 1;
    }
    if (n_out > 0) {                                          // shared memory stays valid until the last store has read it
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // neither CTA frees TMEM / exits while the pair's MMAs or arrives are in flight
  if (warp == 2) {
    __syncwarp();
    tcgen05_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
  if (stamp && threadIdx.x == 0) p.stamps[1] = clock64();
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int MT, int BN, int EPI, bool PAIR>
static int launch_conv(const ConvParams& p, int smem, cudaStream_t stream) {
  static DeviceOnce configured;
  if (configured.pending()) {
    ATDN_CUDA(cudaFuncSetAttribute(tc_conv_kernel<MT, BN, EPI, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    configured.done();
  }
  const int cl = PAIR ? 2 : 1;
  const int max_clusters = num_sms() / cl;
  const int clusters = p.total_tiles < max_clusters ? p.total_tiles : max_clusters;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(clusters * cl);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = env_switches().pdl ? 2 : 1;
  ATDN_CUDA(cudaLaunchKernelEx(&cfg, tc_conv_kernel<MT, BN, EPI, PAIR>, p));
  return 0;
}

template <int EPI, bool PAIR>
static int dispatch_conv(int mt, int bn, const ConvParams& p, int smem, cudaStream_t s) {
  if (mt == 1 && bn == 256) return launch_conv<1, 256, EPI, PAIR>(p, smem, s);
  if (mt == 1 && bn == 192) return launch_conv<1, 192, EPI, PAIR>(p, smem, s);
  if (mt == 1 && bn == 128) return launch_conv<1, 128, EPI, PAIR>(p, smem, s);
  if (mt == 1 && bn == 64) return launch_conv<1, 64, EPI, PAIR>(p, smem, s);    // small problems (batch 1 at 1/8 resolution): more, smaller CTAs
  if (mt == 2 && bn == 128) return launch_conv<2, 128, EPI, PAIR>(p, smem, s);
  if (mt == 2 && bn == 96) return launch_conv<2, 96, EPI, PAIR>(p, smem, s);
  if (mt == 2 && bn == 64) return launch_conv<2, 64, EPI, PAIR>(p, smem, s);
  if (mt == 4 && bn == 64) return launch_conv<4, 64, EPI, PAIR>(p, smem, s);
  if (mt == 4 && bn == 32) return launch_conv<4, 32, EPI, PAIR>(p, smem, s);
  return set_error(ATDN_ERR_UNSUP, "atdn_tc_gemm: halo conv kernel has no (mt=%d, bn=%d) instance", mt, bn);
}
template <int EPI>
static int dispatch_conv(bool pair, int mt, int bn, const ConvParams& p, int smem, cudaStream_t s) {
  return pair ? dispatch_conv<EPI, true>(mt, bn, p, smem, s) : dispatch_conv<EPI, false>(mt, bn, p, smem, s);
}

int launch_conv_halo(const atdn_tc_desc* d, cudaStream_t stream) {
  ATDN_REQUIRE(d->a_mode == ATDN_MODE_PATCH && d->b_mode == ATDN_MODE_ROWS, ATDN_ERR_ARG, "atdn_tc_gemm: mt > 0 needs PATCH A and ROWS B");
  ATDN_REQUIRE(d->stride == 1, ATDN_ERR_UNSUP, "atdn_tc_gemm: the halo conv kernel is stride-1 only");
  ATDN_REQUIRE(d->taps_h >= 1 && d->taps_w >= 1 && d->out_h >= 1 && d->out_w >= 1, ATDN_ERR_ARG, "atdn_tc_gemm: bad conv geometry");
  ATDN_REQUIRE(d->out_pitch % 8 == 0 && d->out_ch_off % 8 == 0, ATDN_ERR_ALIGN, "atdn_tc_gemm: out_pitch/out_ch_off must be multiples of 8");
  const int mt = d->mt, bn = d->bn;
  const bool pair = (d->flags & ATDN_F_PAIR) != 0;
  const int cl = pair ? 2 : 1;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.e.n_valid = d->n_valid;
  p.e.flags = d->flags;
  p.e.alpha = d->alpha;
  p.e.bias = d->bias;
  p.e.out = d->out;
  p.e.out_pitch = d->out_pitch;
  p.e.out_ch_off = d->out_ch_off;
  p.e.resid = static_cast<const __half*>(d->resid16);
  p.e.resid_pitch = d->resid_pitch;
  p.e.resid_ch_off = d->resid_ch_off;
  p.e.h32 = d->h32;
  p.e.z32 = d->z32;
  p.e.rh16 = static_cast<__half*>(d->rh16);
  p.e.aux32 = d->aux32;
  p.e.gamma = d->gamma;
  p.e.img_w = d->out_w;
  p.e.img_h = d->out_h;
  p.stamps = reinterpret_cast<long long*>(d->lvl[2]);   // timing experiments only (normally null)
  if (d->flags & ATDN_F_STATS) {
    ATDN_REQUIRE(d->epi == ATDN_EPI_STORE16 && mt == 4 && bn == 64 && d->n_valid == 64 && d->aux32 && d->bias, ATDN_ERR_ARG,
                 "atdn_tc_gemm: F_STATS needs STORE16, mt 4, bn 64, 64 output channels, a bias and aux32 (the partial-sum buffer)");
    ATDN_REQUIRE(!(d->flags & (ATDN_F_RELU | ATDN_F_RESID | ATDN_F_FLOWTAIL | ATDN_F_TANH_LO)), ATDN_ERR_ARG,
                 "atdn_tc_gemm: F_STATS sums acc + bias (no activation flags)");
  }

  const int box_w = kSubW * mt + d->taps_w - 1, box_h = kSubH + d->taps_h - 1;
  ATDN_REQUIRE(box_w <= 256 && box_h <= 256, ATDN_ERR_UNSUP, "atdn_tc_gemm: halo box %d x %d exceeds the TMA box limit", box_w, box_h);
  const uint32_t box[4] = {64, (uint32_t)box_w, (uint32_t)box_h, 1};
  const uint32_t ones[4] = {1, 1, 1, 1};
  if (int e = make_map_f16(&p.tmA, d->a, d->a_dims, d->a_strides, box, ones, "A")) return e;
  const int64_t c_a = d->a_dims[0];
  p.c_a = (int)c_a;
  p.chunks_a = (d->a_split_chunk > 0) ? d->a_split_chunk : (int)((c_a + 63) / 64);
  if (d->a_split_chunk > 0) {
    ATDN_REQUIRE(d->a2 != nullptr, ATDN_ERR_ARG, "atdn_tc_gemm: a_split_chunk without a2");
    ATDN_REQUIRE(c_a == 64LL * d->a_split_chunk, ATDN_ERR_ARG, "atdn_tc_gemm: a must hold exactly a_split_chunk*64 channels");
    if (int e = make_map_f16(&p.tmA2, d->a2, d->a2_dims, d->a2_strides, box, ones, "A2")) return e;
    p.chunks_a2 = (int)((d->a2_dims[0] + 63) / 64);
    p.c_a2 = (int)d->a2_dims[0];
  }
  const int chunks = p.chunks_a + p.chunks_a2;
  ATDN_REQUIRE(d->b_dims[0] >= (int64_t)(d->taps_h * d->taps_w * chunks - 1) * 64 + 1, ATDN_ERR_ARG,
               "atdn_tc_gemm: B has K extent %lld but the A side iterates %d chunks of 64", (long long)d->b_dims[0],
               d->taps_h * d->taps_w * chunks);
  {
    const uint32_t bbox[4] = {64, (uint32_t)(bn / cl), 1, 1};
    if (int e = make_map_f16(&p.tmB, d->b, d->b_dims, d->b_strides, bbox, ones, "B")) return e;
  }
  p.taps_h = d->taps_h;
  p.taps_w = d->taps_w;
  p.pad_h = d->pad_h;
  p.pad_w = d->pad_w;
  p.out_h = d->out_h;
  p.out_w = d->out_w;
  p.box_w = box_w;
  p.tiles_w = ceil_div(d->out_w, kSubW * mt * cl);   // column groups: a CTA pair owns two adjacent tiles
  p.tiles_h = ceil_div(d->out_h, kSubH);
  p.n_tiles = ceil_div(d->n_valid, bn);
  p.total_tiles = p.tiles_w * p.tiles_h * p.n_tiles * (int)d->a_dims[3];
  p.stats_parts = (d->flags & ATDN_F_STATS) ? p.tiles_w * p.tiles_h * cl * 4 : 0;
  p.a_tx_bytes = box_w * box_h * 128;
  p.a_stage_bytes = (p.a_tx_bytes + 1023) / 1024 * 1024;
  const int b_bytes = (bn / cl) * 128;
  {
    const bool off = env_switches().no_out_tma;
    const void* o16 = d->epi == ATDN_EPI_GRU_ZR ? d->rh16 : d->out;
    p.out_tma = (d->epi == ATDN_EPI_STORE16 || d->epi == ATDN_EPI_GRU_Q || d->epi == ATDN_EPI_GRU_ZR) && o16 != nullptr &&
                !off && aligned16(o16);
  }
  const int out_bytes = p.out_tma ? 8 * 2 * 2048 + 1024 : 0;   // staging slots (+ alignment of the slot base to 1 KiB)
  const int avail = kMaxDynSmem - 1024 - out_bytes;   // 1 KiB alignment slack
  p.a_stages = (3 * p.a_stage_bytes + 4 * b_bytes <= avail) ? 3 : 2;
  if (d->taps_h * d->taps_w == 1) {
    // 1x1: one A box per weight tile, both rings drain at the same rate: split shared memory evenly (measured neutral
    // against 3 A stages -- the 1x1 layers were bound by their epilogue stores, see the TMA-store epilogue below)
    const int n = avail / (p.a_stage_bytes + b_bytes);
    p.a_stages = n > kMaxAStages ? kMaxAStages : (n < 2 ? 2 : n);
  }
  int bs = (avail - p.a_stages * p.a_stage_bytes) / b_bytes;
  p.b_stages = bs > kMaxBStages ? kMaxBStages : bs;
  {
    // resident weights (opt-in, ATDN_B_RESIDENT=1): every (chunk, tap) tile of the single N tile in shared memory next to
    // >= 2 A stages.  Measured on the 1x1 and 64-channel layers: no gain (the weight stream was not their limiter, and
    // the 1x1 324 -> 256 layer loses A stages: 77 vs 74 us), so the streamed ring stays the default.
    const int need = chunks * d->taps_h * d->taps_w;
    if (env_switches().b_resident && p.n_tiles == 1 && need <= kMaxBStages && 2 * p.a_stage_bytes + need * b_bytes <= avail) {
      p.b_resident = 1;
      p.b_stages = need;
      const int as = (avail - need * b_bytes) / p.a_stage_bytes;
      p.a_stages = as > kMaxAStages ? kMaxAStages : as;
    }
  }
  ATDN_REQUIRE(p.b_stages >= 2, ATDN_ERR_UNSUP, "atdn_tc_gemm: halo conv (mt=%d, bn=%d, taps %dx%d) does not fit shared memory", mt, bn, d->taps_h, d->taps_w);
  const int smem = p.a_stages * p.a_stage_bytes + p.b_stages * b_bytes + 1024 + out_bytes;

  if (p.out_tma) {
    const bool zr = d->epi == ATDN_EPI_GRU_ZR;                 // r*h: dense [pix, 128]
    const int64_t opitch = zr ? 128 : d->out_pitch;
    const int64_t odims[4] = {zr ? 128 : d->n_valid, d->out_w, d->out_h, d->a_dims[3]};
    const int64_t ostr[3] = {opitch, (int64_t)d->out_w * opitch, (int64_t)d->out_h * d->out_w * opitch};
    const uint32_t obox[4] = {32, (uint32_t)kSubW, 4, 1};
    const __half* obase = zr ? static_cast<const __half*>(d->rh16) : static_cast<const __half*>(d->out) + d->out_ch_off;
    if (int e = make_map(&p.tmOut, 2, CU_TENSOR_MAP_SWIZZLE_64B, obase, odims, ostr, obox, ones, "conv output")) return e;
  }
  switch (d->epi) {
    case ATDN_EPI_STORE16:
      ATDN_REQUIRE(!(d->flags & ATDN_F_RESID) || (d->resid16 && d->resid_pitch % 8 == 0 && d->resid_ch_off % 8 == 0 && d->n_valid % 8 == 0), ATDN_ERR_ARG, "atdn_tc_gemm: residual");
      ATDN_REQUIRE(!(d->flags & ATDN_F_FLOWTAIL) || d->aux32, ATDN_ERR_ARG, "atdn_tc_gemm: FLOWTAIL needs aux32");
      ATDN_REQUIRE(!(d->flags & ATDN_F_TANH_LO) || d->h32, ATDN_ERR_ARG, "atdn_tc_gemm: TANH_LO needs h32");
      ATDN_REQUIRE(d->out != nullptr, ATDN_ERR_ARG, "atdn_tc_gemm: null output");
      return dispatch_conv<ATDN_EPI_STORE16>(pair, mt, bn, p, smem, stream);
    case ATDN_EPI_STORE32:
      ATDN_REQUIRE(d->out != nullptr, ATDN_ERR_ARG, "atdn_tc_gemm: null output");
      return dispatch_conv<ATDN_EPI_STORE32>(pair, mt, bn, p, smem, stream);
    case ATDN_EPI_GRU_ZR:
      ATDN_REQUIRE(d->n_valid == 256 && (bn == 128 || bn == 256) && d->h32 && d->z32 && d->rh16, ATDN_ERR_ARG, "atdn_tc_gemm: GRU_ZR arguments (bn must be 128 or 256)");
      return dispatch_conv<ATDN_EPI_GRU_ZR>(pair, mt, bn, p, smem, stream);
    case ATDN_EPI_GRU_Q:
      ATDN_REQUIRE(d->n_valid == 128 && d->h32 && d->z32 && d->out, ATDN_ERR_ARG, "atdn_tc_gemm: GRU_Q arguments");
      return dispatch_conv<ATDN_EPI_GRU_Q>(pair, mt, bn, p, smem, stream);
    case ATDN_EPI_FLOW:
      ATDN_REQUIRE(d->n_valid == 2 && bn == 32 && mt == 4 && !pair && d->h32 && d->z32, ATDN_ERR_ARG, "atdn_tc_gemm: FLOW arguments (n_valid 2, mt 4, bn 32, no pair)");
      return launch_conv<4, 32, ATDN_EPI_FLOW, false>(p, smem, stream);
    default:
      return set_error(ATDN_ERR_UNSUP, "atdn_tc_gemm: epilogue %d is not available in the halo conv kernel", d->epi);
  }
}

}  // namespace atdn
