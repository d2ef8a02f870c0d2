// Tensor-core implicit GEMM for sm_100a: TMA-staged fp16 operands (128B swizzle), tcgen05.mma with the
// fp32 accumulator in TMEM, warp-specialised roles, fused epilogues.  See include/atdn_b200.h.
//
// CTA = 192 threads: warps 0-3 epilogue (warp w owns TMEM lanes 32w..32w+31 = tile rows), warp 4 = TMA
// producer (one lane), warp 5 = TMEM allocator + MMA issuer (one lane).  One output tile of
// 128 x BN per CTA; the K loop runs over (filter tap, 64-channel chunk) pairs through a STAGES-deep
// mbarrier ring.  Two CTAs per SM are resident for BN <= 128, so one tile's epilogue overlaps the
// other's MMA phase.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.h"
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"
#include "tc_host.cuh"

namespace atdn {

constexpr int kTileM = 128;
constexpr int kChunkK = 64;                  // fp16 elements per 128-byte swizzled row
constexpr int kABytes = kTileM * kChunkK * 2;  // 16 KiB
constexpr int kThreads = 192;

struct alignas(64) TcParams {
  CUtensorMap tmA, tmA2, tmB, tmOut, tmB8;
  int a_mode, tiles_w, out_h, out_w, m_rows;
  int out_tma;            // STORE16 (single-CTA kernel): fp16 output boxes staged in shared memory, written by TMA stores
  int taps_w, pad_h, pad_w, stride;
  int chunks_a, chunks_a2, c_a, c_a2, num_k_iters;
  int b_batched, a_shared;
  int a_tiled, a_row_blocks;   // ATDN_F_A_TILED: A in blocks of 32 rows x 64 columns, ceil(rows / 32) row blocks per batch element
  const uint8_t* a_hot;        // ATDN_F_A_MIXED: per (batch, 256-row pair tile, 64-column block) 1 = fp16 block, 0 = e4m3 block (tmA2 / tmB8)
  int corr_h, corr_w, corr_tiles_w;
  int lvl_pitch[4];
  float* lvl[3];
  EpiParams e;
};

// Corr-volume epilogue: this thread owns query row `pix` of a tile of 8 x 32 target pixels whose 256
// accumulator columns are ordered (h_local, w_local).  Level 0 is stored as is; levels 1..3 are the
// hierarchical 2x2 means of corr.py:28-30 (floor semantics: only complete 2x2 blocks exist).
__device__ __forceinline__ void epilogue_corr(const TcParams& p, bool valid, long long pix, int bh0, int bw0,
                                              uint32_t tmem_row_base) {
  const int H0 = p.corr_h, W0 = p.corr_w;
  const int H1 = H0 / 2, W1 = W0 / 2, H2 = H1 / 2, W2 = W1 / 2, H3 = H2 / 2, W3 = W2 / 2;
  float hs1[16];   // horizontal pair sums of the previous (even) level-0 row
  float hs2[8];    // horizontal pair sums of the previous (even) level-1 row
  float hs3[4];    // horizontal pair sums of the previous (even) level-2 row
#pragma unroll
  for (int hl = 0; hl < 8; ++hl) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_row_base + hl * 32, v);
    tmem_ld_wait();
    float c[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) c[j] = p.e.alpha * __uint_as_float(v[j]);
    const int h = bh0 + hl;
    if (valid && h < H0) {
      float* dst = reinterpret_cast<float*>(p.e.out) + (pix * H0 + h) * (long long)p.lvl_pitch[0] + bw0;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        if (bw0 + j < p.lvl_pitch[0]) *reinterpret_cast<float4*>(dst + j) = make_float4(c[j], c[j + 1], c[j + 2], c[j + 3]);
    }
    if ((hl & 1) == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) hs1[j] = c[2 * j] + c[2 * j + 1];
    } else {
      float l1[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) l1[j] = (hs1[j] + (c[2 * j] + c[2 * j + 1])) * 0.25f;
      const int r1 = h >> 1, c1 = bw0 >> 1;
      if (valid && r1 < H1) {
        float* dst = p.lvl[0] + (pix * H1 + r1) * (long long)p.lvl_pitch[1] + c1;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          if (c1 + j < p.lvl_pitch[1]) *reinterpret_cast<float4*>(dst + j) = make_float4(l1[j], l1[j + 1], l1[j + 2], l1[j + 3]);
      }
      if ((hl & 3) == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) hs2[j] = l1[2 * j] + l1[2 * j + 1];
      } else {
        float l2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) l2[j] = (hs2[j] + (l1[2 * j] + l1[2 * j + 1])) * 0.25f;
        const int r2 = h >> 2, c2 = bw0 >> 2;
        if (valid && r2 < H2) {
          float* dst = p.lvl[1] + (pix * H2 + r2) * (long long)p.lvl_pitch[2] + c2;
#pragma unroll
          for (int j = 0; j < 8; j += 4)
            if (c2 + j < p.lvl_pitch[2]) *reinterpret_cast<float4*>(dst + j) = make_float4(l2[j], l2[j + 1], l2[j + 2], l2[j + 3]);
        }
        if (hl == 3) {
#pragma unroll
          for (int j = 0; j < 4; ++j) hs3[j] = l2[2 * j] + l2[2 * j + 1];
        } else {
          const int r3 = h >> 3, c3 = bw0 >> 3;
          if (valid && r3 < H3 && c3 < p.lvl_pitch[3]) {
            float* dst = p.lvl[2] + (pix * H3 + r3) * (long long)p.lvl_pitch[3] + c3;
            *reinterpret_cast<float4*>(dst) =
                make_float4((hs3[0] + (l2[0] + l2[1])) * 0.25f, (hs3[1] + (l2[2] + l2[3])) * 0.25f,
                            (hs3[2] + (l2[4] + l2[5])) * 0.25f, (hs3[3] + (l2[6] + l2[7])) * 0.25f);
          }
        }
      }
    }
  }
  (void)W1; (void)W2; (void)W3;
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(kThreads) tc_gemm_kernel(const __grid_constant__ TcParams p) {
  constexpr int kBBytes = BN * kChunkK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  constexpr uint32_t kIdesc = make_idesc_f16(kTileM, BN);

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler as well
  const int lane = threadIdx.x & 31;
  const int batch = blockIdx.z;

  // tile coordinates
  int m0 = 0, h0 = 0, w0 = 0;
  if (p.a_mode == ATDN_MODE_PATCH) {
    h0 = (blockIdx.x / p.tiles_w) * 8;
    w0 = (blockIdx.x % p.tiles_w) * 16;
  } else {
    m0 = blockIdx.x * kTileM;
  }
  const int n0 = blockIdx.y * BN;
  int bh0 = 0, bw0 = 0;
  if constexpr (EPI == ATDN_EPI_CORR) {
    bh0 = (blockIdx.y / p.corr_tiles_w) * 8;
    bw0 = (blockIdx.y % p.corr_tiles_w) * 32;
  }

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if (p.chunks_a2 > 0) tma_prefetch_desc(&p.tmA2);
  }
  if (warp == 5) tmem_alloc(&tmem_base_smem, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();                  // operands, bitmaps and outputs are global memory: not before the preceding grid is complete

  const int chunks = p.chunks_a + p.chunks_a2;

  if (warp == 4) {
    // ===== TMA producer: the whole warp runs the loop, one elected lane issues (see elect_one_sync) =====
    int stage = 0;
    uint32_t phase = 0;
    int chunk = 0, dx = 0, dy = 0;
    for (int it = 0; it < p.num_k_iters; ++it) {
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
        uint8_t* sA = smem + stage * kStageBytes;
        uint8_t* sB = sA + kABytes;
        if (p.a_mode == ATDN_MODE_PATCH) {
          const int cw = w0 * p.stride + dx - p.pad_w;      // (chunk, dx, dy) advance incrementally: no
          const int chh = h0 * p.stride + dy - p.pad_h;     // integer division inside the K loop
          if (chunk < p.chunks_a) tma_load_4d(sA, &p.tmA, &full_bar[stage], chunk * kChunkK, cw, chh, batch);
          else tma_load_4d(sA, &p.tmA2, &full_bar[stage], (chunk - p.chunks_a) * kChunkK, cw, chh, batch);
        } else {
          if (p.a_tiled) tma_load_4d(sA, &p.tmA, &full_bar[stage], 0, 0, it, batch * p.a_row_blocks + (m0 >> 5));
          else tma_load_4d(sA, &p.tmA, &full_bar[stage], it * kChunkK, m0, 0, p.a_shared ? 0 : batch);
        }
        if constexpr (EPI == ATDN_EPI_CORR) {
          tma_load_4d(sB, &p.tmB, &full_bar[stage], it * kChunkK, bw0, bh0, batch);
        } else {
          tma_load_4d(sB, &p.tmB, &full_bar[stage], it * kChunkK, n0, 0, p.b_batched ? batch : 0);
        }
      }
      __syncwarp();
      if (++chunk == chunks) { chunk = 0; if (++dx == p.taps_w) { dx = 0; ++dy; } }
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 5) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues =====
    int stage = 0, mchunk = 0;
    uint32_t phase = 0;
    const uint32_t smem_base_u32 = smem_u32(smem);
    for (int it = 0; it < p.num_k_iters; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      // valid 16-wide K steps in this chunk (zero-padded tails are skipped)
      int rem;
      if (p.a_mode == ATDN_MODE_PATCH) {
        rem = mchunk < p.chunks_a ? p.c_a - mchunk * kChunkK : p.c_a2 - (mchunk - p.chunks_a) * kChunkK;
        if (++mchunk == chunks) mchunk = 0;
      } else {
        rem = p.c_a - it * kChunkK;
      }
      const int ksteps = rem >= kChunkK ? 4 : (rem + 15) >> 4;
      const uint32_t a_addr = smem_base_u32 + stage * kStageBytes;
      const uint64_t a_desc = make_smem_desc_sw128(a_addr);
      const uint64_t b_desc = make_smem_desc_sw128(a_addr + kABytes);
      const uint32_t acc0 = it > 0 ? 1u : 0u;
      if (elect_one_sync()) {
        if (ksteps == 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k)   // +32 bytes (>>4 = 2) per 16-element K step inside the 128-byte swizzle atom
            umma_f16(tmem_base, a_desc + 2u * k, b_desc + 2u * k, kIdesc, k == 0 ? acc0 : 1u);
        } else {
          for (int k = 0; k < ksteps; ++k)
            umma_f16(tmem_base, a_desc + 2u * k, b_desc + 2u * k, kIdesc, k == 0 ? acc0 : 1u);
        }
        umma_commit(&empty_bar[stage]);
        if (it == p.num_k_iters - 1) umma_commit(&tmem_full_bar);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ===== epilogue warps 0..3 =====
    const int r = warp * 32 + lane;
    bool valid;
    long long pix;
    if (p.a_mode == ATDN_MODE_PATCH) {
      const int h = h0 + (r >> 4), w = w0 + (r & 15);
      valid = (h < p.out_h) && (w < p.out_w);
      pix = (static_cast<long long>(batch) * p.out_h + h) * p.out_w + w;
    } else {
      valid = (m0 + r) < p.m_rows;
      pix = static_cast<long long>(batch) * p.m_rows + m0 + r;
    }
    mbar_wait_group128(&tmem_full_bar, 0, threadIdx.x);
    tcgen05_fence_after();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    if constexpr (EPI == ATDN_EPI_CORR) {
      epilogue_corr(p, valid, pix, bh0, bw0, trow);
    } else {
      int n_out = 0;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(trow + c * 32, v);
        tmem_ld_wait();
        if constexpr (EPI == ATDN_EPI_STORE16) {
          // Same staging as the halo kernel (tc_conv.cu): the warp's [32 rows][32 channels] fp16 chunk is one TMA box
          // (PATCH: 2 image rows x 16 pixels; ROWS: 32 GEMM rows).  All MMAs of this CTA have completed, so the
          // operand ring at the start of shared memory is free to hold the 4 x 2 staging slots.
          if (p.out_tma && n0 + c * 32 < p.e.n_valid) {
            uint8_t* slot = smem + (warp * 2 + (n_out & 1)) * 2048;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            epilogue_chunk<EPI>(p.e, valid, pix, n0 + c * 32, v, -1, smem_u32(slot) + lane * 64, static_cast<uint32_t>(lane >> 1) & 3u);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (p.a_mode == ATDN_MODE_PATCH) tma_store_4d(&p.tmOut, slot, n0 + c * 32, w0, h0 + warp * 2, batch);
              else tma_store_4d(&p.tmOut, slot, n0 + c * 32, m0 + warp * 32, 0, batch);
              bulk_commit();
            }
            ++n_out;
            continue;
          }
        }
        epilogue_chunk<EPI>(p.e, valid, pix, n0 + c * 32, v);
      }
      if (n_out > 0) {
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair kernel: one cluster of two CTAs computes a 256 x BN tile with tcgen05.mma.cta_group::2.
// Each CTA stages its own 128 A rows and BN/2 of the B rows, so the operand bytes fetched from L2 per
// FLOP drop by 2x (BN = 256) relative to the single-CTA 128 x 128 tile -- the single-CTA kernel runs at
// the chip's L2->SM throughput cap (profiles/r01_prof_gru_zr_details.txt).  Only the leader CTA
// issues MMAs; its commits are multicast to the mbarriers of both CTAs.
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads) tc_gemm2_kernel(const __grid_constant__ TcParams p) {
  constexpr int kBHalfBytes = (BN / 2) * kChunkK * 2;
  constexpr int kStageBytes = kABytes + kBHalfBytes;
  // P.V keeps a second accumulator half for the lo plane of the e4m3 blocks (ATDN_F_A_MIXED): columns [BN, 2 BN)
  constexpr int kAccN = (EPI == ATDN_EPI_PV && (BN == 128 || BN == 64)) ? 2 * BN : BN;
  constexpr uint32_t kTmemCols = kAccN <= 32 ? 32 : kAccN <= 64 ? 64 : kAccN <= 128 ? 128 : 256;
  constexpr uint32_t kIdesc = make_idesc_f16(2 * kTileM, BN);

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler as well
  const int lane = threadIdx.x & 31;
  const int batch = blockIdx.z;
  const int rank = static_cast<int>(cluster_ctarank());
  const int pair = blockIdx.x >> 1;

  int m0 = 0, h0 = 0, w0 = 0;
  if (p.a_mode == ATDN_MODE_PATCH) {
    h0 = (pair / p.tiles_w) * 16 + rank * 8;
    w0 = (pair % p.tiles_w) * 16;
  } else {
    m0 = pair * (2 * kTileM) + rank * kTileM;
  }
  const int n0 = blockIdx.y * BN;
  int bh0 = 0, bw0 = 0;
  if constexpr (EPI == ATDN_EPI_CORR) {
    bh0 = (blockIdx.y / p.corr_tiles_w) * 8;
    bw0 = (blockIdx.y % p.corr_tiles_w) * 32;
  }

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if (p.chunks_a2 > 0) tma_prefetch_desc(&p.tmA2);
  }
  if constexpr (EPI == ATDN_EPI_PV && (BN == 128 || BN == 64)) {
    if (p.a_hot != nullptr) {   // 8 KiB of zeros behind the ring: the operand tile of the accumulator-clearing MMA
      uint4* z = reinterpret_cast<uint4*>(smem + STAGES * kStageBytes);
      for (int i = threadIdx.x; i < 8192 / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
      fence_proxy_async_smem();
    }
  }
  if (warp == 5) tmem_alloc_pair(&tmem_base_smem, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers are initialised before any remote arrive / complete_tx
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();                  // operands, bitmaps and outputs are global memory: not before the preceding grid is complete

  const int chunks = p.chunks_a + p.chunks_a2;

  // ATDN_F_A_MIXED on CTA pairs: one bitmap row per 256-row tile (atdn_attn_harmonize); an e4m3 block stages 8 KiB of A per
  // CTA and ONE e4m3 plane of the B tile (CTA 0 the hi plane, CTA 1 the lo plane: BN x 64 bytes) -- 16 KiB per CTA and block
  // through the L2 -> SM path instead of the 32 KiB of the single-CTA fp16 kernel.  The pair's MMA then has N = 2 BN: B rows
  // [0, BN) come from CTA 0 (hi -> accumulator columns [0, BN), where the fp16 blocks accumulate as well) and rows
  // [BN, 2 BN) from CTA 1 (lo -> columns [BN, 2 BN)); the epilogue adds the halves.  One N = 2 BN instruction per 32
  // columns reads the A operand from shared memory once for both planes (two N = BN instructions: 96 instead of 64
  // bytes per clock, the regime in which every N = 128 kernel here stops at ~0.65 of the tensor peak).
  bool mixed = false;
  const uint8_t* hot_row = nullptr;
  if constexpr (EPI == ATDN_EPI_PV && (BN == 128 || BN == 64)) {
    mixed = p.a_hot != nullptr;
    if (mixed) hot_row = p.a_hot + (static_cast<long long>(batch) * (gridDim.x >> 1) + pair) * p.num_k_iters;
  }

  if (warp == 4 && mixed) {
    if constexpr (EPI == ATDN_EPI_PV && (BN == 128 || BN == 64)) {
      int stage = 0;
      uint32_t phase = 0, hot32 = 0;
      const int rb = batch * p.a_row_blocks + (m0 >> 5);
      for (int it = 0; it < p.num_k_iters; ++it) {
        if ((it & 31) == 0) hot32 = (it + lane < p.num_k_iters) ? hot_row[it + lane] : 0u;   // 32 blocks of the bitmap row, one per lane
        const uint32_t hot = __shfl_sync(0xffffffffu, hot32, it & 31);
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (elect_one_sync()) {
          uint8_t* sA = smem + stage * kStageBytes;
          uint8_t* sB = sA + kABytes;
          if (hot) {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
            tma_load_4d_pair(sA, &p.tmA, &full_bar[stage], 0, 0, it, rb);
            tma_load_4d_pair(sB, &p.tmB, &full_bar[stage], it * kChunkK, n0 + rank * (BN / 2), 0, batch);
          } else {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (kABytes / 2 + kBHalfBytes));
            tma_load_4d_pair(sA, &p.tmA2, &full_bar[stage], 0, 0, it, rb);
            tma_load_4d_pair(sB, &p.tmB8, &full_bar[stage], it * kChunkK, n0, rank, batch);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 5 && mixed) {
    if constexpr (EPI == ATDN_EPI_PV && (BN == 128 || BN == 64)) {
      if (rank == 0) {
        int stage = 0;
        uint32_t phase = 0, hot32 = 0;
        const uint32_t smem_base_u32 = smem_u32(smem);
        constexpr uint32_t kIdesc2 = make_idesc_f16(2 * kTileM, 2 * BN);
        // both accumulator halves start at zero: one MMA over the zeroed operand tile behind the ring (a first fp16 block
        // would leave the lo half uninitialised, and one instruction cannot overwrite one half and accumulate into the other)
        const uint64_t z_desc = make_smem_desc_sw64(smem_base_u32 + STAGES * kStageBytes);
        if (elect_one_sync()) umma_f8_pair(tmem_base, z_desc, z_desc, kIdesc2, 0u);
        __syncwarp();
        for (int it = 0; it < p.num_k_iters; ++it) {
          if ((it & 31) == 0) hot32 = (it + lane < p.num_k_iters) ? hot_row[it + lane] : 0u;
          const uint32_t hot = __shfl_sync(0xffffffffu, hot32, it & 31);
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_base_u32 + stage * kStageBytes;
          if (hot) {
            const uint64_t a_desc = make_smem_desc_sw128(a_addr);
            const uint64_t b_desc = make_smem_desc_sw128(a_addr + kABytes);
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_pair(tmem_base, a_desc + 2u * k, b_desc + 2u * k, kIdesc, 1u);
              umma_commit_pair(&empty_bar[stage]);
              if (it == p.num_k_iters - 1) umma_commit_pair(&tmem_full_bar);
            }
          } else {
            const uint64_t a_desc = make_smem_desc_sw64(a_addr);
            const uint64_t b_desc = make_smem_desc_sw64(a_addr + kABytes);
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < 2; ++k) umma_f8_pair(tmem_base, a_desc + 2u * k, b_desc + 2u * k, kIdesc2, 1u);
              umma_commit_pair(&empty_bar[stage]);
              if (it == p.num_k_iters - 1) umma_commit_pair(&tmem_full_bar);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 4) {
    int stage = 0;
    uint32_t phase = 0;
    int chunk = 0, dx = 0, dy = 0;
    for (int it = 0; it < p.num_k_iters; ++it) {
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      if (elect_one_sync()) {
        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);   // both CTAs' bytes land here
        uint8_t* sA = smem + stage * kStageBytes;
        uint8_t* sB = sA + kABytes;
        if (p.a_mode == ATDN_MODE_PATCH) {
          const int cw = w0 * p.stride + dx - p.pad_w;
          const int chh = h0 * p.stride + dy - p.pad_h;
          if (chunk < p.chunks_a) tma_load_4d_pair(sA, &p.tmA, &full_bar[stage], chunk * kChunkK, cw, chh, batch);
          else tma_load_4d_pair(sA, &p.tmA2, &full_bar[stage], (chunk - p.chunks_a) * kChunkK, cw, chh, batch);
        } else if (p.a_tiled) {
          tma_load_4d_pair(sA, &p.tmA, &full_bar[stage], 0, 0, it, batch * p.a_row_blocks + (m0 >> 5));
        } else {
          tma_load_4d_pair(sA, &p.tmA, &full_bar[stage], it * kChunkK, m0, 0, p.a_shared ? 0 : batch);
        }
        if constexpr (EPI == ATDN_EPI_CORR) {
          tma_load_4d_pair(sB, &p.tmB, &full_bar[stage], it * kChunkK, bw0, bh0 + rank * 4, batch);
        } else {
          tma_load_4d_pair(sB, &p.tmB, &full_bar[stage], it * kChunkK, n0 + rank * (BN / 2), 0, p.b_batched ? batch : 0);
        }
      }
      __syncwarp();
      if (++chunk == chunks) { chunk = 0; if (++dx == p.taps_w) { dx = 0; ++dy; } }
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 5) {
    if (rank == 0) {
      int stage = 0, mchunk = 0;
      uint32_t phase = 0;
      const uint32_t smem_base_u32 = smem_u32(smem);
      for (int it = 0; it < p.num_k_iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        int rem;
        if (p.a_mode == ATDN_MODE_PATCH) {
          rem = mchunk < p.chunks_a ? p.c_a - mchunk * kChunkK : p.c_a2 - (mchunk - p.chunks_a) * kChunkK;
          if (++mchunk == chunks) mchunk = 0;
        } else {
          rem = p.c_a - it * kChunkK;
        }
        const int ksteps = rem >= kChunkK ? 4 : (rem + 15) >> 4;
        const uint32_t a_addr = smem_base_u32 + stage * kStageBytes;
        const uint64_t a_desc = make_smem_desc_sw128(a_addr);
        const uint64_t b_desc = make_smem_desc_sw128(a_addr + kABytes);
        const uint32_t acc0 = it > 0 ? 1u : 0u;
        if (elect_one_sync()) {
          if (ksteps == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_pair(tmem_base, a_desc + 2u * k, b_desc + 2u * k, kIdesc, k == 0 ? acc0 : 1u);
          } else {
            for (int k = 0; k < ksteps; ++k) umma_f16_pair(tmem_base, a_desc + 2u * k, b_desc + 2u * k, kIdesc, k == 0 ? acc0 : 1u);
          }
          umma_commit_pair(&empty_bar[stage]);
          if (it == p.num_k_iters - 1) umma_commit_pair(&tmem_full_bar);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    const int r = warp * 32 + lane;
    bool valid;
    long long pix;
    if (p.a_mode == ATDN_MODE_PATCH) {
      const int h = h0 + (r >> 4), w = w0 + (r & 15);
      valid = (h < p.out_h) && (w < p.out_w);
      pix = (static_cast<long long>(batch) * p.out_h + h) * p.out_w + w;
    } else {
      valid = (m0 + r) < p.m_rows;
      pix = static_cast<long long>(batch) * p.m_rows + m0 + r;
    }
    mbar_wait_group128(&tmem_full_bar, 0, threadIdx.x);
    tcgen05_fence_after();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    if constexpr (EPI == ATDN_EPI_CORR) {
      epilogue_corr(p, valid, pix, bh0, bw0, trow);
    } else {
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(trow + c * 32, v);
        if constexpr (kAccN == 2 * BN) {
          if (mixed) {             // + the lo-plane half of the accumulator
            uint32_t w[32];
            tmem_ld_32x32(trow + BN + c * 32, w);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
          }
        }
        tmem_ld_wait();
        epilogue_chunk<EPI>(p.e, valid, pix, n0 + c * 32, v);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();          // neither CTA frees TMEM / exits while the pair's MMAs or epilogues are in flight
  if (warp == 5) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES, int EPI>
static int launch(const TcParams& p, dim3 grid, cudaStream_t stream) {
  constexpr int smem = STAGES * (kABytes + BN * kChunkK * 2) + 1024;
  static DeviceOnce configured;
  if (configured.pending()) {
    ATDN_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.done();
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = env_switches().pdl ? 1 : 0;
  ATDN_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BN, STAGES, EPI>, p));
  return 0;
}

template <int BN, int STAGES, int EPI>
static int launch2(const TcParams& p, dim3 grid, cudaStream_t stream) {
  constexpr int smem = STAGES * (kABytes + (BN / 2) * kChunkK * 2) + 1024 + (EPI == ATDN_EPI_PV ? 8192 : 0);   // PV: + the zero tile
  static DeviceOnce configured;
  if (configured.pending()) {
    ATDN_CUDA(cudaFuncSetAttribute(tc_gemm2_kernel<BN, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.done();
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = env_switches().pdl ? 1 : 0;
  ATDN_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm2_kernel<BN, STAGES, EPI>, p));   // (the cluster shape is the kernel's __cluster_dims__)
  return 0;
}

template <int EPI>
static int dispatch_bn2(int bn, const TcParams& p, dim3 grid, cudaStream_t s) {
  switch (bn) {
    case 64:  return launch2<64, 4, EPI>(p, grid, s);
    case 96:  return launch2<96, 4, EPI>(p, grid, s);
    case 128: return launch2<128, 4, EPI>(p, grid, s);
    case 192: return launch2<192, 3, EPI>(p, grid, s);
    case 256: return launch2<256, 3, EPI>(p, grid, s);
    default:  return set_error(ATDN_ERR_UNSUP, "atdn_tc_gemm: unsupported bn %d for pair epilogue %d", bn, EPI);
  }
}

template <int EPI>
static int dispatch_bn(int bn, const TcParams& p, dim3 grid, cudaStream_t s) {
  switch (bn) {
    case 64:  return launch<64, 4, EPI>(p, grid, s);
    case 96:  return launch<96, 3, EPI>(p, grid, s);
    case 128: return launch<128, 3, EPI>(p, grid, s);
    case 192: return launch<192, 4, EPI>(p, grid, s);
    default:  return set_error(ATDN_ERR_UNSUP, "atdn_tc_gemm: unsupported bn %d for epilogue %d", bn, EPI);
  }
}

}  // namespace atdn

using namespace atdn;

extern "C" int atdn_tc_gemm(const atdn_tc_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ATDN_REQUIRE(d != nullptr, ATDN_ERR_ARG, "atdn_tc_gemm: null descriptor");
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(d->a_mode == ATDN_MODE_ROWS || d->a_mode == ATDN_MODE_PATCH, ATDN_ERR_ARG, "atdn_tc_gemm: bad a_mode");
  ATDN_REQUIRE(d->n_valid > 0, ATDN_ERR_ARG, "atdn_tc_gemm: n_valid must be positive");
  if (d->mt > 0) return launch_conv_halo(d, stream);
  const bool corr = d->epi == ATDN_EPI_CORR;
  ATDN_REQUIRE((d->b_mode == ATDN_MODE_PATCH) == corr, ATDN_ERR_ARG, "atdn_tc_gemm: PATCH B operand is only valid with ATDN_EPI_CORR");
  ATDN_REQUIRE(!(d->flags & ATDN_F_PAIR) || d->bn % 32 == 0, ATDN_ERR_ARG, "atdn_tc_gemm: pair kernel needs bn %% 32 == 0");
  ATDN_REQUIRE(!corr || (d->bn == 256 && d->a_mode == ATDN_MODE_ROWS), ATDN_ERR_ARG, "atdn_tc_gemm: CORR needs bn=256 and ROWS A");

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.a_mode = d->a_mode;
  p.e.n_valid = d->n_valid;
  p.e.flags = d->flags;
  p.b_batched = (d->flags & ATDN_F_B_BATCHED) ? 1 : 0;
  p.a_shared = (d->flags & ATDN_F_A_SHARED) ? 1 : 0;
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
  p.e.img_h = d->out_h > 0 ? d->out_h : 1;   // pixel decode of the tiled recurrent-state layout (tc_epilogue.cuh)
  p.e.img_w = d->out_w > 0 ? d->out_w : 1;
  if (d->out8) {
    ATDN_REQUIRE(d->epi == ATDN_EPI_STORE16 && d->a_mode == ATDN_MODE_ROWS && !(d->flags & ATDN_F_PAIR) && aligned16(d->out8) &&
                 d->out_pitch % 32 == 0 && d->out_ch_off == 0 && !d->bias, ATDN_ERR_ARG,
                 "atdn_tc_gemm: out8 needs a single-CTA ROWS STORE16 GEMM without bias, out_pitch %% 32 == 0 and out_ch_off == 0");
    p.e.out8 = static_cast<uint8_t*>(d->out8);
    p.e.out8_rows = (int)d->a_dims[1];
  }
  ATDN_REQUIRE(p.e.out != nullptr || d->epi == ATDN_EPI_GRU_ZR, ATDN_ERR_ARG, "atdn_tc_gemm: null output");

  const uint32_t ones[4] = {1, 1, 1, 1};
  dim3 grid;
  const bool pair = (d->flags & ATDN_F_PAIR) != 0;    // CTA-pair kernel: 256-row tiles, grid.x = 2 * pairs
  const int batch = (d->flags & ATDN_F_A_SHARED) ? (int)d->b_dims[3] : (int)d->a_dims[3];
  ATDN_REQUIRE(!(d->flags & ATDN_F_A_SHARED) || d->a_mode == ATDN_MODE_ROWS, ATDN_ERR_ARG, "atdn_tc_gemm: A_SHARED needs ROWS A");
  grid.z = batch;
  const int64_t c_a = d->a_dims[0];
  if (d->a_mode == ATDN_MODE_PATCH) {
    ATDN_REQUIRE(d->stride == 1 || d->stride == 2, ATDN_ERR_UNSUP, "atdn_tc_gemm: stride %d", d->stride);
    ATDN_REQUIRE(d->taps_h >= 1 && d->taps_w >= 1 && d->out_h >= 1 && d->out_w >= 1, ATDN_ERR_ARG, "atdn_tc_gemm: bad conv geometry");
    const uint32_t s = (uint32_t)d->stride;
    const uint32_t box[4] = {64, 16 * s, 8 * s, 1};
    const uint32_t es[4] = {1, s, s, 1};
    if (int e = make_map_f16(&p.tmA, d->a, d->a_dims, d->a_strides, box, es, "A")) return e;
    p.chunks_a = (d->a_split_chunk > 0) ? d->a_split_chunk : (int)((c_a + 63) / 64);
    p.c_a = (int)c_a;
    if (d->a_split_chunk > 0) {
      ATDN_REQUIRE(d->a2 != nullptr, ATDN_ERR_ARG, "atdn_tc_gemm: a_split_chunk without a2");
      ATDN_REQUIRE(c_a == 64LL * d->a_split_chunk, ATDN_ERR_ARG, "atdn_tc_gemm: a must hold exactly a_split_chunk*64 channels");
      if (int e = make_map_f16(&p.tmA2, d->a2, d->a2_dims, d->a2_strides, box, es, "A2")) return e;
      p.chunks_a2 = (int)((d->a2_dims[0] + 63) / 64);
      p.c_a2 = (int)d->a2_dims[0];
    }
    p.taps_w = d->taps_w;
    p.pad_h = d->pad_h;
    p.pad_w = d->pad_w;
    p.stride = d->stride;
    p.out_h = d->out_h;
    p.out_w = d->out_w;
    p.tiles_w = ceil_div(d->out_w, 16);
    p.num_k_iters = d->taps_h * d->taps_w * (p.chunks_a + p.chunks_a2);
    grid.x = pair ? 2 * p.tiles_w * ceil_div(d->out_h, 16) : p.tiles_w * ceil_div(d->out_h, 8);
  } else {
    if (d->flags & ATDN_F_A_TILED) {
      ATDN_REQUIRE(!(d->flags & ATDN_F_A_SHARED) && c_a % 64 == 0, ATDN_ERR_ARG,
                   "atdn_tc_gemm: A_TILED needs a batched A and a column count that is a multiple of 64");
      p.a_tiled = 1;
      p.a_row_blocks = ceil_div((int)d->a_dims[1], 32);
      const int64_t cb = c_a / 64;
      const int64_t dims[4] = {64, 32, cb, d->a_dims[3] * p.a_row_blocks};
      const int64_t str[3] = {64, 2048, cb * 2048};
      const uint32_t box[4] = {64, 32, 1, 4};            // 4 row blocks = the 128 rows of an M tile, 4 KiB contiguous each
      if (int e = make_map_f16(&p.tmA, d->a, dims, str, box, ones, "A (tiled)")) return e;
      if (d->flags & ATDN_F_A_MIXED) {
        ATDN_REQUIRE(pair && d->epi == ATDN_EPI_PV && (d->bn == 128 || d->bn == 64) && d->b8 && d->a_hot && (d->flags & ATDN_F_B_BATCHED) &&
                     d->b_strides[0] % 16 == 0, ATDN_ERR_ARG,
                     "atdn_tc_gemm: A_MIXED needs ATDN_F_PAIR, EPI_PV, bn 128 or 64, a batched B with its e4m3 planes (b8) and the pair bitmap (a_hot)");
        const int64_t str8[3] = {64, 4096, cb * 4096};     // the same slots seen as bytes: an e4m3 sub-block is the first 2 KiB
        if (int e = make_map(&p.tmA2, 1, CU_TENSOR_MAP_SWIZZLE_64B, d->a, dims, str8, box, ones, "A (tiled, e4m3 blocks)")) return e;
        const int64_t bdims[4] = {d->b_dims[0], d->b_dims[1], 2, d->b_dims[3]};
        const int64_t bstr[3] = {d->b_strides[0], d->b_dims[1] * d->b_strides[0], 2 * d->b_dims[1] * d->b_strides[0]};
        const uint32_t bbox[4] = {64, (uint32_t)d->bn, 1, 1};       // all rows of the tile, one plane per CTA of the pair
        if (int e = make_map(&p.tmB8, 1, CU_TENSOR_MAP_SWIZZLE_64B, d->b8, bdims, bstr, bbox, ones, "B (e4m3 planes)")) return e;
        p.a_hot = d->a_hot;
      }
    } else {
      const uint32_t box[4] = {64, 128, 1, 1};
      if (int e = make_map_f16(&p.tmA, d->a, d->a_dims, d->a_strides, box, ones, "A")) return e;
    }
    p.c_a = (int)c_a;
    p.chunks_a = (int)((c_a + 63) / 64);
    p.num_k_iters = p.chunks_a;
    p.m_rows = (int)d->a_dims[1];
    p.taps_w = 1;
    p.stride = 1;
    grid.x = pair ? 2 * ceil_div(p.m_rows, 256) : ceil_div(p.m_rows, 128);
  }
  // the packed K extent of B must cover every K iteration
  ATDN_REQUIRE(d->b_dims[0] >= (int64_t)(p.num_k_iters - 1) * 64 + 1, ATDN_ERR_ARG,
               "atdn_tc_gemm: B has K extent %lld but the A side iterates %d chunks of 64", (long long)d->b_dims[0], p.num_k_iters);

  if (corr) {
    const uint32_t box[4] = {64, 32, pair ? 4u : 8u, 1};
    if (int e = make_map_f16(&p.tmB, d->b, d->b_dims, d->b_strides, box, ones, "B")) return e;
    p.corr_h = d->corr_h;
    p.corr_w = d->corr_w;
    ATDN_REQUIRE(d->corr_h == d->b_dims[2] && d->corr_w == d->b_dims[1], ATDN_ERR_ARG, "atdn_tc_gemm: corr grid != B image dims");
    p.corr_tiles_w = ceil_div(d->corr_w, 32);
    for (int i = 0; i < 4; ++i) {
      p.lvl_pitch[i] = d->lvl_pitch[i];
      ATDN_REQUIRE(d->lvl_pitch[i] % 4 == 0 && d->lvl_pitch[i] >= (d->corr_w >> i), ATDN_ERR_ALIGN, "atdn_tc_gemm: lvl_pitch[%d]", i);
    }
    for (int i = 0; i < 3; ++i) {
      p.lvl[i] = d->lvl[i];
      ATDN_REQUIRE(d->lvl[i] != nullptr && aligned16(d->lvl[i]), ATDN_ERR_ALIGN, "atdn_tc_gemm: lvl[%d]", i);
    }
    ATDN_REQUIRE(aligned16(d->out), ATDN_ERR_ALIGN, "atdn_tc_gemm: out");
    grid.y = p.corr_tiles_w * ceil_div(d->corr_h, 8);
    return pair ? launch2<256, 3, ATDN_EPI_CORR>(p, grid, stream) : launch<256, 4, ATDN_EPI_CORR>(p, grid, stream);
  }

  {
    const uint32_t box[4] = {64, (uint32_t)(pair ? d->bn / 2 : d->bn), 1, 1};
    if (int e = make_map_f16(&p.tmB, d->b, d->b_dims, d->b_strides, box, ones, "B")) return e;
  }
  grid.y = ceil_div(d->n_valid, d->bn);
  {
    p.out_tma = d->epi == ATDN_EPI_STORE16 && !pair && d->out != nullptr && !env_switches().no_out_tma && aligned16(d->out) &&
                d->out_pitch % 8 == 0 && d->out_ch_off % 8 == 0;
    if (p.out_tma) {
      const bool patch = d->a_mode == ATDN_MODE_PATCH;
      const int64_t nb = (d->flags & ATDN_F_A_SHARED) ? d->b_dims[3] : d->a_dims[3];
      const int64_t rows = patch ? 0 : d->a_dims[1];
      const int64_t odims[4] = {d->n_valid, patch ? d->out_w : rows, patch ? d->out_h : 1, nb};
      const int64_t ostr[3] = {d->out_pitch, patch ? (int64_t)d->out_w * d->out_pitch : rows * d->out_pitch,
                               patch ? (int64_t)d->out_h * d->out_w * d->out_pitch : rows * d->out_pitch};
      const uint32_t obox[4] = {32, patch ? 16u : 32u, patch ? 2u : 1u, 1};
      const __half* obase = static_cast<const __half*>(d->out) + d->out_ch_off;
      if (int e = make_map(&p.tmOut, 2, CU_TENSOR_MAP_SWIZZLE_64B, obase, odims, ostr, obox, ones, "GEMM output")) return e;
    }
  }
  // vector-store alignment of the epilogue
  ATDN_REQUIRE(d->out_pitch % 8 == 0 && d->out_ch_off % 8 == 0, ATDN_ERR_ALIGN, "atdn_tc_gemm: out_pitch/out_ch_off must be multiples of 8");
  switch (d->epi) {
    case ATDN_EPI_STORE16:
      ATDN_REQUIRE(!(d->flags & ATDN_F_RESID) || (d->resid16 && d->resid_pitch % 8 == 0 && d->resid_ch_off % 8 == 0), ATDN_ERR_ARG, "atdn_tc_gemm: residual");
      ATDN_REQUIRE(!(d->flags & ATDN_F_FLOWTAIL) || d->aux32, ATDN_ERR_ARG, "atdn_tc_gemm: FLOWTAIL needs aux32");
      ATDN_REQUIRE(!(d->flags & ATDN_F_TANH_LO) || d->h32, ATDN_ERR_ARG, "atdn_tc_gemm: TANH_LO needs h32");
      return pair ? dispatch_bn2<ATDN_EPI_STORE16>(d->bn, p, grid, stream) : dispatch_bn<ATDN_EPI_STORE16>(d->bn, p, grid, stream);
    case ATDN_EPI_STORE32:
      return pair ? dispatch_bn2<ATDN_EPI_STORE32>(d->bn, p, grid, stream) : dispatch_bn<ATDN_EPI_STORE32>(d->bn, p, grid, stream);
    case ATDN_EPI_GRU_ZR:
      ATDN_REQUIRE(d->n_valid == 256 && d->h32 && d->z32 && d->rh16, ATDN_ERR_ARG, "atdn_tc_gemm: GRU_ZR arguments");
      if (pair) {
        ATDN_REQUIRE(d->bn == 256 || d->bn == 128, ATDN_ERR_ARG, "atdn_tc_gemm: GRU_ZR pair kernel needs bn 128 or 256");
        return d->bn == 256 ? launch2<256, 3, ATDN_EPI_GRU_ZR>(p, grid, stream) : launch2<128, 4, ATDN_EPI_GRU_ZR>(p, grid, stream);
      }
      ATDN_REQUIRE(d->bn == 128, ATDN_ERR_ARG, "atdn_tc_gemm: GRU_ZR needs bn 128");
      return launch<128, 3, ATDN_EPI_GRU_ZR>(p, grid, stream);
    case ATDN_EPI_GRU_Q:
      ATDN_REQUIRE(d->n_valid == 128 && d->h32 && d->z32, ATDN_ERR_ARG, "atdn_tc_gemm: GRU_Q arguments");
      if (pair) {
        ATDN_REQUIRE(d->bn == 128 || d->bn == 64, ATDN_ERR_ARG, "atdn_tc_gemm: GRU_Q pair kernel needs bn 64 or 128");
        return d->bn == 128 ? launch2<128, 4, ATDN_EPI_GRU_Q>(p, grid, stream) : launch2<64, 4, ATDN_EPI_GRU_Q>(p, grid, stream);
      }
      return dispatch_bn<ATDN_EPI_GRU_Q>(d->bn, p, grid, stream);
    case ATDN_EPI_PV:
      ATDN_REQUIRE(d->resid16 && d->aux32 && d->gamma && d->resid_pitch % 8 == 0 && d->resid_ch_off % 8 == 0, ATDN_ERR_ARG, "atdn_tc_gemm: PV arguments");
      if (pair) {
        ATDN_REQUIRE(d->bn == 128 || d->bn == 64, ATDN_ERR_ARG, "atdn_tc_gemm: PV pair kernel needs bn 64 or 128");
        return d->bn == 128 ? launch2<128, 4, ATDN_EPI_PV>(p, grid, stream) : launch2<64, 4, ATDN_EPI_PV>(p, grid, stream);
      }
      return dispatch_bn<ATDN_EPI_PV>(d->bn, p, grid, stream);
    default:
      return set_error(ATDN_ERR_UNSUP, "atdn_tc_gemm: unknown epilogue %d", d->epi);
  }
}
