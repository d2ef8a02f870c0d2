// Fused attention probabilities for sm_100a: P = exp(q k^T * scale - rowmax) (fp16, un-normalised) and
// 1 / rowsum, without ever writing the fp32 logits to HBM (gma.py:72-73 of the reference's GMA wheel).
//
// The unfused path (QK^T GEMM -> fp32 S in HBM -> row softmax kernel) moves 4 + 4 + 2 bytes per logit and ran
// at 8% of the tensor peak; the GEMM itself is tiny (K = 128).  Here one CTA owns 128 query rows and streams
// the key tiles of its image TWICE through the tensor cores:
//   pass 1: S tile -> TMEM -> row maxima (registers only);
//   pass 2: S tile -> TMEM -> exp2((s - max) * scale * log2 e) -> fp16 -> swizzled shared memory -> TMA store.
// Recomputing S costs 2 x 13.9 GFLOP per pair (0.02 ms of tensor time) and removes 0.42 GB of HBM traffic per
// pair; the only HBM traffic left is the P write (2 bytes per logit) that the P.V GEMM needs anyway.
//
// Warp roles (384 threads): warp 0 = TMA producer (Q once, K tiles through a ring), warp 1 = TMEM allocator +
// MMA issuer, warps 4..11 = epilogue.  The 512 TMEM columns hold two 128 x 256 fp32 accumulators; epilogue group
// g (warps 4+4g .. 7+4g, warp w reads TMEM lanes 32 (w & 3) ..) drains accumulator g, i.e. every second key
// tile, while the MMA warp fills the other one.  Row maxima / sums of the two groups meet in shared memory.
#include <math.h>
#include <string.h>

#include "common.h"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace atdn {

constexpr int kAttnThreads = 384;
constexpr int kAttnBN = 256;                       // keys per tile
constexpr int kAttnKStages = 4;                    // ring of 256 x 64 fp16 K chunks (32 KiB each)
constexpr int kAttnQBytes = 2 * 128 * 128;         // 2 chunks of 128 rows x 64 fp16
constexpr int kAttnKStageBytes = kAttnBN * 128;
constexpr int kAttnStoreBytes = 32 * 128;          // one 32-row x 64-column fp16 box per warp and buffer
constexpr int kAttnSmem = kAttnQBytes + kAttnKStages * kAttnKStageBytes + 8 * 2 * kAttnStoreBytes + 1024;

struct alignas(64) AttnParams {
  CUtensorMap tmQ, tmK, tmP;
  int n, tiles;
  int p_tiled, row_blocks;    // P in blocks of 32 rows x 64 columns ([batch][row block][column block][32][64]); ceil(n / 32)
  float scale_log2;       // softmax scale * log2(e)
  float* inv_sum;         // [batch * n]
};

__global__ void __launch_bounds__(kAttnThreads, 1) attn_probs_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, k_full[kAttnKStages], k_empty[kAttnKStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float xbuf[2][128];                   // row maxima, later row sums, of the two epilogue groups

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + kAttnQBytes;
  uint8_t* smem_st = smem_k + kAttnKStages * kAttnKStageBytes;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int batch = blockIdx.y;
  const int T = p.tiles;

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < kAttnKStages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 2 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmP);
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===== producer =====
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(&q_full, kAttnQBytes);
      tma_load_4d(smem_q, &p.tmQ, &q_full, 0, m0, 0, batch);
      tma_load_4d(smem_q + 128 * 128, &p.tmQ, &q_full, 64, m0, 0, batch);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int gi = 0; gi < 2 * T; ++gi) {
      const int j = gi >= T ? gi - T : gi;
      for (int c = 0; c < 2; ++c) {
        mbar_wait(&k_empty[stage], phase ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&k_full[stage], kAttnKStageBytes);
          tma_load_4d(smem_k + stage * kAttnKStageBytes, &p.tmK, &k_full[stage], c * 64, j * kAttnBN, 0, batch);
        }
        __syncwarp();
        if (++stage == kAttnKStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t kIdesc = make_idesc_f16(128, kAttnBN);
    const uint32_t q_u32 = smem_u32(smem_q), k_u32 = smem_u32(smem_k);
    int stage = 0;
    uint32_t phase = 0, pe0 = 0, pe1 = 0;
    mbar_wait(&q_full, 0);
    for (int gi = 0; gi < 2 * T; ++gi) {
      const int buf = gi & 1;
      const uint32_t pe = buf ? pe1 : pe0;
      mbar_wait(&acc_empty[buf], pe ^ 1u);
      if (buf) pe1 ^= 1u; else pe0 ^= 1u;
      tcgen05_fence_after();
      const uint32_t d = tmem_base + static_cast<uint32_t>(buf * kAttnBN);
      for (int c = 0; c < 2; ++c) {
        mbar_wait(&k_full[stage], phase);
        tcgen05_fence_after();
        const uint64_t a_desc = make_smem_desc_sw128(q_u32 + c * 128 * 128);
        const uint64_t b_desc = make_smem_desc_sw128(k_u32 + stage * kAttnKStageBytes);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, a_desc + 2u * k, b_desc + 2u * k, kIdesc, (c | k) ? 1u : 0u);
          umma_commit(&k_empty[stage]);
          if (c == 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++stage == kAttnKStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp & 3, g = (warp - 4) >> 2;
    const int row_local = q * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g * kAttnBN);
    const uint32_t st_u32 = smem_u32(smem_st + (warp - 4) * 2 * kAttnStoreBytes);
    const uint8_t* st_ptr = smem_st + (warp - 4) * 2 * kAttnStoreBytes;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    float mx = -INFINITY, mxs = 0.0f, sum = 0.0f;
    uint32_t pf = 0;
    bool exchanged = false;
    int nstore = 0;

    auto exchange = [&]() {
      xbuf[g][row_local] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(xbuf[0][row_local], xbuf[1][row_local]);
      mxs = mx * p.scale_log2;
      asm volatile("bar.sync 1, 256;" ::: "memory");   // xbuf is reused for the row sums
      exchanged = true;
    };

    for (int gi = g; gi < 2 * T; gi += 2) {
      const bool second = gi >= T;
      if (second && !exchanged) exchange();
      const int col0 = (second ? gi - T : gi) * kAttnBN;
      mbar_wait(&acc_full[g], pf);
      pf ^= 1u;
      tcgen05_fence_after();
      if (!second) {
#pragma unroll 1
        for (int c = 0; c < kAttnBN / 32; ++c) {
          const int cb = col0 + c * 32;
          if (cb >= p.n) break;
          uint32_t v[32];
          tmem_ld_32x32(trow + c * 32, v);
          tmem_ld_wait();
          if (cb + 32 <= p.n) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cb + j < p.n) mx = fmaxf(mx, __uint_as_float(v[j]));
          }
        }
      } else {
#pragma unroll 1
        for (int cg = 0; cg < kAttnBN / 64; ++cg) {
          const int cb = col0 + cg * 64;
          if (cb >= p.n) break;
          const int b = nstore & 1;
          if (lane == 0) bulk_wait_read<1>();       // the store that last read this buffer has drained it
          __syncwarp();
          uint32_t v[64];
          tmem_ld_32x32(trow + cg * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld_32x32(trow + cg * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
          const bool full = cb + 64 <= p.n;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {          // 16-byte chunk = 8 probabilities
            uint32_t w[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int j = ch * 8 + h * 2;
              float e0 = ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2, -mxs));
              float e1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, -mxs));
              if (!full) {
                if (cb + j >= p.n) e0 = 0.0f;
                if (cb + j + 1 >= p.n) e1 = 0.0f;
              }
              const __half2 hh = __floats2half2_rn(e0, e1);
              const float2 r = __half22float2(hh);   // normalise by what the P.V GEMM will actually read
              sum += r.x + r.y;
              w[h] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            st_shared_v4(st_u32 + b * kAttnStoreBytes + lane * 128 + ((static_cast<uint32_t>(ch) ^ sw) << 4),
                         make_uint4(w[0], w[1], w[2], w[3]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (!p.p_tiled) {
              tma_store_4d(&p.tmP, st_ptr + b * kAttnStoreBytes, cb, m0 + q * 32, 0, batch);
            } else if ((m0 >> 5) + q < p.row_blocks) {   // (the last CTA overhangs the last row block of the batch element)
              tma_store_4d(&p.tmP, st_ptr + b * kAttnStoreBytes, 0, 0, cb >> 6, batch * p.row_blocks + (m0 >> 5) + q);
            }
            bulk_commit();
          }
          ++nstore;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[g]);
    }
    if (!exchanged) exchange();
    xbuf[g][row_local] = sum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (g == 0 && m0 + row_local < p.n)
      p.inv_sum[static_cast<long long>(batch) * p.n + m0 + row_local] = 1.0f / (xbuf[0][row_local] + xbuf[1][row_local]);
    if (lane == 0) bulk_wait_read<0>();             // shared memory stays valid until the last store has read it
    __syncwarp();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace atdn

using namespace atdn;

extern "C" int atdn_attn_probs(const void* qk16, int64_t qk_pitch, void* p16, int64_t p_pitch, int32_t p_tiled, float* inv_sum,
                               int32_t batch, int32_t n, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(qk16 && p16 && inv_sum && batch > 0 && n > 0, ATDN_ERR_ARG, "atdn_attn_probs: null / empty argument");
  ATDN_REQUIRE(qk_pitch >= 256 && qk_pitch % 8 == 0 && p_pitch >= n && p_pitch % 8 == 0, ATDN_ERR_ALIGN,
               "atdn_attn_probs: qk_pitch %lld / p_pitch %lld", (long long)qk_pitch, (long long)p_pitch);
  ATDN_REQUIRE(scale > 0.0f, ATDN_ERR_ARG, "atdn_attn_probs: scale must be positive (row maxima are taken before scaling)");
  ATDN_REQUIRE(!p_tiled || p_pitch % 64 == 0, ATDN_ERR_ALIGN, "atdn_attn_probs: the tiled layout needs p_pitch %% 64 == 0, got %lld", (long long)p_pitch);
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.n = n;
  p.tiles = ceil_div(n, kAttnBN);
  p.scale_log2 = scale * 1.4426950408889634f;
  p.inv_sum = inv_sum;
  const uint32_t ones[4] = {1, 1, 1, 1};
  const int64_t qdims[4] = {128, n, 1, batch};
  const int64_t qstr[3] = {qk_pitch, (int64_t)n * qk_pitch, (int64_t)n * qk_pitch};
  const uint32_t qbox[4] = {64, 128, 1, 1}, kbox[4] = {64, kAttnBN, 1, 1}, pbox[4] = {64, 32, 1, 1};
  if (int e = make_map_f16(&p.tmQ, qk16, qdims, qstr, qbox, ones, "Q")) return e;
  if (int e = make_map_f16(&p.tmK, static_cast<const __half*>(qk16) + 128, qdims, qstr, kbox, ones, "K")) return e;
  p.p_tiled = p_tiled ? 1 : 0;
  p.row_blocks = ceil_div(n, 32);
  if (p_tiled) {
    const int64_t cb = p_pitch / 64;
    const int64_t pdims[4] = {64, 32, cb, (int64_t)batch * p.row_blocks};
    const int64_t pstr[3] = {64, 2048, cb * 2048};
    if (int e = make_map_f16(&p.tmP, p16, pdims, pstr, pbox, ones, "P (tiled)")) return e;
  } else {
    const int64_t pdims[4] = {n, n, 1, batch};
    const int64_t pstr[3] = {p_pitch, (int64_t)n * p_pitch, (int64_t)n * p_pitch};
    if (int e = make_map_f16(&p.tmP, p16, pdims, pstr, pbox, ones, "P")) return e;
  }
  static DeviceOnce configured;
  if (configured.pending()) {
    ATDN_CUDA(cudaFuncSetAttribute(attn_probs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    configured.done();
  }
  attn_probs_kernel<<<dim3(ceil_div(n, 128), batch), kAttnThreads, kAttnSmem, stream>>>(p);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}
