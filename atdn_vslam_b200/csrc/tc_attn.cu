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
// Warp roles (576 threads): warp 0 = TMA producer (Q once, K tiles through a ring), warp 1 = TMEM allocator +
// MMA issuer, warps 2..17 = epilogue.  The 512 TMEM columns hold two 128 x 256 fp32 accumulators; four epilogue
// groups of four warps (warp w reads TMEM lanes 32 (w & 3) ..): groups 0 / 1 drain the two 128-column halves of
// accumulator 0, groups 2 / 3 those of accumulator 1, i.e. every second key tile, while the MMA warp fills the
// other one.  (Eight epilogue warps left the issue slots 57% used: the epilogue is a chain of TMEM load -> exp ->
// pack -> staging store per chunk, and two warps per scheduler cannot hide it.)  Row maxima / sums of the four
// groups meet in shared memory.
//
// MIXED storage (block_hot != NULL): P.V re-reads P twelve times per pair at the HBM roofline, so every 32-row x
// 64-column sub-block whose rounding error cannot matter is stored as e4m3 (64-byte rows, half the bytes) and read back
// by a kind::f8f6f4 MMA; the sub-blocks that carry a row's mass stay fp16.  All values are stored scaled by 256 (e4m3 then keeps
// its 3 mantissa bits down to 2^-14 of the row maximum; the factor cancels against the row sum).  A sub-block is "hot"
// (fp16) when for one of its rows  sqrt(sum_block p^2) > hot_energy * sum_row p  -- the e4m3 rounding error of the block
// relative to that row's output (tools/fp8_attention_sensitivity.py, variants mixed-e*).  The row sums this needs are
// estimated in pass 1 with an online softmax; pass 2 still sums the ROUNDED values for the normalisation.
#include <math.h>
#include <string.h>

#include "common.h"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace atdn {

constexpr int kAttnEpiWarps = 16;                  // 4 groups of 4 warps: (accumulator buffer 0 / 1) x (column half 0 / 1)
constexpr int kAttnThreads = (2 + kAttnEpiWarps) * 32;   // warp 0 producer, warp 1 TMEM allocator + MMA issuer
constexpr int kAttnBN = 256;                       // keys per tile
constexpr int kAttnKStages = 3;                    // ring of 256 x 64 fp16 K chunks (32 KiB each); 16 staging boxes take the rest
constexpr int kAttnQBytes = 2 * 128 * 128;         // 2 chunks of 128 rows x 64 fp16
constexpr int kAttnKStageBytes = kAttnBN * 128;
constexpr int kAttnStoreBytes = 32 * 128;          // one 32-row x 64-column fp16 box per epilogue warp
constexpr int kAttnSmem = kAttnQBytes + kAttnKStages * kAttnKStageBytes + kAttnEpiWarps * kAttnStoreBytes + 1024;

struct alignas(64) AttnParams {
  CUtensorMap tmQ, tmK, tmP, tmP8;
  int n, tiles;
  int p_tiled, row_blocks;    // P in blocks of 32 rows x 64 columns ([batch][row block][column block][32][64]); ceil(n / 32)
  float scale_log2;       // softmax scale * log2(e)
  float* inv_sum;         // [batch * n]
  uint8_t* block_hot;     // MIXED: [batch][row_blocks][col_blocks] 1 = fp16 sub-block (32 rows x 64 columns), 0 = e4m3
  int col_blocks;
  float hot_thr;          // (hot_energy * 256)^2
};

constexpr float kAttnStoreLog2 = 8.0f;   // MIXED: stored value = 256 * exp(s - rowmax)

template <bool MIXED>
__global__ void __launch_bounds__(kAttnThreads, 1) attn_probs_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, k_full[kAttnKStages], k_empty[kAttnKStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float xbuf[4][128];                   // row maxima, later row sums, of the four epilogue groups

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
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    fence_barrier_init();
  }
  if (warp == 3 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmP);
    if constexpr (MIXED) tma_prefetch_desc(&p.tmP8);
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
  } else {
    // ===== epilogue: warp e = warp - 2 reads TMEM lanes 32 (warp & 3) .. (any four consecutive warps cover the 128 rows);
    // group e >> 2 drains accumulator buffer g = group >> 1 (every second key tile) and, of its 256 columns, the half hc = group & 1
    const int e = warp - 2;
    const int q = warp & 3, g = e >> 3, hc = (e >> 2) & 1, grp = e >> 2;
    const int row_local = q * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g * kAttnBN + hc * (kAttnBN / 2));
    const uint32_t st_u32 = smem_u32(smem_st + e * kAttnStoreBytes);
    const uint8_t* st_ptr = smem_st + e * kAttnStoreBytes;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    float mx = -INFINITY, mxs = 0.0f, sum = 0.0f;
    float s1 = 0.0f, thr = 0.0f;                     // MIXED: pass-1 row sum (relative to mx), hot threshold on sum_block p^2
    uint32_t pf = 0;
    bool exchanged = false;

    auto exchange = [&]() {
      xbuf[grp][row_local] = mx;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      const float mv[4] = {xbuf[0][row_local], xbuf[1][row_local], xbuf[2][row_local], xbuf[3][row_local]};
      mx = fmaxf(fmaxf(mv[0], mv[1]), fmaxf(mv[2], mv[3]));
      mxs = mx * p.scale_log2;
      if constexpr (MIXED) {   // second round through the same buffer: the pass-1 row sums, each relative to its group's maximum
        asm volatile("bar.sync 1, 512;" ::: "memory");
        xbuf[grp][row_local] = s1;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        float srow = 0.0f;     // a group that saw no column holds (-inf, 0): ex2(-inf) = 0
#pragma unroll
        for (int i = 0; i < 4; ++i) srow += xbuf[i][row_local] * ex2_approx((mv[i] - mx) * p.scale_log2);
        const float sfl = fmaxf(srow, 1.0f);          // the row maximum itself contributes 1: a floor under the sampled estimate
        thr = p.hot_thr * sfl * sfl;
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");   // xbuf is reused for the row sums
      exchanged = true;
    };

    for (int gi = g; gi < 2 * T; gi += 2) {
      const bool second = gi >= T;
      if (second && !exchanged) exchange();
      const int col0 = (second ? gi - T : gi) * kAttnBN + hc * (kAttnBN / 2);
      mbar_wait(&acc_full[g], pf);
      pf ^= 1u;
      tcgen05_fence_after();
      if (!second) {
#pragma unroll 1
        for (int c = 0; c < kAttnBN / 64; ++c) {
          const int cb = col0 + c * 32;
          if (cb >= p.n) break;
          uint32_t v[32];
          tmem_ld_32x32(trow + c * 32, v);
          tmem_ld_wait();
          if constexpr (!MIXED) {
            if (cb + 32 <= p.n) {
#pragma unroll
              for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cb + j < p.n) mx = fmaxf(mx, __uint_as_float(v[j]));
            }
          } else {   // online softmax: running maximum and the row sum relative to it
            const bool full = cb + 32 <= p.n;
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full || cb + j < p.n) cm = fmaxf(cm, __uint_as_float(v[j]));
            if (cm > mx) {
              s1 *= ex2_approx((mx - cm) * p.scale_log2);
              mx = cm;
              mxs = mx * p.scale_log2;
            }
            // The sum only feeds the hot / cold criterion: every 4th column stands for its group of four (exact for the flat
            // rows where the criterion matters; a peaked row is over-estimated by at most 4x, i.e. still decided at the
            // 4 * hot_energy level, tools/fp8_attention_sensitivity.py mixed-e30).  25% of the MUFU work of a full pass.
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float e0 = ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2, -mxs));
              float e1 = ex2_approx(fmaf(__uint_as_float(v[j + 4]), p.scale_log2, -mxs));
              if (!full) {
                if (cb + j >= p.n) e0 = 0.0f;
                if (cb + j + 4 >= p.n) e1 = 0.0f;
              }
              a0 += e0; a1 += e1;
            }
            s1 += 4.0f * (a0 + a1);
          }
        }
      } else {
#pragma unroll 1
        for (int cg = 0; cg < kAttnBN / 128; ++cg) {
          const int cb = col0 + cg * 64;
          if (cb >= p.n) break;
          constexpr int b = 0;                      // one staging box per warp (16 warps x 4 KiB)
          if (lane == 0) bulk_wait_read<0>();       // the previous store of this warp has drained the box
          __syncwarp();
          uint32_t v[64];
          tmem_ld_32x32(trow + cg * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld_32x32(trow + cg * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
          const bool full = cb + 64 <= p.n;
          bool hot = true;
          if constexpr (MIXED) {
            // exp once, in place; the sub-block's energy decides its storage format
            const float off = kAttnStoreLog2 - mxs;
            float en0 = 0.0f, en1 = 0.0f;
#pragma unroll
            for (int j = 0; j < 64; j += 2) {
              float e0 = ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2, off));
              float e1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, off));
              if (!full) {
                if (cb + j >= p.n) e0 = 0.0f;
                if (cb + j + 1 >= p.n) e1 = 0.0f;
              }
              en0 = fmaf(e0, e0, en0);
              en1 = fmaf(e1, e1, en1);
              v[j] = __float_as_uint(e0);
              v[j + 1] = __float_as_uint(e1);
            }
            // one decision per warp = per 32-row sub-block (no cross-warp traffic); atdn_attn_harmonize makes the eight
            // sub-blocks of a 256-row P.V tile agree afterwards
            hot = __any_sync(0xffffffffu, (m0 + row_local < p.n) && (en0 + en1 > thr));
            if (lane == 0 && (m0 >> 5) + q < p.row_blocks)
              p.block_hot[(static_cast<long long>(batch) * p.row_blocks + (m0 >> 5) + q) * p.col_blocks + (cb >> 6)] = hot ? 1 : 0;
          }
          if (!MIXED || hot) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {          // 16-byte chunk = 8 probabilities
              uint32_t w[4];
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                const int j = ch * 8 + h * 2;
                float e0, e1;
                if constexpr (MIXED) {
                  e0 = __uint_as_float(v[j]);
                  e1 = __uint_as_float(v[j + 1]);
                } else {
                  e0 = ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2, -mxs));
                  e1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, -mxs));
                  if (!full) {
                    if (cb + j >= p.n) e0 = 0.0f;
                    if (cb + j + 1 >= p.n) e1 = 0.0f;
                  }
                }
                const __half2 hh = __floats2half2_rn(e0, e1);
                const float2 r = __half22float2(hh);   // normalise by what the P.V GEMM will actually read
                sum += r.x + r.y;
                w[h] = *reinterpret_cast<const uint32_t*>(&hh);
              }
              st_shared_v4(st_u32 + b * kAttnStoreBytes + lane * 128 + ((static_cast<uint32_t>(ch) ^ sw) << 4),
                           make_uint4(w[0], w[1], w[2], w[3]));
            }
          } else {
            const uint32_t sw8 = static_cast<uint32_t>(lane >> 1) & 3u;   // 64-byte rows, SWIZZLE_64B
            // The row sum of the ROUNDED values: the e4m3 values are unpacked to fp16 (exact) and added pairwise in fp16 -- exact
            // while the partial sums need <= 11 significant bits, i.e. for neighbours within 2^6 of each other, ~2^-12 relative
            // per level otherwise -- then once in fp32 per 16 values (a third of the instructions of 64 fp32 conversions + adds).
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {          // 16-byte chunk = 16 probabilities
              uint32_t w[4];
              __half2 part[4];
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                const int j = ch * 16 + h * 4;
                const uint32_t lo = pack2_e4m3(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                const uint32_t hi = pack2_e4m3(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                part[h] = __hadd2(unpack2_e4m3_h2(lo), unpack2_e4m3_h2(hi));
                w[h] = lo | (hi << 16);
              }
              const float2 r = __half22float2(__hadd2(__hadd2(part[0], part[1]), __hadd2(part[2], part[3])));
              sum += r.x + r.y;
              st_shared_v4(st_u32 + b * kAttnStoreBytes + lane * 64 + ((static_cast<uint32_t>(ch) ^ sw8) << 4),
                           make_uint4(w[0], w[1], w[2], w[3]));
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (!p.p_tiled) {
              tma_store_4d(&p.tmP, st_ptr + b * kAttnStoreBytes, cb, m0 + q * 32, 0, batch);
            } else if ((m0 >> 5) + q < p.row_blocks) {   // (the last CTA overhangs the last row block of the batch element)
              if (!MIXED || hot) tma_store_4d(&p.tmP, st_ptr + b * kAttnStoreBytes, 0, 0, cb >> 6, batch * p.row_blocks + (m0 >> 5) + q);
              else tma_store_4d(&p.tmP8, st_ptr + b * kAttnStoreBytes, 0, 0, cb >> 6, batch * p.row_blocks + (m0 >> 5) + q);
            }
            bulk_commit();
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[g]);
    }
    if (!exchanged) exchange();
    xbuf[grp][row_local] = sum;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (grp == 0 && m0 + row_local < p.n)
      p.inv_sum[static_cast<long long>(batch) * p.n + m0 + row_local] =
          1.0f / ((xbuf[0][row_local] + xbuf[1][row_local]) + (xbuf[2][row_local] + xbuf[3][row_local]));
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

// P.V on CTA pairs (tcgen05 cta_group::2) multiplies 256 rows per MMA, so the eight 32-row sub-blocks of a pair tile must
// hold a column block in the same format.  Where they disagree, the e4m3 sub-blocks are rewritten as fp16 IN PLACE (the
// same values: their rounding already happened and was judged harmless) and the pair bitmap says "fp16".  One warp per
// (pair tile, column block); a sub-block is [32][64] bytes at the start of its 4 KiB slot and becomes [32][64] halfs.
__global__ void __launch_bounds__(256) attn_harmonize_kernel(uint8_t* pbuf, const uint8_t* hot, uint8_t* pair_hot, int pairs, int cb,
                                                             int row_blocks) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + warp, pr = blockIdx.y, b = blockIdx.z;
  if (c >= cb) return;
  const int rb0 = pr * 8;
  uint32_t f = 0;
  if (lane < 8 && rb0 + lane < row_blocks) f = hot[(static_cast<long long>(b) * row_blocks + rb0 + lane) * cb + c];
  const uint32_t hot_mask = __ballot_sync(0xffffffffu, f != 0);
  if (lane == 0) pair_hot[(static_cast<long long>(b) * pairs + pr) * cb + c] = hot_mask ? 1 : 0;
  if (hot_mask == 0) return;
  for (int sub = 0; sub < 8; ++sub) {
    const int rbk = rb0 + sub;
    if (rbk >= row_blocks) break;
    if ((hot_mask >> sub) & 1u) continue;
    uint8_t* slot = pbuf + ((static_cast<long long>(b) * row_blocks + rbk) * cb + c) * 4096;
    uint4 in[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) in[i] = reinterpret_cast<const uint4*>(slot + lane * 64)[i];
    __syncwarp();                                  // every lane holds its row before any lane overwrites the slot
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t w[4] = {in[i].x, in[i].y, in[i].z, in[i].w};
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(o[2 * j]) : "h"(static_cast<uint16_t>(w[j] & 0xffffu)));
        asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(o[2 * j + 1]) : "h"(static_cast<uint16_t>(w[j] >> 16)));
      }
      uint4* dst = reinterpret_cast<uint4*>(slot + lane * 128 + i * 32);
      dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
      dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
    __syncwarp();
  }
}

}  // namespace atdn

using namespace atdn;

extern "C" int atdn_attn_harmonize(void* p16, int64_t p_pitch, const uint8_t* block_hot, uint8_t* pair_hot, int32_t batch, int32_t n,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(p16 && block_hot && pair_hot && batch > 0 && n > 0 && p_pitch == (int64_t)ceil_div(n, 64) * 64, ATDN_ERR_ARG,
               "atdn_attn_harmonize: null / empty argument or p_pitch != ceil64(n)");
  const int pairs = ceil_div(n, 256), cb = (int)(p_pitch / 64);
  attn_harmonize_kernel<<<dim3(ceil_div(cb, 8), pairs, batch), 256, 0, stream>>>(static_cast<uint8_t*>(p16), block_hot, pair_hot, pairs, cb,
                                                                               ceil_div(n, 32));
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_attn_probs(const void* qk16, int64_t qk_pitch, void* p16, int64_t p_pitch, int32_t p_tiled, float* inv_sum,
                               int32_t batch, int32_t n, float scale, uint8_t* block_hot, float hot_energy, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(qk16 && p16 && inv_sum && batch > 0 && n > 0, ATDN_ERR_ARG, "atdn_attn_probs: null / empty argument");
  ATDN_REQUIRE(qk_pitch >= 256 && qk_pitch % 8 == 0 && p_pitch >= n && p_pitch % 8 == 0, ATDN_ERR_ALIGN,
               "atdn_attn_probs: qk_pitch %lld / p_pitch %lld", (long long)qk_pitch, (long long)p_pitch);
  ATDN_REQUIRE(scale > 0.0f, ATDN_ERR_ARG, "atdn_attn_probs: scale must be positive (row maxima are taken before scaling)");
  ATDN_REQUIRE(!p_tiled || p_pitch % 64 == 0, ATDN_ERR_ALIGN, "atdn_attn_probs: the tiled layout needs p_pitch %% 64 == 0, got %lld", (long long)p_pitch);
  ATDN_REQUIRE(!block_hot || (p_tiled && p_pitch == (int64_t)ceil_div(n, 64) * 64 && hot_energy > 0.0f), ATDN_ERR_ARG,
               "atdn_attn_probs: mixed storage needs the tiled layout with p_pitch = ceil64(n) and hot_energy > 0");
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
    if (block_hot) {   // the same 4 KiB slots seen as bytes: an e4m3 block fills the first 2 KiB, 64-byte rows
      const int64_t dims8[4] = {64, 32, cb, (int64_t)batch * p.row_blocks};
      const int64_t str8[3] = {64, 4096, cb * 4096};
      if (int e = make_map(&p.tmP8, 1, CU_TENSOR_MAP_SWIZZLE_64B, p16, dims8, str8, pbox, ones, "P (tiled, e4m3 blocks)")) return e;
      p.block_hot = block_hot;
      p.col_blocks = (int)cb;
      p.hot_thr = (hot_energy * 256.0f) * (hot_energy * 256.0f);
    }
  } else {
    const int64_t pdims[4] = {n, n, 1, batch};
    const int64_t pstr[3] = {p_pitch, (int64_t)n * p_pitch, (int64_t)n * p_pitch};
    if (int e = make_map_f16(&p.tmP, p16, pdims, pstr, pbox, ones, "P")) return e;
  }
  static DeviceOnce configured;
  if (configured.pending()) {
    ATDN_CUDA(cudaFuncSetAttribute(attn_probs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    ATDN_CUDA(cudaFuncSetAttribute(attn_probs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    configured.done();
  }
  if (block_hot) attn_probs_kernel<true><<<dim3(ceil_div(n, 128), batch), kAttnThreads, kAttnSmem, stream>>>(p);
  else attn_probs_kernel<false><<<dim3(ceil_div(n, 128), batch), kAttnThreads, kAttnSmem, stream>>>(p);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}
