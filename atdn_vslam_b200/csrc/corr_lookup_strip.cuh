// Correlation lookup on the fp16 STRIP-layout pyramid (atdn_corr_pyramid half_levels = 4, include/atdn_b200.h).
// GMA.whl!/GMA/core/corr.py:32-53 + utils/utils.py:59-73 (bilinear grid_sample, align_corners=True, zero padding).
//
// What round 1's kernel (one warp per query, 8-byte loads into registers, fp32 windows in shared memory) measured
// on 27 pairs (profiles/r02a_ncu_corr_pyramid_lookup_b27_*): issue slots 85% busy at ~600 warp instructions per query,
// DRAM 4.3 TB/s of which only a third is algorithmic -- the L2 fetches 64-byte granules from HBM (9.1 M sectors
// delivered to the L1, 14.8 M read from DRAM, unchanged under ld.global.cg / L1::no_allocate /
// cudaLimitMaxL2FetchGranularity = 32), and a 64-byte granule of the tile-row layout was 1 x 32 texels: 14 granules per
// 10 x 10 window.  This kernel attacks both:
//   * layout: level l = [query][tile][strip][rows][8 cols] (see the header) -- a granule is a 4 x 8 texel block at
//     levels 0 and 1 (6.9 granules per window), rows are 16-byte pieces;
//   * staging: cp.async 16-byte pieces straight into shared memory (L1 bypass, zero fill for pieces outside the padded
//     maps, no conversion / masking pass: texels between the map edge and the tile edge are stored as zeros by the
//     pyramid kernel), double-buffered: a warp walks 8 consecutive queries and the pieces of query i + 1 are in
//     flight while query i is blended -- 2 x 3 KiB per warp, 32 warps per SM;
//   * arithmetic: the eight window origins of a query (4 levels x 2 axes) are computed by eight lanes and broadcast by
//     shuffles instead of being recomputed by every lane in every pass; address = row part + strip * S_l.
// The blend itself is the separable form of the first kernel: all 81 taps of a level share one fractional offset, so
// the 9 x 9 samples are 10 horizontal + 9 vertical lerps per window column; lane = (level, column pair): one pass of 20 lanes.
#pragma once
#include <math.h>
#include <stdint.h>

#include <cuda_fp16.h>

#include "tc_ptx.cuh"

namespace atdn {
namespace lks {

constexpr int kWarps = 8;                       // warps per CTA
constexpr int kQueriesPerWarp = 8;              // consecutive queries walked by one warp
constexpr int kRows = 10, kCols = 24;           // staged window per level: rows iy .. iy + 9, cols x0 .. x0 + 23 (x0 = ix & ~7)
constexpr int kLevelBytes = kRows * kCols * 2 + 64;  // 480 + 64 = 136 words: the four levels of the blend pass (5 lanes each, <= 7 words wide) start 8 banks apart
constexpr int kWinBytes = 4 * kLevelBytes;      // 2176 per query
constexpr int kOutBytes = 672;                  // 324 fp16 results (648 B) rounded up to 16 bytes
constexpr int kWarpBytes = 2 * kWinBytes + kOutBytes;
constexpr int kSmemBytes = kWarps * kWarpBytes; // 40192: dynamic shared memory, 4 CTAs per SM (64 registers per thread)

struct Params {
  const __half* lvl[4];     // level l: [query][tiles_l][chunk_l], chunk = 256 / 64 / 16 / 4 halves (strip layout)
  long long qstride[4];     // halves per query at level l
  int rowmul[4];            // halves between consecutive tile rows: tiles_w * chunk_l (level 3: tiles_w3 * 4)
  int hp[4], xs[4];         // padded extents: rows, and 8-column strips per row
  int h0, w0;
  const float* coords;      // [nq][2] (x, y)
  __half* out16;            // [nq][out_pitch] (may be null)
  float* out32;             // [nq][324] un-rounded results (tests; may be null)
  long long out_pitch;
  long long nq;
};

// window origin of a level: same arithmetic as the first kernels (level coordinate c / 2^l, exact; clamped so that
// far-away coordinates cannot overflow -- windows that start 8 texels outside the map are all zeros either way)
__device__ __forceinline__ void origin(float c, float inv, int size, int& i0, float& frac) {
  const float v = fminf(fmaxf(c * inv, -8.0f), static_cast<float>(size + 8));
  const float f = floorf(v);
  frac = v - f;
  i0 = static_cast<int>(f) - 4;
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// per-lane constants of the lane's staging level (lane >> 3).  The level's 30 pieces (10 window rows x 3 strips) are
// dealt so that the row arithmetic is shared: lane j (0..7) of the group takes window row j with strips 0, 1, 2, and
// lanes 0..5 take (row 8 + j / 3, strip j % 3) as a fourth piece.
struct LevelLane {
  const __half* base;       // level pointer
  long long qstride;
  float inv;                // 2^-l
  int w, h;                 // map size of the level (for the origin clamp)
  int rowmul, hp, xs;
  int rshift, rmask, sshift;
  int r0, r3, sx3;          // window row of pieces 0..2; row / strip of piece 3 (r3 = kRows: none)
  uint32_t dst0, dst3;      // shared-memory offsets inside a window buffer (pieces 1, 2: dst0 + 16, + 32)
};

// issue the pieces of query q for this lane's level; returns the lane's window origin (ix) and fractional offsets
__device__ __forceinline__ void stage(const LevelLane& c, const float* coords, long long q, uint32_t win_u32, int& ix_out, float& fx, float& fy) {
  const float2 cxy = __ldg(reinterpret_cast<const float2*>(coords) + q);
  int ix, iy;
  origin(cxy.x, c.inv, c.w, ix, fx);
  origin(cxy.y, c.inv, c.h, iy, fy);
  ix_out = ix;
  const __half* base = c.base + q * c.qstride;
  const int xs0 = ix >> 3;                                        // arithmetic shift: floor for negative origins
  const int last = ((ix & 7) + 9) >> 3;                           // last strip (counted from xs0) the window touches: 1 or 2
  {
    const int y = iy + c.r0;
    const bool yok = static_cast<unsigned>(y) < static_cast<unsigned>(c.hp);
    const __half* row = base + ((y >> c.rshift) * c.rowmul + ((y & c.rmask) << 3) + (xs0 << c.sshift));
    const int step = 1 << c.sshift;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k <= last) {
        const bool ok = yok && static_cast<unsigned>(xs0 + k) < static_cast<unsigned>(c.xs);
        cp_async_16(win_u32 + c.dst0 + k * 16, ok ? row + k * step : base, ok ? 16u : 0u);
      }
    }
  }
  if (c.r3 < kRows && c.sx3 <= last) {
    const int y = iy + c.r3, xs = xs0 + c.sx3;
    const bool ok = static_cast<unsigned>(y) < static_cast<unsigned>(c.hp) && static_cast<unsigned>(xs) < static_cast<unsigned>(c.xs);
    const int off = (y >> c.rshift) * c.rowmul + ((y & c.rmask) << 3) + (xs << c.sshift);
    cp_async_16(win_u32 + c.dst3, ok ? base + off : base, ok ? 16u : 0u);
  }
}

template <bool OUT32>
__global__ void __launch_bounds__(kWarps * 32, 4) corr_lookup_strip_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  pdl_launch_dependents();
  pdl_wait();               // the coordinates come from the preceding kernel (ATDN_PDL: only the launch overlaps its tail)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* wbase = smem + warp * kWarpBytes;
  const uint32_t win_u32 = static_cast<uint32_t>(__cvta_generic_to_shared(wbase));
  __half* out_s = reinterpret_cast<__half*>(wbase + 2 * kWinBytes);
  const long long q0 = (static_cast<long long>(blockIdx.x) * kWarps + warp) * kQueriesPerWarp;
  if (q0 >= p.nq) return;
  const long long left = p.nq - q0;
  const int n = left < kQueriesPerWarp ? static_cast<int>(left) : kQueriesPerWarp;

  // staging role: level lg = lane >> 3 (see LevelLane)
  const int lg = lane >> 3;
  LevelLane c;
  c.base = p.lvl[lg];
  c.qstride = p.qstride[lg];
  c.inv = __int_as_float((127 - lg) << 23);
  c.w = p.w0 >> lg;
  c.h = p.h0 >> lg;
  c.rowmul = p.rowmul[lg];
  c.hp = p.hp[lg];
  c.xs = p.xs[lg];
  c.rshift = 3 - lg;
  c.rmask = (8 >> lg) - 1;
  c.sshift = 6 - lg;
  {
    const int j = lane & 7;
    c.r0 = j;
    c.r3 = j < 6 ? 8 + j / 3 : kRows;
    c.sx3 = j % 3;
    c.dst0 = lg * kLevelBytes + c.r0 * (kCols * 2);
    c.dst3 = lg * kLevelBytes + c.r3 * (kCols * 2) + c.sx3 * 16;
  }
  // blend role: lane (level bl, column pair ap), lanes 0..19
  const int bl = lane / 5, ap = lane - 5 * bl;

  int ix_cur, ix_nxt = 0;
  float fx_cur, fy_cur, fx_nxt = 0.0f, fy_nxt = 0.0f;
  stage(c, p.coords, q0, win_u32, ix_cur, fx_cur, fy_cur);
  cp_async_commit();

#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    const uint32_t cur_off = (i & 1) * kWinBytes;
    if (i + 1 < n) stage(c, p.coords, q0 + i + 1, win_u32 + (kWinBytes - cur_off), ix_nxt, fx_nxt, fy_nxt);
    cp_async_commit();
    cp_async_wait<1>();                              // everything but the pieces just issued has landed (this lane's share)
    __syncwarp();                                    // ... and every other lane's
    const uint8_t* win = wbase + cur_off;

    // separable blend: lane (level l = lane / 5, ap = lane % 5), lanes 0..19, sweeps the 10 window rows of columns a = 2 ap and
    // a + 1 (three texel loads for two columns); channel = l*81 + a*9 + b.  The origin / fractions of level l live in the
    // lanes of its staging group (8 l .. 8 l + 7).
    {
      const int l = bl < 4 ? bl : 3;
      const int ix = __shfl_sync(0xffffffffu, ix_cur, 8 * l);
      const float fxl = __shfl_sync(0xffffffffu, fx_cur, 8 * l), fyl = __shfl_sync(0xffffffffu, fy_cur, 8 * l);
      if (lane < 20) {
        const __half* wp = reinterpret_cast<const __half*>(win + l * kLevelBytes) + (ix & 7) + 2 * ap;
        float va[9], vb[9];
        float ha_prev = 0.0f, hb_prev = 0.0f;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
          const float t0 = __half2float(wp[r * kCols]), t1 = __half2float(wp[r * kCols + 1]), t2 = __half2float(wp[r * kCols + 2]);
          const float ha = fmaf(fxl, t1 - t0, t0), hb = fmaf(fxl, t2 - t1, t1);
          if (r > 0) {
            va[r - 1] = fmaf(fyl, ha - ha_prev, ha_prev);
            vb[r - 1] = fmaf(fyl, hb - hb_prev, hb_prev);
            if constexpr (OUT32) {
              p.out32[(q0 + i) * 324 + l * 81 + (2 * ap) * 9 + r - 1] = va[r - 1];
              if (ap < 4) p.out32[(q0 + i) * 324 + l * 81 + (2 * ap + 1) * 9 + r - 1] = vb[r - 1];
            }
          }
          ha_prev = ha;
          hb_prev = hb;
        }
        // the lane's 18 (ap = 4: 9) consecutive halves start at half l * 81 + 18 ap, odd for odd levels: aligned words plus
        // one single half at the front (odd) or at the back
        const int start = l * 81 + 18 * ap;
        const bool odd = (l & 1) != 0;
        uint32_t* ow = reinterpret_cast<uint32_t*>(out_s) + ((start + 1) >> 1);
        const float s9[18] = {va[0], va[1], va[2], va[3], va[4], va[5], va[6], va[7], va[8], vb[0], vb[1], vb[2], vb[3], vb[4], vb[5], vb[6], vb[7], vb[8]};
        const int nwords = ap < 4 ? 8 : 4;                        // whole words after the (possible) leading single
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (k < nwords) {
            const __half2 pk = __floats2half2_rn(odd ? s9[2 * k + 1] : s9[2 * k], odd ? s9[2 * k + 2] : s9[2 * k + 1]);
            ow[k] = *reinterpret_cast<const uint32_t*>(&pk);
          }
        }
        if (ap < 4) {
          if (odd) {                                               // halves 0 and 17 are singles
            out_s[start] = __float2half_rn(s9[0]);
            out_s[start + 17] = __float2half_rn(s9[17]);
          } else {                                                 // ninth word
            const __half2 pk = __floats2half2_rn(s9[16], s9[17]);
            ow[8] = *reinterpret_cast<const uint32_t*>(&pk);
          }
        } else {
          out_s[odd ? start : start + 8] = __float2half_rn(odd ? s9[0] : s9[8]);
        }
      }
    }
    __syncwarp();

    // coalesced output: 40 16-byte vectors + channels 320..323
    if (p.out16) {
      __half* dst = p.out16 + (q0 + i) * p.out_pitch;
      const uint4* src = reinterpret_cast<const uint4*>(out_s);
      *reinterpret_cast<uint4*>(dst + lane * 8) = src[lane];
      if (lane < 8) *reinterpret_cast<uint4*>(dst + 256 + lane * 8) = src[32 + lane];
      else if (lane == 8) *reinterpret_cast<uint2*>(dst + 320) = *reinterpret_cast<const uint2*>(out_s + 320);
    }
    __syncwarp();                                    // out_s and the window of query i are free again
    ix_cur = ix_nxt;
    fx_cur = fx_nxt;
    fy_cur = fy_nxt;
  }
}

}  // namespace lks
}  // namespace atdn
