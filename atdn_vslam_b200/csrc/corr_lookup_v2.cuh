// Correlation lookup on the tiled fp16 pyramid, second layout (EXPERIMENTAL: selected with ATDN_LOOKUP_V2=1; not yet
// run on a GPU -- its logic is checked on the host by tools/lookup_v2_emulate.cu, which executes the phase functions
// below thread by thread against the formulas of corr_lookup_half_kernel).
// GMA.whl!/GMA/core/corr.py:32-53 + utils/utils.py:59-73.
//
// Why: corr_lookup_half_kernel (one warp per query) executes ~600 warp instructions per query and is bound by
// instruction issue (1.8 TB/s of algorithmic traffic, 28% of the HBM copy bandwidth): its blend runs on 18 of 32
// lanes and every lane repeats the per-query coordinate arithmetic.  Here a CTA of 128 threads owns 32 queries:
//   phase 0  thread (level, query): window origin, fractional offsets and the tiled-layout offsets of the 10 window
//            rows and 4 row segments, once, into shared-memory tables;
//   phase 1  warp w stages the windows of queries 8w..8w+7 with the access pattern of the first kernel (lane = (row,
//            4-texel segment), 5 passes per query, coalesced 8-byte loads, 4 queries in flight) as raw fp16; an item
//            costs two table reads and an add instead of ~25 integer instructions;
//   phase 2  thread (level = warp, query = lane) blends its whole 9 x 9 level window from shared memory in registers
//            (all 32 lanes busy, no cross-lane traffic): ~90 warp instructions per query instead of ~170;
//   phase 3  the 81 fp16 results per thread go to a [query][324] staging tile (aliasing the windows) and leave as
//            16-byte vectors of whole 648-byte rows.
// Same fp32 formulas and operand order as the first kernel, so the fp16 outputs are bit-identical.
// Static SASS count: see DESIGN.md section 6.
#pragma once
#include <math.h>
#include <stdint.h>

#include <cuda_fp16.h>

namespace atdn {
namespace lk2 {

#define LK2_HD __host__ __device__ __forceinline__

constexpr int kQ = 32;                          // queries per CTA
constexpr int kThreads = 128;                   // 4 warps = 4 levels in phase 2
constexpr int kWinBytes = 336;                  // 10 rows x 16 fp16 + 16: lane stride 84 words -> conflict-free 128-bit reads
constexpr int kWinRegion = 4 * kQ * kWinBytes;  // 43008
constexpr int kOutPitch = 326;                  // halves per staged output row: 163 words (odd) -> conflict-free stores
constexpr int kOrgOffset = kWinRegion;
constexpr int kOrgBytes = 80;                   // per (level, query) record, see Org
constexpr int kSmemBytes = kWinRegion + 4 * kQ * kOrgBytes;   // 53248: dynamic shared memory (opt-in above 48 KiB)
constexpr int kInvalid = -(1 << 30);            // row / segment outside the map: any sum with it stays negative
static_assert(kQ * kOutPitch * 2 <= kWinRegion, "output staging aliases the window region");

struct Params {
  const __half* lvl[4];     // tiled: level l = [query][tile][(8 >> l) x (32 >> l)]
  int tiles, tiles_w;
  int h0, w0;
  const float* coords;      // [nq][2] (x, y)
  __half* out16;            // [nq][out_pitch]
  long long out_pitch;
  long long nq;
};

struct Org {                  // written once per (level, query) in phase 0: everything phase 1 and 2 need
  float fx, fy;               // fractional offsets shared by the 81 taps of the level
  int o;                      // ix & 3: first needed texel inside the staged row (row base = ix & ~3)
  int nvalid;                 // texels of the staged 16-texel row that lie inside the map (counted from the row base)
  int row[10];                // half offset of window row r inside the query's level maps (tile row part), or kInvalid
  int seg[4];                 // ... of 4-texel segment s (tile column part), or kInvalid
  int pad[2];
};
static_assert(sizeof(Org) == kOrgBytes, "Org layout");

struct Regs {               // what a thread carries from the blend (phase 2) across the CTA barrier into the stores
  uint32_t pk[41];          // halves c = a * 9 + b of this thread's level window, two per word
};

LK2_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t shift) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, shift);
#else
  shift &= 31u;
  return shift ? ((lo >> shift) | (hi << (32u - shift))) : lo;
#endif
}

LK2_HD uint32_t pack2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

LK2_HD float2 unpack2(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }

// window origin of level l: same arithmetic as lk_origin in flow_ops.cu
LK2_HD void origin(float c, int l, int size, int& i0, float& frac) {
  const float inv = 1.0f / static_cast<float>(1 << l);          // exact power of two
  const float v = fminf(fmaxf(c * inv, -8.0f), static_cast<float>(size + 8));
  const float f = floorf(v);
  frac = v - f;
  i0 = static_cast<int>(f) - 4;
}

LK2_HD Org* org_slot(uint8_t* smem, int l, int ql) { return reinterpret_cast<Org*>(smem + kOrgOffset) + l * kQ + ql; }

// ---- phase 0: thread (l = tid / 32, ql = tid % 32) ----------------------------------------------------------------
// Tiled level l: half offset of texel (y, x) inside a query's maps = row[y] + seg[x]:
//   row part = ((y >> (3-l)) * tiles_w) << (8-2l)  +  ((y & ((8>>l)-1)) << (5-l)),   seg part = ((x >> (5-l)) << (8-2l)) + (x & ((32>>l)-1))
LK2_HD void phase_origin(const Params& p, long long qbase, int tid, uint8_t* smem) {
  const int l = tid >> 5, ql = tid & 31;
  const long long q = qbase + ql;
  Org* o = org_slot(smem, l, ql);
  if (q >= p.nq) return;                                 // never staged, blended on garbage, never copied out
  const int H = p.h0 >> l, W = p.w0 >> l;
  const float2 cxy = *reinterpret_cast<const float2*>(p.coords + q * 2);
  int ix, iy;
  float fx, fy;
  origin(cxy.x, l, W, ix, fx);
  origin(cxy.y, l, H, iy, fy);
  o->fx = fx;
  o->fy = fy;
  o->o = ix & 3;
  const int xb = ix & ~3;
  const int nv = W - xb;
  o->nvalid = nv < 0 ? 0 : (nv > 16 ? 16 : nv);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const int y = iy + r;
    o->row[r] = (static_cast<unsigned>(y) < static_cast<unsigned>(H))
                    ? ((((y >> (3 - l)) * p.tiles_w) << (8 - 2 * l)) + ((y & ((8 >> l) - 1)) << (5 - l))) : kInvalid;
  }
#pragma unroll
  for (int sgm = 0; sgm < 4; ++sgm) {
    const int x = xb + 4 * sgm;
    o->seg[sgm] = (static_cast<unsigned>(x) < static_cast<unsigned>(W)) ? (((x >> (5 - l)) << (8 - 2 * l)) + (x & ((32 >> l) - 1))) : kInvalid;
  }
}

// ---- phase 1: warp `warp` stages queries ql = 8 * warp + j; item k of a lane = (level, window row, 4-texel segment):
// passes 0..3 = rows 0..7 of level k (lane = row * 4 + segment), pass 4 = rows 8, 9 of level lane / 8.  All index
// arithmetic comes from the phase-0 tables: two shared-memory reads and an add per item.
struct StageLane {   // per-lane constants of the five passes (byte offsets for query 0 of the CTA)
  int row_at[5];     // Org::row[r] of the item's level
  int seg_at[5];     // Org::seg[s]
  int dst[5];        // the lane's 8-byte slot in the level's window
  int l4;            // level of pass 4
  const __half* lvl4;
};

LK2_HD void stage_lane_init(const Params& p, int lane, StageLane& c) {
  c.l4 = lane >> 3;
  c.lvl4 = p.lvl[c.l4];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int l = k < 4 ? k : c.l4;
    const int r = k < 4 ? (lane >> 2) : 8 + ((lane >> 2) & 1);
    const int sgm = lane & 3;
    c.row_at[k] = kOrgOffset + l * kQ * kOrgBytes + 16 + r * 4;
    c.seg_at[k] = kOrgOffset + l * kQ * kOrgBytes + 56 + sgm * 4;
    c.dst[k] = l * kQ * kWinBytes + r * 32 + sgm * 8;
  }
}

// `base` = first half of query q's level maps; rows / segments outside the map are staged as zeros (zero padding of
// grid_sample); texels past the right edge inside a loaded segment are masked in phase 2
LK2_HD uint2 stage_load(const __half* base, int k, int ql, const StageLane& c, const uint8_t* smem) {
  const int off = *reinterpret_cast<const int*>(smem + c.row_at[k] + ql * kOrgBytes) +
                  *reinterpret_cast<const int*>(smem + c.seg_at[k] + ql * kOrgBytes);
  uint2 raw = make_uint2(0u, 0u);
  if (off >= 0) {
    const uint2* src = reinterpret_cast<const uint2*>(base + off);
#ifdef __CUDA_ARCH__
    raw = __ldg(src);
#else
    raw = *src;
#endif
  }
  return raw;
}

LK2_HD void phase_stage(const Params& p, long long qbase, int warp, int lane, uint8_t* smem) {
  constexpr int G = 4;                                   // queries in flight per warp: 20 independent 8-byte loads per lane
  StageLane c;
  stage_lane_init(p, lane, c);
#pragma unroll 1
  for (int g = 0; g < 8; g += G) {
    uint2 raw[G][5];
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int ql = warp * 8 + g + j;
      const long long q = qbase + ql;
      if (q < p.nq) {                                    // warp-uniform
        const long long qt = q * p.tiles;
#pragma unroll
        for (int k = 0; k < 4; ++k) raw[j][k] = stage_load(p.lvl[k] + (qt << (8 - 2 * k)), k, ql, c, smem);
        raw[j][4] = stage_load(c.lvl4 + (qt << (8 - 2 * c.l4)), 4, ql, c, smem);
      }
    }
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int ql = warp * 8 + g + j;
      if (qbase + ql < p.nq) {
#pragma unroll
        for (int k = 0; k < 5; ++k) *reinterpret_cast<uint2*>(smem + c.dst[k] + ql * kWinBytes) = raw[j][k];
      }
    }
  }
}

// ---- phase 2: thread (level l = warp, query ql = lane) blends its 9 x 9 window -------------------------------------
// Output channel of the level = a * 9 + b, a = x offset (slow), b = y offset (corr.py:40-46).  Row r of the window
// yields the horizontal lerps h[a]; rows r-1 and r give the outputs (a, b = r - 1).
LK2_HD void phase_blend(int l, int ql, const uint8_t* smem, Regs& R) {
  const Org* o = reinterpret_cast<const Org*>(smem + kOrgOffset) + l * kQ + ql;
  const uint8_t* win = smem + (l * kQ + ql) * kWinBytes;
  const int first_texel = o->o, nvalid = o->nvalid;
  const bool skip_word = (first_texel & 2) != 0;        // first needed texel sits in word 0 or 1 of the row ...
  const uint32_t sh = (first_texel & 1) ? 16u : 0u;     // ... in its low or high half
  const float fx = o->fx, fy = o->fy;
  // texels past the right map edge inside a tile hold pooling leftovers: word j of the (word-aligned) row keeps
  // both / the low / none of its halves
  uint32_t keep[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int left = nvalid - 2 * (j + (skip_word ? 1 : 0));
    keep[j] = left >= 2 ? 0xffffffffu : (left == 1 ? 0xffffu : 0u);
  }
  float hprev[9];
  float pend[9];                                         // (a, b) results waiting for the other half of their word
  float first[9];                                        // (a, 0) for odd a: high half of the word that (a-1, 8) completes
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint4 lo = *reinterpret_cast<const uint4*>(win + r * 32);
    const uint4 hi = *reinterpret_cast<const uint4*>(win + r * 32 + 16);
    const uint32_t raw[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t v[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) v[j] = (skip_word ? raw[j + 1] : raw[j]) & keep[j];
    float t[10];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float2 f = unpack2(funnel_r(v[j], v[j + 1], sh));
      t[2 * j] = f.x;
      t[2 * j + 1] = f.y;
    }
#pragma unroll
    for (int a = 0; a < 9; ++a) {
      const float h = fmaf(fx, t[a + 1] - t[a], t[a]);
      if (r > 0) {
        const int b = r - 1;
        const float res = fmaf(fy, h - hprev[a], hprev[a]);
        const int c = a * 9 + b;
        if ((c & 1) == 0) {
          if (c == 80) R.pk[40] = pack2(res, 0.0f);                       // last half of the level: low half only
          else if (b == 8) R.pk[c >> 1] = pack2(res, first[a + 1]);       // partner (a + 1, 0) was computed at r = 1
          else pend[a] = res;                                              // partner (a, b + 1) comes with the next row
        } else {
          if (b == 0) first[a] = res;                                      // completed by (a - 1, 8) at r = 9
          else R.pk[c >> 1] = pack2(pend[a], res);
        }
      }
      hprev[a] = h;
    }
  }
}

// ---- phase 3a: the same thread writes its 81 halves into the [ql][324] staging tile (after a CTA barrier: the tile
// aliases the windows).  Level l starts at half 81 * l: word-aligned for even l, off by one half for odd l.
LK2_HD void phase_scatter(int l, int ql, uint8_t* smem, const Regs& R) {
  uint8_t* row = smem + ql * (kOutPitch * 2);
  if ((l & 1) == 0) {
    uint32_t* w = reinterpret_cast<uint32_t*>(row + 81 * l * 2);
#pragma unroll
    for (int m = 0; m < 40; ++m) w[m] = R.pk[m];
    *reinterpret_cast<uint16_t*>(row + (81 * l + 80) * 2) = static_cast<uint16_t>(R.pk[40] & 0xffffu);
  } else {
    *reinterpret_cast<uint16_t*>(row + 81 * l * 2) = static_cast<uint16_t>(R.pk[0] & 0xffffu);
    uint32_t* w = reinterpret_cast<uint32_t*>(row + (81 * l + 1) * 2);
#pragma unroll
    for (int m = 0; m < 40; ++m) w[m] = funnel_r(R.pk[m], R.pk[m + 1], 16u);
  }
}

// ---- phase 3b: warp w copies the staged rows 8w..8w+7 to global memory: 40 16-byte vectors + one 8-byte tail per row
LK2_HD void phase_copy(const Params& p, long long qbase, int warp, int lane, const uint8_t* smem) {
  const long long q0 = qbase + warp * 8;
  const long long left = p.nq - q0;
  const int rows = left >= 8 ? 8 : (left > 0 ? static_cast<int>(left) : 0);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(smem) + (warp * 8) * (kOutPitch / 2) + lane * 4;
  __half* dst = p.out16 + q0 * p.out_pitch + lane * 8;
#pragma unroll 1
  for (int j = 0; j < rows; ++j) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(src[0], src[1], src[2], src[3]);                       // vectors 0..31
    if (lane < 8) *reinterpret_cast<uint4*>(dst + 256) = make_uint4(src[128], src[129], src[130], src[131]);   // 32..39
    else if (lane == 8) *reinterpret_cast<uint2*>(dst + 256) = make_uint2(src[128], src[129]);         // channels 320..323
    src += kOutPitch / 2;
    dst += p.out_pitch;
  }
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(kThreads) corr_lookup_v2_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) uint8_t smem[];        // kSmemBytes, dynamic (above the 48 KiB static limit)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long qbase = static_cast<long long>(blockIdx.x) * kQ;
  phase_origin(p, qbase, tid, smem);
  __syncthreads();
  phase_stage(p, qbase, warp, lane, smem);
  __syncthreads();
  Regs R;
  phase_blend(warp, lane, smem, R);
  __syncthreads();                                       // every window has been read: reuse the region as output staging
  phase_scatter(warp, lane, smem, R);
  __syncthreads();
  phase_copy(p, qbase, warp, lane, smem);
}
#endif

}  // namespace lk2
}  // namespace atdn
