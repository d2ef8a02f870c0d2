// Gather / element-wise kernels of the GMA flow path (HBM- or latency-bound, CUDA cores).
// See include/atdn_b200.h for the reference call sites each entry point replaces.
#include <math.h>
#include <stdlib.h>

#include <cuda_fp16.h>

#include "common.h"
#include "tc_host.cuh"
#include "tc_ptx.cuh"
#include "corr_lookup_strip.cuh"

namespace atdn {

// ------------------------------------------------------------------------------------------------
// Correlation lookup: one WARP per query pixel, 8 queries per CTA.
//   1. lanes compute the 4 x (9 + 9) tap coordinates of the query (per level: 9 x and 9 y positions shared by
//      the 81 taps) exactly like the reference's round trip through normalised grid coordinates (utils.py:63-64
//      then ATen's align_corners un-normalisation; explicit _rn intrinsics so nvcc cannot contract into FMAs);
//   2. the 12 x 12 texel window of every level that covers all taps (+-1 texel of rounding slack) is staged in
//      shared memory with row-contiguous loads: all 18 loads of a lane are independent and in flight together,
//      and out-of-image texels are staged as the zeros of grid_sample's zero padding;
//   3. each lane blends its taps from shared memory and writes two adjacent channels per store.
// The previous version (one CTA per query, 4 dependent scattered loads per tap) ran at 16% of the HBM roofline.
// ------------------------------------------------------------------------------------------------
struct LookupParams {
  const float* lvl[4];
  int pitch[4];
  int h[4], w[4];
};

__device__ __forceinline__ float grid_round_trip(float x, int size) {
  const float sm1 = static_cast<float>(size - 1);
  const float xn = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, x), sm1), 1.0f);
  return __fmul_rn(__fdiv_rn(__fadd_rn(xn, 1.0f), 2.0f), sm1);
}

constexpr int kLkWarps = 8;
constexpr int kLkWinW = 16, kLkWinH = 10;        // staged window: 10 rows x up to 16 columns (column base aligned to 4 texels)

// One WARP per query pixel, 8 queries per CTA.  The first version of this layout (scalar window loads, per-tap index
// arithmetic in the blend loop) executed ~1200 instructions per lane and query and was issue-bound at 21% of the HBM
// roofline (ncu: SM throughput 81%).  Now
//   2. the window base is aligned to 4 texels, so the 10 x 10 texels the taps touch arrive as (at most) two 128-bit
//      loads per lane, all issued before the first use;
//      texels outside the image (incl. the row pad) are zeroed = grid_sample's zero padding;
//   3. lane (b, a0) blends 3 x-taps of one y-tap per level (27 lanes): the y coordinate / weight is fetched once per
//      level, results go to shared memory and leave as 16-byte vectors of 8 consecutive channels.
__global__ void __launch_bounds__(kLkWarps * 32) corr_lookup_kernel(LookupParams p, const float* __restrict__ coords,
                                                                    __half* __restrict__ out16, long long out_pitch,
                                                                    float* __restrict__ out32, long long nq) {
  // win doubles as the output staging buffer (324 floats) once the blend results sit in registers
  __shared__ __align__(16) float win[kLkWarps][4][kLkWinH * kLkWinW];
  __shared__ float cfr[kLkWarps][72];            // fractional part of the 4 x (9 x + 9 y) tap coordinates
  __shared__ int cin[kLkWarps][72];              // integer part (floor), clamped to [-2, size]
  const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = static_cast<long long>(blockIdx.x) * kLkWarps + wq;
  if (q >= nq) return;
  const float2 cxy = *reinterpret_cast<const float2*>(coords + q * 2);
  const float cx = cxy.x, cy = cxy.y;
  const int h0 = p.h[0], w0 = p.w[0];            // level l is (h0 >> l) x (w0 >> l): no indexed parameter reads

  // 1. tap coordinates: idx = level * 18 + axis * 9 + k
  for (int idx = lane; idx < 72; idx += 32) {
    const int l = idx / 18, rem = idx - l * 18, axis = rem / 9, k = rem - axis * 9;
    const float inv = 1.0f / static_cast<float>(1 << l);     // exact: division by a power of two
    const int size = (axis ? h0 : w0) >> l;
    const float c = (axis ? cy : cx) * inv;
    const float v = grid_round_trip(__fadd_rn(c, static_cast<float>(k - 4)), size);
    const float vf = floorf(v);
    cfr[wq][idx] = v - vf;
    // clamp before the int conversion so that far-away coordinates cannot overflow
    cin[wq][idx] = static_cast<int>(fminf(fmaxf(vf, -2.0f), static_cast<float>(size)));
  }
  // 2. window loads: lane -> (row = lane / 4 [+ 8], 4-texel segment = lane % 4); every lane derives the origins itself
  //    and ALL loads are issued before the first use
  float4 v[4][2];
  int xa[4], y0w[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const float inv = 1.0f / static_cast<float>(1 << l);
    const int H = h0 >> l, W = w0 >> l;
    // the 81 taps touch texels floor(c) - 4 .. floor(c) + 5 on each axis; a coordinate that the normalised-grid round
    // trip pushes across an integer (1 ulp) falls outside this window and takes the exact direct-load path below
    const int ox = static_cast<int>(fminf(fmaxf(floorf(cx * inv), -16.0f), static_cast<float>(W + 16))) - 4;
    const int oy = static_cast<int>(fminf(fmaxf(floorf(cy * inv), -16.0f), static_cast<float>(H + 16))) - 4;
    xa[l] = ox & ~3;                              // aligned down to a multiple of 4 (two's complement floor)
    y0w[l] = oy;
    const int pitch = l == 0 ? p.pitch[0] : l == 1 ? p.pitch[1] : l == 2 ? p.pitch[2] : p.pitch[3];
    const float* lv = l == 0 ? p.lvl[0] : l == 1 ? p.lvl[1] : l == 2 ? p.lvl[2] : p.lvl[3];
    const float* __restrict__ base = lv + q * static_cast<long long>(H) * pitch;
    const int x = xa[l] + (lane & 3) * 4;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int r = pass * 8 + (lane >> 2);
      const int y = oy + r;
      v[l][pass] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (r < kLkWinH && y >= 0 && y < H && x >= 0 && x + 4 <= pitch && x < ox + 10)   // segments right of the taps are never read
        v[l][pass] = __ldg(reinterpret_cast<const float4*>(base + static_cast<long long>(y) * pitch + x));
    }
  }
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const int W = w0 >> l;
    const int x = xa[l] + (lane & 3) * 4;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int r = pass * 8 + (lane >> 2);
      if (r < kLkWinH) {
        float4 t = v[l][pass];
        if (x + 3 >= W) {                          // row pad (and the columns TMA clipping left behind) is not image
          if (x >= W) t.x = 0.0f;
          if (x + 1 >= W) t.y = 0.0f;
          if (x + 2 >= W) t.z = 0.0f;
          t.w = 0.0f;
        }
        *reinterpret_cast<float4*>(&win[wq][l][r * kLkWinW + (lane & 3) * 4]) = t;
      }
    }
  }
  __syncwarp();

  // 3. blend: channel ch = l * 81 + a * 9 + b  (a offsets x, b offsets y -- corr.py:40-46); lane -> (b, a0..a0+2)
  float res[4][3];
  const int b = lane / 3, a0 = (lane - b * 3) * 3;
  if (lane < 27) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int y0 = cin[wq][l * 18 + 9 + b];
      const float fy = cfr[wq][l * 18 + 9 + b];
      const int XA = xa[l], wy = y0 - y0w[l];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int a = a0 + i;
        const int x0 = cin[wq][l * 18 + a];
        const float fx = cfr[wq][l * 18 + a];
        const int wx = x0 - XA;
        float v00, v01, v10, v11;
        if (wx >= 0 && wx + 1 < kLkWinW && wy >= 0 && wy + 1 < kLkWinH) {
          const float* wp = &win[wq][l][wy * kLkWinW + wx];
          v00 = wp[0]; v01 = wp[1]; v10 = wp[kLkWinW]; v11 = wp[kLkWinW + 1];
        } else {   // not reachable for finite coordinates near the image; kept exact for robustness
          const int H = h0 >> l, W = w0 >> l;
          const int pitch = l == 0 ? p.pitch[0] : l == 1 ? p.pitch[1] : l == 2 ? p.pitch[2] : p.pitch[3];
          const float* lv = l == 0 ? p.lvl[0] : l == 1 ? p.lvl[1] : l == 2 ? p.lvl[2] : p.lvl[3];
          const float* base = lv + q * static_cast<long long>(H) * pitch;
          const bool xin0 = x0 >= 0 && x0 < W, xin1 = x0 + 1 >= 0 && x0 + 1 < W;
          const bool yin0 = y0 >= 0 && y0 < H, yin1 = y0 + 1 >= 0 && y0 + 1 < H;
          const float* r0 = base + static_cast<long long>(y0) * pitch + x0;
          v00 = (xin0 && yin0) ? __ldg(r0) : 0.0f;
          v01 = (xin1 && yin0) ? __ldg(r0 + 1) : 0.0f;
          v10 = (xin0 && yin1) ? __ldg(r0 + pitch) : 0.0f;
          v11 = (xin1 && yin1) ? __ldg(r0 + pitch + 1) : 0.0f;
        }
        res[l][i] = v00 * ((1.0f - fx) * (1.0f - fy)) + v01 * (fx * (1.0f - fy)) + v10 * ((1.0f - fx) * fy) + v11 * (fx * fy);
      }
    }
  }
  __syncwarp();                                    // every lane is done with the windows: reuse them as output staging
  float* outs = &win[wq][0][0];
  if (lane < 27) {
#pragma unroll
    for (int l = 0; l < 4; ++l)
#pragma unroll
      for (int i = 0; i < 3; ++i) outs[l * 81 + (a0 + i) * 9 + b] = res[l][i];
  }
  __syncwarp();

  // 4. coalesced output: 8 consecutive channels per lane and store
  for (int c8 = lane; c8 < 41; c8 += 32) {
    const float4 f0 = *reinterpret_cast<const float4*>(&outs[c8 * 8]);
    if (c8 < 40) {
      const float4 f1 = *reinterpret_cast<const float4*>(&outs[c8 * 8 + 4]);
      if (out16) {
        __half2 h0 = __floats2half2_rn(f0.x, f0.y), h1 = __floats2half2_rn(f0.z, f0.w);
        __half2 h2 = __floats2half2_rn(f1.x, f1.y), h3 = __floats2half2_rn(f1.z, f1.w);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2);
        u.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(out16 + q * out_pitch + c8 * 8) = u;
      }
      if (out32) {
        *reinterpret_cast<float4*>(out32 + q * 324 + c8 * 8) = f0;
        *reinterpret_cast<float4*>(out32 + q * 324 + c8 * 8 + 4) = f1;
      }
    } else {   // channels 320..323
      if (out16) {
        __half2 h0 = __floats2half2_rn(f0.x, f0.y), h1 = __floats2half2_rn(f0.z, f0.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(out16 + q * out_pitch + 320) = u;
      }
      if (out32) *reinterpret_cast<float4*>(out32 + q * 324 + 320) = f0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The two 7x7 layers whose input has 3 / 2 channels (too thin for a 64-channel K chunk) run on the halo
// implicit-GEMM kernel after folding the horizontal taps into channels ("row packing"): the packed tensor is
// 3.2x (stem) / 6.5x (flow) smaller than a full im2col matrix and the vertical taps stay implicit.
//
// Stem (extractor.py:173, 7x7 stride 2 pad 3 on the normalised image, network.py:75-76):
//   X[b, y2, ox, (ry*3 + c)*8 + xx] = 2*(img[b, c, 2*y2 + ry, 2*(ox-2) + xx]/255) - 1   (0 outside the image)
//   out[oy, ox] = sum_{ai<4} sum_ch W4[ai, ch] X[oy + ai - 2, ox, ch],  W4[ai, (ry,c,xx)] = w[c, 2*ai+ry-1, xx-1]
// i.e. a 4x1 convolution with 48 input channels, top padding 2 (rows past the image are TMA zero fill).
// One thread = one 16-byte group of 8 consecutive xx = 8 consecutive image pixels of one (row, plane).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_pack_kernel(const float* __restrict__ img, __half* __restrict__ x, int B, int H,
                                                        int W, int OH, int OW) {
  pdl_launch_dependents();
  pdl_wait();
  // grid = (ceil(OW*6/256), OH, B): no 64-bit divisions
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= static_cast<unsigned>(OW) * 6u) return;
  const int g = static_cast<int>(idx % 6u), ox = static_cast<int>(idx / 6u);
  const int y2 = blockIdx.y, b = blockIdx.z;
  const long long plane = static_cast<long long>(H) * W;
  const long long pix = (static_cast<long long>(b) * OH + y2) * OW + ox;
  {
    const int ry = g / 3, c = g - ry * 3;
    const float* row = img + (static_cast<long long>(b) * 3 + c) * plane + static_cast<long long>(2 * y2 + ry) * W;
    const int x0 = 2 * ox - 4;
    float v[8];
    if (x0 >= 0 && x0 + 8 <= W) {   // x0 is even: 8-byte aligned pairs (W is even)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __ldg(reinterpret_cast<const float2*>(row + x0) + j);
        v[2 * j] = f.x;
        v[2 * j + 1] = f.y;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 2.0f * (v[j] / 255.0f) - 1.0f;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ix = x0 + j;
        v[j] = (ix >= 0 && ix < W) ? 2.0f * (__ldg(row + ix) / 255.0f) - 1.0f : 0.0f;
      }
    }
    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2);
    u.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(x + pix * 48 + g * 8) = u;
  }
}

// Flow (update.py:79 convf1, 7x7 pad 3 on the 2-channel flow): X[b, y, x, dx*2 + c] = flow[b, y, x + dx - 3, c]
// (14 channels + 2 zeros, 32 bytes per pixel); the 7 vertical taps stay implicit (7x1 convolution, pad 3).
__global__ void __launch_bounds__(256) flow_pack_kernel(const float* __restrict__ flow, __half* __restrict__ x, int B, int H,
                                                        int W) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(B) * H * W * 2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx & 1);   // dx 0..3 | dx 4..6 + pad
    const long long pix = idx >> 1;
    const int xw = static_cast<int>(pix % W);
    const float2* row = reinterpret_cast<const float2*>(flow) + (pix - xw);
    float2 f[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int dx = g * 4 + j, ix = xw + dx - 3;
      f[j] = (dx < 7 && ix >= 0 && ix < W) ? __ldg(row + ix) : make_float2(0.0f, 0.0f);
    }
    __half2 h0 = __floats2half2_rn(f[0].x, f[0].y), h1 = __floats2half2_rn(f[1].x, f[1].y);
    __half2 h2 = __floats2half2_rn(f[2].x, f[2].y), h3 = __floats2half2_rn(f[3].x, f[3].y);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2);
    u.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(x + pix * 16 + g * 8) = u;
  }
}

// ------------------------------------------------------------------------------------------------
// Instance norm (per image, per channel over H*W) on NHWC fp16
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inorm_partial_kernel(const __half* __restrict__ x, long long pitch, int HW,
                                                            int C, int parts, float* __restrict__ scratch) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[256 * 16];
  const int b = blockIdx.y, part = blockIdx.x;
  const int groups = C >> 3;
  const int lanes = 256 / groups;
  const int g = threadIdx.x % groups, ln = threadIdx.x / groups;
  const int per = (HW + parts - 1) / parts;
  const int p0 = part * per, p1 = min(HW, p0 + per);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ss[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (ln < lanes) {
#pragma unroll 4
    for (int pidx = p0 + ln; pidx < p1; pidx += lanes) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + (static_cast<long long>(b) * HW + pidx) * pitch + g * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        s[2 * j] += f.x; s[2 * j + 1] += f.y;
        ss[2 * j] += f.x * f.x; ss[2 * j + 1] += f.y * f.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[threadIdx.x * 16 + j] = s[j];
    red[threadIdx.x * 16 + 8 + j] = ss[j];
  }
  __syncthreads();
  // thread t < C*2 reduces one (channel, moment) over the pixel lanes in a fixed order
  for (int t = threadIdx.x; t < C * 2; t += 256) {
    const int c = t >> 1, mom = t & 1;
    const int gg = c >> 3, j = c & 7;
    float acc = 0.0f;
    for (int l2 = 0; l2 < lanes; ++l2) acc += red[(l2 * groups + gg) * 16 + mom * 8 + j];
    scratch[((static_cast<long long>(b) * parts + part) * C + c) * 2 + mom] = acc;
  }
}

// one warp per (image, channel): lanes stride over the partial sums in a fixed order, fixed shuffle tree
__global__ void __launch_bounds__(256) inorm_finalize_kernel(const float* __restrict__ scratch, int parts, int C, int HW,
                                                             int BC, float* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  const int wid = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wid >= BC) return;
  const int b = wid / C, c = wid - b * C;
  double s = 0.0, ss = 0.0;
  for (int p = lane; p < parts; p += 32) {
    const float2 v = *reinterpret_cast<const float2*>(scratch + ((static_cast<long long>(b) * parts + p) * C + c) * 2);
    s += v.x;
    ss += v.y;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (lane == 0) {
    const double mean = s / HW;
    double var = ss / HW - mean * mean;   // biased variance, as nn.InstanceNorm2d
    if (var < 0.0) var = 0.0;
    stats[static_cast<long long>(wid) * 2] = static_cast<float>(mean);
    stats[static_cast<long long>(wid) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + 1e-5));
  }
}

// One thread handles 4 (pixel, 8-channel group) items 256 apart and issues all of its loads before the first use:
// with one 16-byte load per thread the kernel ran at 3.6 TB/s (latency-bound: 32 KB in flight per SM).
// grid = (ceil(HW * C/8 / 1024), batch); 32-bit index math only.
constexpr int kApplyU = 4;
__global__ void __launch_bounds__(256) inorm_apply_kernel(const __half* __restrict__ x, long long pitch, const float* __restrict__ stats,
                                                          const __half* __restrict__ resid, long long rpitch, __half* __restrict__ y,
                                                          long long ypitch, int B, int HW, int C, int relu) {
  pdl_launch_dependents();
  pdl_wait();
  const unsigned groups = static_cast<unsigned>(C) >> 3;
  const unsigned per_image = static_cast<unsigned>(HW) * groups;
  const unsigned base = blockIdx.x * (256u * kApplyU) + threadIdx.x;
  const int b = blockIdx.y;
  uint4 u[kApplyU], ur[kApplyU];
  long long pix[kApplyU];
  unsigned g[kApplyU];
  bool ok[kApplyU];
#pragma unroll
  for (int k = 0; k < kApplyU; ++k) {
    const unsigned idx = base + k * 256u;
    ok[k] = idx < per_image;
    g[k] = idx % groups;
    pix[k] = static_cast<long long>(b) * HW + idx / groups;
    u[k] = ur[k] = make_uint4(0, 0, 0, 0);
    if (ok[k]) {
      u[k] = *reinterpret_cast<const uint4*>(x + pix[k] * pitch + g[k] * 8);
      if (resid) ur[k] = *reinterpret_cast<const uint4*>(resid + pix[k] * rpitch + g[k] * 8);
    }
  }
#pragma unroll
  for (int k = 0; k < kApplyU; ++k) {
    if (!ok[k]) continue;
    const __half2* h = reinterpret_cast<const __half2*>(&u[k]);
    const float4* st = reinterpret_cast<const float4*>(stats + (static_cast<long long>(b) * C + g[k] * 8) * 2);
    float v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      const float4 ms = __ldg(st + j);   // (mean0, rstd0, mean1, rstd1)
      v[2 * j] = (f.x - ms.x) * ms.y;
      v[2 * j + 1] = (f.y - ms.z) * ms.w;
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
    }
    if (resid) {
      const __half2* hr = reinterpret_cast<const __half2*>(&ur[k]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(hr[j]);
        v[2 * j] = fmaxf(f.x + v[2 * j], 0.0f);
        v[2 * j + 1] = fmaxf(f.y + v[2 * j + 1], 0.0f);
      }
    }
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) ho[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(y + pix[k] * ypitch + g[k] * 8) = o;
  }
}

// Specialised apply for the encoder widths (C = 64 / 96 / 128): the generic kernel above is instruction-bound
// (runtime division, four statistics loads and ~110 instructions per 16 bytes: 3.6 TB/s).  Here the channel group of a
// thread is fixed (block size is a multiple of C/8), so mean / rstd are loaded once and reused for 8 pixels, the
// divisions are by compile-time constants, and all 8 loads are issued before the first use.
template <int C>
__global__ void __launch_bounds__(C == 96 ? 192 : 256) inorm_apply_fixed_kernel(const __half* __restrict__ x, long long pitch,
                                                                              const float* __restrict__ stats,
                                                                              const __half* __restrict__ resid, long long rpitch,
                                                                              __half* __restrict__ y, long long ypitch, int HW,
                                                                              int relu) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int G = C / 8, T = (C == 96 ? 192 : 256), LANES = T / G, U = 8;
  const int g = threadIdx.x % G, pl = threadIdx.x / G, b = blockIdx.y;
  float mean[8], rstd[8];
  {
    const float4* st = reinterpret_cast<const float4*>(stats + (static_cast<long long>(b) * C + g * 8) * 2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 ms = __ldg(st + j);
      mean[2 * j] = ms.x; rstd[2 * j] = ms.y; mean[2 * j + 1] = ms.z; rstd[2 * j + 1] = ms.w;
    }
  }
  const int p0 = blockIdx.x * (LANES * U) + pl;
  const long long img = static_cast<long long>(b) * HW;
  uint4 u[U], ur[U];
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const int p = p0 + k * LANES;
    u[k] = ur[k] = make_uint4(0, 0, 0, 0);
    if (p < HW) {
      u[k] = *reinterpret_cast<const uint4*>(x + (img + p) * pitch + g * 8);
      if (resid) ur[k] = *reinterpret_cast<const uint4*>(resid + (img + p) * rpitch + g * 8);
    }
  }
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const int p = p0 + k * LANES;
    if (p >= HW) continue;
    const __half2* h = reinterpret_cast<const __half2*>(&u[k]);
    const __half2* hr = reinterpret_cast<const __half2*>(&ur[k]);
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      float v0 = (f.x - mean[2 * j]) * rstd[2 * j], v1 = (f.y - mean[2 * j + 1]) * rstd[2 * j + 1];
      if (relu) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
      if (resid) {
        const float2 r = __half22float2(hr[j]);
        v0 = fmaxf(r.x + v0, 0.0f);
        v1 = fmaxf(r.y + v1, 0.0f);
      }
      ho[j] = __floats2half2_rn(v0, v1);
    }
    *reinterpret_cast<uint4*>(y + (img + p) * ypitch + g * 8) = o;
  }
}

template <int C>
static cudaError_t launch_inorm_apply_fixed(const __half* x, long long pitch, const float* stats, const __half* resid, long long rpitch,
                                     __half* y, long long ypitch, int batch, int hw, int relu, cudaStream_t s) {
  constexpr int T = (C == 96 ? 192 : 256), LANES = T / (C / 8), U = 8;
  return launch_pdl(inorm_apply_fixed_kernel<C>, dim3((hw + LANES * U - 1) / (LANES * U), batch), dim3(T), 0, s, x, pitch, stats, resid, rpitch, y, ypitch,
                    hw, relu);
}

// flow_head.conv2 as "1x1 conv + gather": the tensor-core kernel evaluates all nine taps on the UNSHIFTED pixel,
//   d[q, tap*2 + co] = sum_c w2[co, c, tap] * x[q, c]        (a 256 -> 18 1x1 convolution: x is read once, not 9x),
// and this kernel sums the shifted contributions  delta[p, co] = bias[co] + sum_tap d[p + off(tap), tap*2 + co]
// (zero for neighbours outside the image = the conv zero padding) and applies network.py:116.  The 3x3 conv with two
// real output channels in a 32-wide MMA tile re-read its A tile for every tap and took 4x longer than both together.
__global__ void __launch_bounds__(256) flow_head_gather_kernel(const float* __restrict__ d, long long pitch,
                                                               const float* __restrict__ bias, float* __restrict__ coords1,
                                                               float* __restrict__ flow, int B, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  const long long npix = static_cast<long long>(B) * H * W;
  const long long pix = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (pix >= npix) return;
  const int xw = static_cast<int>(pix % W);
  const int yh = static_cast<int>((pix / W) % H);
  float a0 = __ldg(bias), a1 = __ldg(bias + 1);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int iy = yh + dy, ix = xw + dx;
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
    const float2 v = __ldg(reinterpret_cast<const float2*>(d + (pix + dy * W + dx) * pitch + tap * 2));
    a0 += v.x;
    a1 += v.y;
  }
  const float2 c = *reinterpret_cast<const float2*>(coords1 + pix * 2);
  const float cxn = c.x + a0, cyn = c.y + a1;
  *reinterpret_cast<float2*>(coords1 + pix * 2) = make_float2(cxn, cyn);
  *reinterpret_cast<float2*>(flow + pix * 2) = make_float2(cxn - static_cast<float>(xw), cyn - static_cast<float>(yh));
}

// ------------------------------------------------------------------------------------------------
// Convex upsampling: block = 4 low-res pixels (x) x 8 x 8 sub-pixels
// ------------------------------------------------------------------------------------------------
template <typename MaskT>
__global__ void __launch_bounds__(256) convex_upsample_kernel(const MaskT* __restrict__ mask, long long mpitch,
                                                              const float* __restrict__ flow,
                                                              float* __restrict__ up, float* __restrict__ lo, int B,
                                                              int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  const int j = threadIdx.x & 7, px = threadIdx.x >> 3;   // blockDim.x = 32
  const int i = threadIdx.y;                              // blockDim.y = 8
  const int x = blockIdx.x * 4 + px, y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const long long pix = (static_cast<long long>(b) * H + y) * W + x;
  const MaskT* mrow = mask + pix * mpitch + i * 8 + j;
  float mk[9], mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    mk[k] = static_cast<float>(mrow[k * 64]);
    mx = fmaxf(mx, mk[k]);
  }
  float sum = 0.0f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    mk[k] = expf(mk[k] - mx);
    sum += mk[k];
  }
  float ox = 0.0f, oy = 0.0f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const float2 f = *reinterpret_cast<const float2*>(flow + ((static_cast<long long>(b) * H + yy) * W + xx) * 2);
      const float wgt = mk[k] / sum;
      ox += wgt * (8.0f * f.x);
      oy += wgt * (8.0f * f.y);
    }
  }
  const long long HW8 = static_cast<long long>(H) * 8 * W * 8;
  const long long o = (static_cast<long long>(y) * 8 + i) * (W * 8) + x * 8 + j;
  up[(static_cast<long long>(b) * 2) * HW8 + o] = ox;
  up[(static_cast<long long>(b) * 2 + 1) * HW8 + o] = oy;
  if (lo && i == 0 && j == 0) {
    const float2 f = *reinterpret_cast<const float2*>(flow + pix * 2);
    lo[((static_cast<long long>(b) * 2) * H + y) * W + x] = f.x;
    lo[((static_cast<long long>(b) * 2 + 1) * H + y) * W + x] = f.y;
  }
}

__global__ void coords_init_kernel(float* __restrict__ coords1, float* __restrict__ flow,
                                   const float* __restrict__ init, int B, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  const long long n = static_cast<long long>(B) * H * W;
  for (long long pix = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; pix < n;
       pix += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(pix % W);
    const long long t = pix / W;
    const int y = static_cast<int>(t % H);
    const long long b = t / H;
    float fx = 0.0f, fy = 0.0f;
    if (init) {
      fx = init[((b * 2) * H + y) * W + x];
      fy = init[((b * 2 + 1) * H + y) * W + x];
    }
    coords1[pix * 2] = static_cast<float>(x) + fx;
    coords1[pix * 2 + 1] = static_cast<float>(y) + fy;
    flow[pix * 2] = fx;
    flow[pix * 2 + 1] = fy;
  }
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;   // a few waves of the 148 SMs; kernels are grid-stride
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

// ------------------------------------------------------------------------------------------------
// Caller-side resize of NeuralSLAM.__call__ (neural_slam.py:197-199: TF.resize = bilinear with antialias) as one kernel:
// the separable triangle filter of ATen's upsample_bilinear2d_aa (support = max(scale, 1) input pixels around
// center = scale * (o + 0.5), weights normalised to 1; rows first, then the column of row results), one thread per
// output pixel.  PyTorch's generic kernel takes 1.4 ms for the 55 frames of a batch (2.5% of the step); this one is
// bound by the 0.6 GB it moves.
// ------------------------------------------------------------------------------------------------
constexpr int kAaRows = 8;         // output rows per block: the column weights are computed once per thread and reused

template <int TAPS>
struct AaSpan { int lo, size; float w[TAPS]; };

template <int TAPS>
__device__ __forceinline__ AaSpan<TAPS> aa_span(int o, int in_size, float scale) {
  AaSpan<TAPS> s;
  const float support = scale >= 1.0f ? scale : 1.0f;
  const float center = scale * (static_cast<float>(o) + 0.5f);
  s.lo = max(static_cast<int>(center - support + 0.5f), 0);
  s.size = min(min(static_cast<int>(center + support + 0.5f), in_size) - s.lo, TAPS);
  const float invscale = scale >= 1.0f ? static_cast<float>(1.0 / static_cast<double>(scale)) : 1.0f;
  const float lo_m_center = static_cast<float>(s.lo) - center;
  float total = 0.0f;
#pragma unroll
  for (int j = 0; j < TAPS; ++j) {
    float w = 0.0f;
    if (j < s.size) {
      const float x = fabsf((static_cast<float>(j) + lo_m_center + 0.5f) * invscale);
      w = x < 1.0f ? 1.0f - x : 0.0f;
    }
    s.w[j] = w;
    total += w;
  }
  if (total != 0.0f) {
#pragma unroll
    for (int j = 0; j < TAPS; ++j) s.w[j] /= total;
  }
  return s;
}

// TAPS = 2 * ceil(support) + 1 (5 up to 2x down-scaling, 11 up to 5x); SAME_H: the height is unchanged and the row filter is
// the identity (its two taps are 1 and 0: the result is the row value itself, bit for bit)
template <typename T, int TAPS, bool SAME_H>
__global__ void __launch_bounds__(128) resize_aa_kernel(const T* __restrict__ src, float* __restrict__ dst, int ih, int iw, int oh, int ow,
                                                        float hscale, float wscale) {
  __shared__ AaSpan<TAPS> sy_s[kAaRows];
  const int x = blockIdx.x * 128 + threadIdx.x, y0 = blockIdx.y * kAaRows;
  if constexpr (!SAME_H) {
    if (threadIdx.x < kAaRows && y0 + threadIdx.x < oh) sy_s[threadIdx.x] = aa_span<TAPS>(y0 + threadIdx.x, ih, hscale);
    __syncthreads();
  }
  if (x >= ow) return;
  const AaSpan<TAPS> sx = aa_span<TAPS>(x, iw, wscale);
  const T* plane = src + static_cast<long long>(blockIdx.z) * ih * iw + sx.lo;
  float* out = dst + static_cast<long long>(blockIdx.z) * oh * ow + x;
  auto row_value = [&](int r) {
    const T* row = plane + static_cast<long long>(r) * iw;
    float t = 0.0f;
#pragma unroll
    for (int j = 0; j < TAPS; ++j)
      if (j < sx.size) t = fmaf(static_cast<float>(row[j]), sx.w[j], t);
    return t;
  };
#pragma unroll 1
  for (int k = 0; k < kAaRows && y0 + k < oh; ++k) {
    if constexpr (SAME_H) {
      out[static_cast<long long>(y0 + k) * ow] = row_value(y0 + k);
    } else {
      const AaSpan<TAPS>& sy = sy_s[k];
      float acc = 0.0f;
      for (int r = 0; r < sy.size; ++r) acc = fmaf(row_value(sy.lo + r), sy.w[r], acc);
      out[static_cast<long long>(y0 + k) * ow] = acc;
    }
  }
}

}  // namespace atdn

using namespace atdn;

extern "C" int atdn_corr_lookup(const void* const lvl[4], const int32_t lvl_pitch[4], int32_t half_levels, const float* coords,
                                void* out16, int64_t out_pitch, float* out32, int32_t batch, int32_t h8, int32_t w8, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(half_levels == 0 || half_levels == 4, ATDN_ERR_UNSUP, "atdn_corr_lookup: half_levels must be 0 (fp32 pyramid) or 4 (fp16 pyramid), got %d", half_levels);
  ATDN_REQUIRE(lvl && lvl_pitch && coords && (out16 || out32), ATDN_ERR_ARG, "atdn_corr_lookup: null argument");
  ATDN_REQUIRE(h8 >= 16 && w8 >= 16, ATDN_ERR_ARG, "atdn_corr_lookup: grid %dx%d is smaller than 16x16 (level 3 would be < 2x2)", h8, w8);
  ATDN_REQUIRE(!out16 || out_pitch >= 324, ATDN_ERR_ARG, "atdn_corr_lookup: out_pitch < 324");
  const long long nq = static_cast<long long>(batch) * h8 * w8;
  ATDN_REQUIRE(!out16 || (out_pitch % 8 == 0 && aligned16(out16)), ATDN_ERR_ALIGN, "atdn_corr_lookup: out16 must be 16-byte aligned with a pitch that is a multiple of 8");
  ATDN_REQUIRE(!out32 || aligned16(out32), ATDN_ERR_ALIGN, "atdn_corr_lookup: out32 must be 16-byte aligned");
  int w = w8;
  for (int l = 0; l < 4; ++l) {
    ATDN_REQUIRE(lvl[l] != nullptr && (half_levels || lvl_pitch[l] >= w), ATDN_ERR_ARG, "atdn_corr_lookup: level %d", l);
    ATDN_REQUIRE(lvl_pitch[l] % 4 == 0 && aligned16(lvl[l]), ATDN_ERR_ALIGN, "atdn_corr_lookup: level %d must be 16-byte aligned with a pitch that is a multiple of 4", l);
    w /= 2;
  }
  const unsigned grid = static_cast<unsigned>((nq + kLkWarps - 1) / kLkWarps);
  if (half_levels) {
    // fp16 STRIP layout written by atdn_corr_pyramid (corr_lookup_strip.cuh)
    lks::Params p;
    const int tiles_w = (w8 + 31) / 32, tiles_h = (h8 + 7) / 8;
    const int tiles = tiles_w * tiles_h, tiles_w3 = (tiles_w + 1) & ~1;
    const int expect[4] = {256, 64, 16, 4};
    for (int l = 0; l < 4; ++l) {
      ATDN_REQUIRE(lvl_pitch[l] == expect[l], ATDN_ERR_ARG, "atdn_corr_lookup: strip-layout level %d has %d elements per tile, expected %d", l, lvl_pitch[l], expect[l]);
      p.lvl[l] = static_cast<const __half*>(lvl[l]);
      p.qstride[l] = l < 3 ? static_cast<long long>(tiles) * expect[l] : static_cast<long long>(tiles_h) * tiles_w3 * 4;
      p.rowmul[l] = l < 3 ? tiles_w * expect[l] : tiles_w3 * 4;
      p.hp[l] = l < 3 ? tiles_h * (8 >> l) : tiles_h;
      p.xs[l] = l < 3 ? (tiles_w * 4) >> l : tiles_w3 / 2;
    }
    p.h0 = h8;
    p.w0 = w8;
    p.coords = coords;
    p.out16 = static_cast<__half*>(out16);
    p.out32 = out32;
    p.out_pitch = out_pitch;
    p.nq = nq;
    static DeviceOnce configured;
    if (configured.pending()) {
      ATDN_CUDA(cudaFuncSetAttribute(lks::corr_lookup_strip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lks::kSmemBytes));
      ATDN_CUDA(cudaFuncSetAttribute(lks::corr_lookup_strip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lks::kSmemBytes));
      configured.done();
    }
    const long long per_cta = static_cast<long long>(lks::kWarps) * lks::kQueriesPerWarp;
    const unsigned ctas = static_cast<unsigned>((nq + per_cta - 1) / per_cta);
    if (out32) ATDN_CUDA(launch_pdl(lks::corr_lookup_strip_kernel<true>, dim3(ctas), dim3(lks::kWarps * 32), lks::kSmemBytes, static_cast<cudaStream_t>(stream), p));
    else ATDN_CUDA(launch_pdl(lks::corr_lookup_strip_kernel<false>, dim3(ctas), dim3(lks::kWarps * 32), lks::kSmemBytes, static_cast<cudaStream_t>(stream), p));
  } else {
    LookupParams p;
    int h = h8;
    w = w8;
    for (int l = 0; l < 4; ++l) {
      p.lvl[l] = static_cast<const float*>(lvl[l]);
      p.pitch[l] = lvl_pitch[l];
      p.h[l] = h;
      p.w[l] = w;
      h /= 2;
      w /= 2;
    }
    corr_lookup_kernel<<<grid, kLkWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(p, coords, static_cast<__half*>(out16), out_pitch, out32, nq);
  }
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_resize_aa(const void* src, int32_t src_is_u8, float* dst, int32_t planes, int32_t in_h, int32_t in_w, int32_t out_h,
                              int32_t out_w, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(src && dst && planes > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, ATDN_ERR_ARG, "atdn_resize_aa: null / empty argument");
  ATDN_REQUIRE(planes <= 65535, ATDN_ERR_UNSUP, "atdn_resize_aa: %d planes exceed the grid", planes);
  const float hscale = static_cast<float>(in_h) / out_h, wscale = static_cast<float>(in_w) / out_w;
  ATDN_REQUIRE(hscale <= 5.0f && wscale <= 5.0f, ATDN_ERR_UNSUP, "atdn_resize_aa: down-scaling by more than 5x (%d x %d -> %d x %d)", in_h, in_w, out_h, out_w);
  const dim3 grid((out_w + 127) / 128, (out_h + kAaRows - 1) / kAaRows, planes);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool small = hscale <= 2.0f && wscale <= 2.0f, same_h = in_h == out_h;
#define ATDN_RESIZE(T, TAPS, SAME)                                                                                                \
  resize_aa_kernel<T, TAPS, SAME><<<grid, 128, 0, s>>>(static_cast<const T*>(src), dst, in_h, in_w, out_h, out_w, hscale, wscale)
  if (src_is_u8) {
    if (small && same_h) ATDN_RESIZE(uint8_t, 5, true);
    else if (small) ATDN_RESIZE(uint8_t, 5, false);
    else ATDN_RESIZE(uint8_t, 11, false);
  } else {
    if (small && same_h) ATDN_RESIZE(float, 5, true);
    else if (small) ATDN_RESIZE(float, 5, false);
    else ATDN_RESIZE(float, 11, false);
  }
#undef ATDN_RESIZE
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_stem_pack(const float* image, void* x16, int32_t batch, int32_t h, int32_t w, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(image && x16 && aligned16(x16) && (reinterpret_cast<uintptr_t>(image) & 7u) == 0 && h % 2 == 0 && w % 2 == 0 && w >= 8,
               ATDN_ERR_ARG, "atdn_stem_pack: bad arguments (even h, w >= 8; 8-byte aligned image, 16-byte aligned output)");
  const int oh = h / 2, ow = w / 2;
  ATDN_REQUIRE(oh <= 65535 && batch <= 65535, ATDN_ERR_UNSUP, "atdn_stem_pack: image too large");
  ATDN_CUDA(launch_pdl(stem_pack_kernel, dim3((ow * 6 + 255) / 256, oh, batch), dim3(256), 0, static_cast<cudaStream_t>(stream), image,
                       static_cast<__half*>(x16), batch, h, w, oh, ow));
  return 0;
}

extern "C" int atdn_flow_pack(const float* flow, void* x16, int32_t batch, int32_t h8, int32_t w8, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(flow && x16 && aligned16(x16) && (reinterpret_cast<uintptr_t>(flow) & 7u) == 0, ATDN_ERR_ARG, "atdn_flow_pack: bad arguments");
  const long long total = static_cast<long long>(batch) * h8 * w8 * 2;
  ATDN_CUDA(launch_pdl(flow_pack_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), flow, static_cast<__half*>(x16),
                       batch, h8, w8));
  return 0;
}

extern "C" int atdn_inorm_stats(const void* x16, int64_t pitch, int32_t batch, int32_t hw, int32_t c, float* scratch,
                                int32_t parts, float* stats, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(x16 && scratch && stats && parts >= 1, ATDN_ERR_ARG, "atdn_inorm_stats: null argument");
  ATDN_REQUIRE(c % 8 == 0 && c >= 8 && c <= 256 && pitch % 8 == 0 && aligned16(x16), ATDN_ERR_ALIGN, "atdn_inorm_stats: C=%d pitch=%lld", c, (long long)pitch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ATDN_CUDA(launch_pdl(inorm_partial_kernel, dim3(parts, batch), dim3(256), 0, s, static_cast<const __half*>(x16), pitch, hw, c, parts, scratch));
  ATDN_CUDA(launch_pdl(inorm_finalize_kernel, dim3((batch * c + 7) / 8), dim3(256), 0, s, scratch, parts, c, hw, batch * c, stats));
  return 0;
}

extern "C" int atdn_inorm_finalize(const float* scratch, int32_t parts, int32_t batch, int32_t c, int32_t hw, float* stats, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(scratch && stats && parts >= 1 && batch >= 1 && c >= 1 && hw >= 1, ATDN_ERR_ARG, "atdn_inorm_finalize: bad arguments");
  ATDN_CUDA(launch_pdl(inorm_finalize_kernel, dim3((batch * c + 7) / 8), dim3(256), 0, static_cast<cudaStream_t>(stream), scratch, parts, c, hw,
                       batch * c, stats));
  return 0;
}

extern "C" int atdn_inorm_apply(const void* x16, int64_t pitch, const float* stats, const void* resid16, int64_t resid_pitch,
                                void* y16, int64_t y_pitch, int32_t batch, int32_t hw, int32_t c, int32_t relu, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(x16 && stats && y16, ATDN_ERR_ARG, "atdn_inorm_apply: null argument");
  ATDN_REQUIRE(c % 8 == 0 && pitch % 8 == 0 && y_pitch % 8 == 0 && (!resid16 || resid_pitch % 8 == 0) && aligned16(x16) && aligned16(y16) && aligned16(stats),
               ATDN_ERR_ALIGN, "atdn_inorm_apply: alignment");
  ATDN_REQUIRE(static_cast<long long>(hw) * (c / 8) < (1LL << 31) && batch <= 65535, ATDN_ERR_UNSUP, "atdn_inorm_apply: image too large");
  if (c == 64 || c == 96 || c == 128) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const __half *xx = static_cast<const __half*>(x16), *rr = static_cast<const __half*>(resid16);
    __half* yy = static_cast<__half*>(y16);
    if (c == 64) ATDN_CUDA(launch_inorm_apply_fixed<64>(xx, pitch, stats, rr, resid_pitch, yy, y_pitch, batch, hw, relu, st));
    else if (c == 96) ATDN_CUDA(launch_inorm_apply_fixed<96>(xx, pitch, stats, rr, resid_pitch, yy, y_pitch, batch, hw, relu, st));
    else ATDN_CUDA(launch_inorm_apply_fixed<128>(xx, pitch, stats, rr, resid_pitch, yy, y_pitch, batch, hw, relu, st));
    return 0;
  }
  const unsigned per_image = static_cast<unsigned>(hw) * static_cast<unsigned>(c / 8);
  inorm_apply_kernel<<<dim3((per_image + 256 * kApplyU - 1) / (256 * kApplyU), batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x16), pitch, stats, static_cast<const __half*>(resid16), resid_pitch,
      static_cast<__half*>(y16), y_pitch, batch, hw, c, relu);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_flow_head_gather(const float* d32, int64_t pitch, const float* bias, float* coords1, float* flow,
                                     int32_t batch, int32_t h8, int32_t w8, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(d32 && bias && coords1 && flow, ATDN_ERR_ARG, "atdn_flow_head_gather: null argument");
  ATDN_REQUIRE(pitch >= 18 && pitch % 2 == 0 && (reinterpret_cast<uintptr_t>(d32) & 7u) == 0, ATDN_ERR_ALIGN, "atdn_flow_head_gather: d32 must be 8-byte aligned with an even pitch >= 18");
  const long long npix = static_cast<long long>(batch) * h8 * w8;
  ATDN_CUDA(launch_pdl(flow_head_gather_kernel, dim3(static_cast<unsigned>((npix + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                       d32, static_cast<long long>(pitch), bias, coords1, flow, batch, h8, w8));
  return 0;
}

extern "C" int atdn_convex_upsample(const void* mask, int32_t mask_is_half, int64_t mask_pitch, const float* flow, float* flow_up,
                                    float* flow_lo, int32_t batch, int32_t h8, int32_t w8, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(mask && flow && flow_up && mask_pitch >= 576, ATDN_ERR_ARG, "atdn_convex_upsample: bad arguments");
  const dim3 grid((w8 + 3) / 4, h8, batch), block(32, 8);
  if (mask_is_half)
    ATDN_CUDA(launch_pdl(convex_upsample_kernel<__half>, grid, block, 0, static_cast<cudaStream_t>(stream), static_cast<const __half*>(mask), mask_pitch, flow,
                         flow_up, flow_lo, batch, h8, w8));
  else
    ATDN_CUDA(launch_pdl(convex_upsample_kernel<float>, grid, block, 0, static_cast<cudaStream_t>(stream), static_cast<const float*>(mask), mask_pitch, flow,
                         flow_up, flow_lo, batch, h8, w8));
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_coords_init(float* coords1, float* flow, const float* flow_init, int32_t batch, int32_t h8, int32_t w8,
                                void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(coords1 && flow, ATDN_ERR_ARG, "atdn_coords_init: null argument");
  const long long n = static_cast<long long>(batch) * h8 * w8;
  ATDN_CUDA(launch_pdl(coords_init_kernel, dim3(grid_for(n, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), coords1, flow, flow_init, batch, h8, w8));
  ATDN_CUDA(cudaGetLastError());
  return 0;
}
