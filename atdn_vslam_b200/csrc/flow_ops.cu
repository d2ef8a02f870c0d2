// Gather / element-wise kernels of the GMA flow path (HBM- or latency-bound, CUDA cores).
// See include/atdn_b200.h for the reference call sites each entry point replaces.
#include <math.h>

#include <cuda_fp16.h>

#include "common.h"

namespace atdn {

// ------------------------------------------------------------------------------------------------
// Correlation lookup: one CTA (4 warps) per query pixel, warp l samples pyramid level l.
// Each of the 81 taps of a level is an independent bilinear sample; the 10x10 texel footprint of a
// level stays in L1 across the 3 rounds of a warp.  Coordinates follow the reference's round trip
// through normalised grid coordinates (utils.py:63-64 then ATen's align_corners un-normalisation),
// with explicit _rn intrinsics so that nvcc cannot contract the sequence into FMAs.
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

__global__ void __launch_bounds__(128) corr_lookup_kernel(LookupParams p, const float* __restrict__ coords,
                                                          __half* __restrict__ out16, long long out_pitch,
                                                          float* __restrict__ out32) {
  const long long q = blockIdx.x;
  const int l = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float cx = coords[q * 2], cy = coords[q * 2 + 1];
  const float inv = 1.0f / static_cast<float>(1 << l);
  const int H = p.h[l], W = p.w[l], pitch = p.pitch[l];
  const float* __restrict__ base = p.lvl[l] + q * static_cast<long long>(H) * pitch;
  const float xl = cx * inv, yl = cy * inv;   // exact: division by a power of two
  for (int t = lane; t < 81; t += 32) {
    const int a = t / 9, b = t - a * 9;       // a offsets x, b offsets y (corr.py:40-46)
    const float x = grid_round_trip(__fadd_rn(xl, static_cast<float>(a - 4)), W);
    const float y = grid_round_trip(__fadd_rn(yl, static_cast<float>(b - 4)), H);
    const float xf = floorf(x), yf = floorf(y);
    const float fx = x - xf, fy = y - yf;
    // clamp before the int conversion so that far-away coordinates cannot overflow
    const int x0 = static_cast<int>(fminf(fmaxf(xf, -2.0f), static_cast<float>(W)));
    const int y0 = static_cast<int>(fminf(fmaxf(yf, -2.0f), static_cast<float>(H)));
    const bool xin0 = x0 >= 0 && x0 < W, xin1 = x0 + 1 >= 0 && x0 + 1 < W;
    const bool yin0 = y0 >= 0 && y0 < H, yin1 = y0 + 1 >= 0 && y0 + 1 < H;
    const float* r0 = base + static_cast<long long>(y0) * pitch + x0;
    const float* r1 = r0 + pitch;
    const float v00 = (xin0 && yin0) ? __ldg(r0) : 0.0f;
    const float v01 = (xin1 && yin0) ? __ldg(r0 + 1) : 0.0f;
    const float v10 = (xin0 && yin1) ? __ldg(r1) : 0.0f;
    const float v11 = (xin1 && yin1) ? __ldg(r1 + 1) : 0.0f;
    const float val = v00 * ((1.0f - fx) * (1.0f - fy)) + v01 * (fx * (1.0f - fy)) + v10 * ((1.0f - fx) * fy) +
                      v11 * (fx * fy);
    const int ch = l * 81 + t;
    if (out16) out16[q * out_pitch + ch] = __float2half_rn(val);
    if (out32) out32[q * 324 + ch] = val;
  }
}

// ------------------------------------------------------------------------------------------------
// im2col for the two 7x7 layers whose input has 3 / 2 channels (too thin for a 64-channel K chunk)
// ------------------------------------------------------------------------------------------------
__global__ void stem_im2col_kernel(const float* __restrict__ img, __half* __restrict__ rows, long long pitch, int B,
                                   int H, int W, int OH, int OW) {
  const long long total = static_cast<long long>(B) * OH * OW * pitch;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(idx % pitch);
    const long long pix = idx / pitch;
    float v = 0.0f;
    if (k < 147) {
      const int c = k % 3, tap = k / 3, dy = tap / 7, dx = tap - dy * 7;
      const int ox = static_cast<int>(pix % OW);
      const long long t = pix / OW;
      const int oy = static_cast<int>(t % OH), b = static_cast<int>(t / OH);
      const int iy = oy * 2 + dy - 3, ix = ox * 2 + dx - 3;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        const float x = __ldg(img + ((static_cast<long long>(b) * 3 + c) * H + iy) * W + ix);
        v = 2.0f * (x / 255.0f) - 1.0f;        // network.py:75-76
      }
    }
    rows[idx] = __float2half_rn(v);
  }
}

__global__ void flow_im2col_kernel(const float* __restrict__ flow, __half* __restrict__ rows, long long pitch, int B,
                                   int H, int W) {
  const long long total = static_cast<long long>(B) * H * W * pitch;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(idx % pitch);
    const long long pix = idx / pitch;
    float v = 0.0f;
    if (k < 98) {
      const int c = k & 1, tap = k >> 1, dy = tap / 7, dx = tap - dy * 7;
      const int x = static_cast<int>(pix % W);
      const long long t = pix / W;
      const int y = static_cast<int>(t % H), b = static_cast<int>(t / H);
      const int iy = y + dy - 3, ix = x + dx - 3;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W)
        v = __ldg(flow + ((static_cast<long long>(b) * H + iy) * W + ix) * 2 + c);
    }
    rows[idx] = __float2half_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------
// Instance norm (per image, per channel over H*W) on NHWC fp16
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inorm_partial_kernel(const __half* __restrict__ x, long long pitch, int HW,
                                                            int C, int parts, float* __restrict__ scratch) {
  __shared__ float red[256 * 16];
  const int b = blockIdx.y, part = blockIdx.x;
  const int groups = C >> 3;
  const int lanes = 256 / groups;
  const int g = threadIdx.x % groups, ln = threadIdx.x / groups;
  const int per = (HW + parts - 1) / parts;
  const int p0 = part * per, p1 = min(HW, p0 + per);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ss[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (ln < lanes) {
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

__global__ void inorm_finalize_kernel(const float* __restrict__ scratch, int parts, int C, int HW,
                                      float* __restrict__ stats) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s = 0.0, ss = 0.0;
    for (int p = 0; p < parts; ++p) {
      s += scratch[((static_cast<long long>(b) * parts + p) * C + c) * 2];
      ss += scratch[((static_cast<long long>(b) * parts + p) * C + c) * 2 + 1];
    }
    const double mean = s / HW;
    double var = ss / HW - mean * mean;   // biased variance, as nn.InstanceNorm2d
    if (var < 0.0) var = 0.0;
    stats[(static_cast<long long>(b) * C + c) * 2] = static_cast<float>(mean);
    stats[(static_cast<long long>(b) * C + c) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + 1e-5));
  }
}

__global__ void inorm_apply_kernel(const __half* __restrict__ x, long long pitch, const float* __restrict__ stats,
                                   const __half* __restrict__ resid, long long rpitch, __half* __restrict__ y,
                                   long long ypitch, int B, int HW, int C, int relu) {
  const int groups = C >> 3;
  const long long total = static_cast<long long>(B) * HW * groups;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    const long long pix = idx / groups;
    const int b = static_cast<int>(pix / HW);
    const uint4 u = *reinterpret_cast<const uint4*>(x + pix * pitch + g * 8);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    const float4* st = reinterpret_cast<const float4*>(stats + (static_cast<long long>(b) * C + g * 8) * 2);
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
      const uint4 ur = *reinterpret_cast<const uint4*>(resid + pix * rpitch + g * 8);
      const __half2* hr = reinterpret_cast<const __half2*>(&ur);
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
    *reinterpret_cast<uint4*>(y + pix * ypitch + g * 8) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// Row softmax of the attention logits: un-normalised fp16 probabilities + 1/sum
// ------------------------------------------------------------------------------------------------
constexpr int kSoftmaxThreads = 256;
constexpr int kSoftmaxMaxPerThread = 40;   // cols <= 10240

__global__ void __launch_bounds__(kSoftmaxThreads) softmax_rows_kernel(const float* __restrict__ s, long long spitch,
                                                                       __half* __restrict__ p, long long ppitch,
                                                                       float* __restrict__ inv_sum, int cols) {
  __shared__ float red[kSoftmaxThreads / 32];
  __shared__ float bcast;
  const long long row = blockIdx.x;
  const float* sr = s + row * spitch;
  float v[kSoftmaxMaxPerThread];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < kSoftmaxMaxPerThread; ++i) {
    const int c = threadIdx.x + i * kSoftmaxThreads;
    v[i] = c < cols ? sr[c] : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
    for (int i = 1; i < kSoftmaxThreads / 32; ++i) t = fmaxf(t, red[i]);
    bcast = t;
  }
  __syncthreads();
  m = bcast;
  float sum = 0.0f;
  __half* pr = p + row * ppitch;
#pragma unroll
  for (int i = 0; i < kSoftmaxMaxPerThread; ++i) {
    const int c = threadIdx.x + i * kSoftmaxThreads;
    if (c < cols) {
      const __half e = __float2half_rn(expf(v[i] - m));
      pr[c] = e;
      sum += __half2float(e);   // normalise by what the P.V GEMM will actually read
    }
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < kSoftmaxThreads / 32; ++i) t += red[i];
    inv_sum[row] = 1.0f / t;
  }
}

// ------------------------------------------------------------------------------------------------
// flow_head.conv2 (3x3, 256 -> 2) fused with the coordinate update: one warp per pixel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flow_head_update_kernel(const __half* __restrict__ x, long long pitch,
                                                               const float* __restrict__ w,
                                                               const float* __restrict__ bias,
                                                               float* __restrict__ coords1, float* __restrict__ flow,
                                                               int B, int H, int W) {
  __shared__ float2 ws[9 * 256];   // [tap][c] -> (w_out0, w_out1)
  for (int i = threadIdx.x; i < 9 * 256; i += blockDim.x) {
    const int tap = i / 256, c = i - tap * 256;
    ws[i] = make_float2(w[(0 * 256 + c) * 9 + tap], w[(1 * 256 + c) * 9 + tap]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pix = static_cast<long long>(blockIdx.x) * 8 + warp;
  const long long npix = static_cast<long long>(B) * H * W;
  if (pix >= npix) return;
  const int xw = static_cast<int>(pix % W);
  const long long t = pix / W;
  const int yh = static_cast<int>(t % H);
  const long long b = t / H;
  float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int iy = yh + tap / 3 - 1, ix = xw + tap % 3 - 1;
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
    const uint4 u = *reinterpret_cast<const uint4*>(x + ((b * H + iy) * W + ix) * pitch + lane * 8);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      const float2 w0 = ws[tap * 256 + lane * 8 + 2 * j], w1 = ws[tap * 256 + lane * 8 + 2 * j + 1];
      a0 += f.x * w0.x + f.y * w1.x;
      a1 += f.x * w0.y + f.y * w1.y;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if (lane == 0) {
    const float cxn = coords1[pix * 2] + (a0 + bias[0]);
    const float cyn = coords1[pix * 2 + 1] + (a1 + bias[1]);
    coords1[pix * 2] = cxn;
    coords1[pix * 2 + 1] = cyn;
    flow[pix * 2] = cxn - static_cast<float>(xw);
    flow[pix * 2 + 1] = cyn - static_cast<float>(yh);
  }
}

// ------------------------------------------------------------------------------------------------
// Convex upsampling: block = 4 low-res pixels (x) x 8 x 8 sub-pixels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) convex_upsample_kernel(const float* __restrict__ mask, long long mpitch,
                                                              const float* __restrict__ flow,
                                                              float* __restrict__ up, float* __restrict__ lo, int B,
                                                              int H, int W) {
  const int j = threadIdx.x & 7, px = threadIdx.x >> 3;   // blockDim.x = 32
  const int i = threadIdx.y;                              // blockDim.y = 8
  const int x = blockIdx.x * 4 + px, y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const long long pix = (static_cast<long long>(b) * H + y) * W + x;
  const float* mrow = mask + pix * mpitch + i * 8 + j;
  float mk[9], mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    mk[k] = mrow[k * 64];
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

}  // namespace atdn

using namespace atdn;

extern "C" int atdn_corr_lookup(const float* const lvl[4], const int32_t lvl_pitch[4], const float* coords, void* out16,
                                int64_t out_pitch, float* out32, int32_t batch, int32_t h8, int32_t w8, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(lvl && lvl_pitch && coords && (out16 || out32), ATDN_ERR_ARG, "atdn_corr_lookup: null argument");
  ATDN_REQUIRE(h8 >= 16 && w8 >= 16, ATDN_ERR_ARG, "atdn_corr_lookup: grid %dx%d is smaller than 16x16 (level 3 would be < 2x2)", h8, w8);
  ATDN_REQUIRE(!out16 || out_pitch >= 324, ATDN_ERR_ARG, "atdn_corr_lookup: out_pitch < 324");
  LookupParams p;
  int h = h8, w = w8;
  for (int l = 0; l < 4; ++l) {
    ATDN_REQUIRE(lvl[l] != nullptr && lvl_pitch[l] >= w, ATDN_ERR_ARG, "atdn_corr_lookup: level %d", l);
    p.lvl[l] = lvl[l];
    p.pitch[l] = lvl_pitch[l];
    p.h[l] = h;
    p.w[l] = w;
    h /= 2;
    w /= 2;
  }
  const long long nq = static_cast<long long>(batch) * h8 * w8;
  corr_lookup_kernel<<<static_cast<unsigned>(nq), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      p, coords, static_cast<__half*>(out16), out_pitch, out32);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_stem_im2col(const float* image, void* rows16, int64_t pitch, int32_t batch, int32_t h, int32_t w,
                                void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(image && rows16 && pitch >= 147 && h % 2 == 0 && w % 2 == 0, ATDN_ERR_ARG, "atdn_stem_im2col: bad arguments");
  const int oh = h / 2, ow = w / 2;
  const long long total = static_cast<long long>(batch) * oh * ow * pitch;
  stem_im2col_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      image, static_cast<__half*>(rows16), pitch, batch, h, w, oh, ow);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_flow_im2col(const float* flow, void* rows16, int64_t pitch, int32_t batch, int32_t h8, int32_t w8,
                                void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(flow && rows16 && pitch >= 98, ATDN_ERR_ARG, "atdn_flow_im2col: bad arguments");
  const long long total = static_cast<long long>(batch) * h8 * w8 * pitch;
  flow_im2col_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      flow, static_cast<__half*>(rows16), pitch, batch, h8, w8);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_inorm_stats(const void* x16, int64_t pitch, int32_t batch, int32_t hw, int32_t c, float* scratch,
                                int32_t parts, float* stats, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(x16 && scratch && stats && parts >= 1, ATDN_ERR_ARG, "atdn_inorm_stats: null argument");
  ATDN_REQUIRE(c % 8 == 0 && c >= 8 && c <= 256 && pitch % 8 == 0 && aligned16(x16), ATDN_ERR_ALIGN, "atdn_inorm_stats: C=%d pitch=%lld", c, (long long)pitch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  inorm_partial_kernel<<<dim3(parts, batch), 256, 0, s>>>(static_cast<const __half*>(x16), pitch, hw, c, parts, scratch);
  ATDN_CUDA(cudaGetLastError());
  inorm_finalize_kernel<<<batch, 128, 0, s>>>(scratch, parts, c, hw, stats);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_inorm_apply(const void* x16, int64_t pitch, const float* stats, const void* resid16, int64_t resid_pitch,
                                void* y16, int64_t y_pitch, int32_t batch, int32_t hw, int32_t c, int32_t relu, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(x16 && stats && y16, ATDN_ERR_ARG, "atdn_inorm_apply: null argument");
  ATDN_REQUIRE(c % 8 == 0 && pitch % 8 == 0 && y_pitch % 8 == 0 && (!resid16 || resid_pitch % 8 == 0) && aligned16(x16) && aligned16(y16) && aligned16(stats),
               ATDN_ERR_ALIGN, "atdn_inorm_apply: alignment");
  const long long total = static_cast<long long>(batch) * hw * (c / 8);
  inorm_apply_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x16), pitch, stats, static_cast<const __half*>(resid16), resid_pitch,
      static_cast<__half*>(y16), y_pitch, batch, hw, c, relu);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_softmax_rows(const float* s32, int64_t s_pitch, void* p16, int64_t p_pitch, float* inv_sum, int64_t rows,
                                 int32_t cols, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(s32 && p16 && inv_sum && rows > 0, ATDN_ERR_ARG, "atdn_softmax_rows: null argument");
  ATDN_REQUIRE(cols > 0 && cols <= kSoftmaxThreads * kSoftmaxMaxPerThread, ATDN_ERR_UNSUP, "atdn_softmax_rows: cols=%d > %d", cols, kSoftmaxThreads * kSoftmaxMaxPerThread);
  softmax_rows_kernel<<<static_cast<unsigned>(rows), kSoftmaxThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      s32, s_pitch, static_cast<__half*>(p16), p_pitch, inv_sum, cols);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_flow_head_update(const void* x16, int64_t pitch, const float* w, const float* bias, float* coords1,
                                     float* flow, int32_t batch, int32_t h8, int32_t w8, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(x16 && w && bias && coords1 && flow, ATDN_ERR_ARG, "atdn_flow_head_update: null argument");
  ATDN_REQUIRE(pitch % 8 == 0 && pitch >= 256 && aligned16(x16), ATDN_ERR_ALIGN, "atdn_flow_head_update: alignment");
  const long long npix = static_cast<long long>(batch) * h8 * w8;
  flow_head_update_kernel<<<static_cast<unsigned>((npix + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x16), pitch, w, bias, coords1, flow, batch, h8, w8);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_convex_upsample(const float* mask32, int64_t mask_pitch, const float* flow, float* flow_up, float* flow_lo,
                                    int32_t batch, int32_t h8, int32_t w8, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(mask32 && flow && flow_up && mask_pitch >= 576, ATDN_ERR_ARG, "atdn_convex_upsample: bad arguments");
  convex_upsample_kernel<<<dim3((w8 + 3) / 4, h8, batch), dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      mask32, mask_pitch, flow, flow_up, flow_lo, batch, h8, w8);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_coords_init(float* coords1, float* flow, const float* flow_init, int32_t batch, int32_t h8, int32_t w8,
                                void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(coords1 && flow, ATDN_ERR_ARG, "atdn_coords_init: null argument");
  const long long n = static_cast<long long>(batch) * h8 * w8;
  coords_init_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(coords1, flow, flow_init, batch, h8, w8);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}
