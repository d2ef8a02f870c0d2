// fp32 CUDA-core kernels for the small-channel networks: CLVO (ATDNVO) encoder + LSTM + heads, the
// MappingVAE keyframe encoder, and the keyframe L2 search.  These layers have 2..128 channels on
// small maps: memory/latency-bound, not tensor-core work.
#include <math.h>
#include <string.h>

#include "common.h"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace atdn {

__device__ __forceinline__ float mishf(float x) {
  // x * tanh(softplus(x)) with softplus thresholded at 20 like torch.  tanh(log(1 + e^x)) = (n^2 + 2n) / (n^2 + 2n + 2),
  // n = e^x: one MUFU exponential and one division (~2 ulp, all terms positive: no cancellation) instead of
  // expf + log1pf + tanhf (~100 instructions, which made the 16-channel conv layers epilogue-bound).
  const float n = __expf(x);
  const float t = n * (n + 2.0f);
  return x > 20.0f ? x : x * __fdividef(t, t + 2.0f);
}
// Branch-free variant for the register-tiled kernels (64 calls per thread): clamping the exponent argument at the
// softplus threshold makes t / (t + 2) round to exactly 1.0f there (t = 2.4e17), so x > 20 returns x without a select;
// ex2.approx + rcp.approx: ~3e-7 relative.
__device__ __forceinline__ float mish_fast(float x) {
  const float n = ex2_approx(fminf(x, 20.0f) * 1.4426950408889634f);
  const float t = n * (n + 2.0f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 2.0f));
  return x * (t * r);
}

// ------------------------------------------------------------------------------------------------
// Direct convolution, NCHW fp32.  CTA = 16x16 output pixels x 16 output channels; input channels are
// streamed through shared memory 4 at a time together with their [ci][tap][co16] weight slab.
// ------------------------------------------------------------------------------------------------
constexpr int kCoT = 16, kCiT = 4, kTile = 16;

struct Conv32Params {
  const float *x, *w, *bias, *in_scale, *in_shift, *skip, *bn_scale, *bn_shift, *bn2_scale, *bn2_shift;
  float* y;
  int B, Cin, Cout, H, W, OH, OW, K, stride, pad, mish, tiles_x, in_tile;
  int xp, yp;             // row pitch (elements) of x and of y / skip: W / OW when dense
};

__global__ void __launch_bounds__(256) conv32_kernel(Conv32Params p) {
  extern __shared__ float sm[];
  const int IT = p.in_tile;                       // (kTile-1)*stride + K
  float* s_in = sm;                               // [kCiT][IT][IT]
  float* s_w = sm + kCiT * IT * IT;               // [kCiT][K*K][kCoT]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int tile_x = blockIdx.x % p.tiles_x, tile_y = blockIdx.x / p.tiles_x;
  const int co0 = blockIdx.y * kCoT, b = blockIdx.z;
  const int ox = tile_x * kTile + tx, oy = tile_y * kTile + ty;
  const int ix0 = tile_x * kTile * p.stride - p.pad, iy0 = tile_y * kTile * p.stride - p.pad;
  const int KK = p.K * p.K;
  float acc[kCoT];
#pragma unroll
  for (int i = 0; i < kCoT; ++i) acc[i] = 0.0f;

  for (int ci0 = 0; ci0 < p.Cin; ci0 += kCiT) {
    __syncthreads();
    for (int i = threadIdx.x; i < kCiT * IT * IT; i += 256) {
      const int c = i / (IT * IT), r = i - c * IT * IT;
      const int yy = iy0 + r / IT, xx = ix0 + r % IT;
      float v = 0.0f;
      const int ci = ci0 + c;
      if (ci < p.Cin && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
        v = __ldg(p.x + ((static_cast<long long>(b) * p.Cin + ci) * p.H + yy) * p.xp + xx);
        if (p.in_scale) v = v * __ldg(p.in_scale + ci) + __ldg(p.in_shift + ci);
      }
      s_in[i] = v;
    }
    for (int i = threadIdx.x; i < kCiT * KK * kCoT; i += 256) {
      const int co = i % kCoT, t = (i / kCoT) % KK, c = i / (kCoT * KK);
      const int ci = ci0 + c;
      float v = 0.0f;
      if (ci < p.Cin && co0 + co < p.Cout) v = __ldg(p.w + (static_cast<long long>(co0 + co) * p.Cin + ci) * KK + t);
      s_w[i] = v;
    }
    __syncthreads();
    for (int c = 0; c < kCiT; ++c) {
      const float* in_c = s_in + c * IT * IT + (ty * p.stride) * IT + tx * p.stride;
      const float* w_c = s_w + c * KK * kCoT;
      for (int ky = 0; ky < p.K; ++ky) {
        for (int kx = 0; kx < p.K; ++kx) {
          const float v = in_c[ky * IT + kx];
          const float4* w4 = reinterpret_cast<const float4*>(w_c + (ky * p.K + kx) * kCoT);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 ww = w4[q];
            acc[4 * q] += v * ww.x;
            acc[4 * q + 1] += v * ww.y;
            acc[4 * q + 2] += v * ww.z;
            acc[4 * q + 3] += v * ww.w;
          }
        }
      }
    }
  }
  if (ox >= p.OW || oy >= p.OH) return;
#pragma unroll
  for (int i = 0; i < kCoT; ++i) {
    const int co = co0 + i;
    if (co >= p.Cout) break;
    const long long o = ((static_cast<long long>(b) * p.Cout + co) * p.OH + oy) * p.yp + ox;
    float v = acc[i] + (p.bias ? __ldg(p.bias + co) : 0.0f);
    if (p.mish) v = mishf(v);
    if (p.bn_scale) v = v * __ldg(p.bn_scale + co) + __ldg(p.bn_shift + co);
    if (p.skip) {
      v = mishf(v + p.skip[o]);
      if (p.bn2_scale) v = v * __ldg(p.bn2_scale + co) + __ldg(p.bn2_shift + co);
    }
    p.y[o] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Register-tiled direct convolution for the 16-output-channel layers of the CLVO encoder (the per-pair hot
// ones: 7x7/2 stem on the flow, 3x3/1 and 3x3/2 of the residual blocks).  The generic kernel above issues one
// scalar + four 128-bit shared loads per 16 FMAs (LSU-bound, 9 TFLOP/s); here
//   * CTA = 64 x 8 output pixels x 16 output channels, 128 threads, thread = 4 consecutive pixels x 16 channels
//     (64 fp32 accumulators): per (input channel, filter row) a thread loads its 4*S+K-S input values once
//     (128-bit loads) and reuses them for K taps x 16 channels; the 16 weights of a tap are 4 broadcast loads
//     feeding 64 FMAs;
//   * stride-2 layers stage the input tile de-interleaved (even | odd columns) so the per-thread runs stay
//     contiguous and conflict-free.
// fp32 throughout (1e-4 relative pose tolerance, north star), same epilogue as the generic kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kC16TW = 64, kC16TH = 8, kC16Threads = 128;

template <int K, int S>
struct C16Geom {
  static constexpr int IH = (kC16TH - 1) * S + K;
  static constexpr int IW = (kC16TW - 1) * S + K;
  static constexpr int NV = 3 * S + K;                       // input values per thread and filter row
  // stride 1: one row of IW (padded to a multiple of 4); stride 2: even columns then odd columns, each padded
  static constexpr int HALF = ((IW + 1) / 2 + 3) / 4 * 4 + 4;
  static constexpr int PITCH = S == 1 ? (IW + 3) / 4 * 4 + 4 : 2 * HALF;
};

template <int K, int S, int CIT>
__global__ void __launch_bounds__(kC16Threads) conv16_kernel(Conv32Params p) {
  using G = C16Geom<K, S>;
  extern __shared__ float sm[];
  float* s_in = sm;                                 // [CIT][IH][PITCH]
  float* s_w = sm + CIT * G::IH * G::PITCH;         // [CIT][K*K][16]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int tile_x = blockIdx.x % p.tiles_x, tile_y = blockIdx.x / p.tiles_x;
  const int b = blockIdx.z;
  const int ox0 = tile_x * kC16TW + tx * 4, oy = tile_y * kC16TH + ty;
  const int ix0 = tile_x * kC16TW * S - p.pad, iy0 = tile_y * kC16TH * S - p.pad;
  float acc[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[i][c] = 0.0f;

  for (int ci0 = 0; ci0 < p.Cin; ci0 += CIT) {
    __syncthreads();
    // 8 independent global loads in flight per thread before the first shared store (a load -> store loop
    // serialises one L2 round trip per element: measured 3x slower end to end)
    constexpr int kN = CIT * G::IH * G::IW;
    for (int i0 = threadIdx.x; i0 < kN; i0 += kC16Threads * 8) {
      float v[8];
      int dst[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * kC16Threads;
        v[u] = 0.0f;
        dst[u] = -1;
        if (i < kN) {
          const int c = i / (G::IH * G::IW), r = i - c * (G::IH * G::IW);
          const int ry = r / G::IW, rx = r - ry * G::IW;
          const int yy = iy0 + ry, xx = ix0 + rx, ci = ci0 + c;
          dst[u] = (c * G::IH + ry) * G::PITCH + (S == 1 ? rx : (rx & 1) * G::HALF + (rx >> 1));
          if (ci < p.Cin && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
            v[u] = __ldg(p.x + ((static_cast<long long>(b) * p.Cin + ci) * p.H + yy) * p.xp + xx);
            if (p.in_scale) v[u] = v[u] * __ldg(p.in_scale + ci) + __ldg(p.in_shift + ci);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (dst[u] >= 0) s_in[dst[u]] = v[u];
    }
    for (int i = threadIdx.x; i < CIT * K * K * 16; i += kC16Threads) {
      const int co = i & 15, t = (i >> 4) % (K * K), c = i / (16 * K * K);
      const int ci = ci0 + c;
      s_w[i] = ci < p.Cin ? __ldg(p.w + (static_cast<long long>(co) * p.Cin + ci) * (K * K) + t) : 0.0f;
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CIT; ++c) {
#pragma unroll 1
      for (int ky = 0; ky < K; ++ky) {
        const float* row = s_in + (c * G::IH + ty * S + ky) * G::PITCH;
        float in[G::NV + 3];
        if (S == 1) {
#pragma unroll
          for (int q = 0; q < (G::NV + 3) / 4; ++q) {
            const float4 f = *reinterpret_cast<const float4*>(row + tx * 4 + q * 4);
            in[4 * q] = f.x; in[4 * q + 1] = f.y; in[4 * q + 2] = f.z; in[4 * q + 3] = f.w;
          }
        } else {
          constexpr int NE = (G::NV + 1) / 2, NO = G::NV / 2;      // even / odd values needed
          float ev[(NE + 3) / 4 * 4], od[(NO + 3) / 4 * 4];
#pragma unroll
          for (int q = 0; q < (NE + 3) / 4; ++q) {
            const float4 f = *reinterpret_cast<const float4*>(row + tx * 4 + q * 4);
            ev[4 * q] = f.x; ev[4 * q + 1] = f.y; ev[4 * q + 2] = f.z; ev[4 * q + 3] = f.w;
          }
#pragma unroll
          for (int q = 0; q < (NO + 3) / 4; ++q) {
            const float4 f = *reinterpret_cast<const float4*>(row + G::HALF + tx * 4 + q * 4);
            od[4 * q] = f.x; od[4 * q + 1] = f.y; od[4 * q + 2] = f.z; od[4 * q + 3] = f.w;
          }
#pragma unroll
          for (int j = 0; j < G::NV; ++j) in[j] = (j & 1) ? od[j >> 1] : ev[j >> 1];
        }
        const float4* w4 = reinterpret_cast<const float4*>(s_w + (c * K * K + ky * K) * 16);
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          float w[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 f = w4[kx * 4 + q];
            w[4 * q] = f.x; w[4 * q + 1] = f.y; w[4 * q + 2] = f.z; w[4 * q + 3] = f.w;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float v = in[i * S + kx];
#pragma unroll
            for (int co = 0; co < 16; ++co) acc[i][co] += v * w[co];
          }
        }
      }
    }
  }
  if (oy >= p.OH) return;
  const bool vec = (p.yp & 3) == 0 && ox0 + 4 <= p.yp;
#pragma unroll
  for (int co = 0; co < 16; ++co) {
    const long long o = ((static_cast<long long>(b) * 16 + co) * p.OH + oy) * p.yp + ox0;
    const float bias = p.bias ? __ldg(p.bias + co) : 0.0f;
    const float s1 = p.bn_scale ? __ldg(p.bn_scale + co) : 1.0f, h1 = p.bn_scale ? __ldg(p.bn_shift + co) : 0.0f;
    const float s2 = p.bn2_scale ? __ldg(p.bn2_scale + co) : 1.0f, h2 = p.bn2_scale ? __ldg(p.bn2_shift + co) : 0.0f;
    float v[4], sk[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (p.skip) {
      if (vec) {
        const float4 f = *reinterpret_cast<const float4*>(p.skip + o);
        sk[0] = f.x; sk[1] = f.y; sk[2] = f.z; sk[3] = f.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (ox0 + i < p.OW) sk[i] = p.skip[o + i];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t = acc[i][co] + bias;
      if (p.mish) t = mishf(t);
      if (p.bn_scale) t = t * s1 + h1;
      if (p.skip) {
        t = mishf(t + sk[i]);
        if (p.bn2_scale) t = t * s2 + h2;
      }
      v[i] = t;
    }
    if (vec) {
      *reinterpret_cast<float4*>(p.y + o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (ox0 + i < p.OW) p.y[o + i] = v[i];
    }
  }
}

template <int K, int S, int CIT>
static int launch_conv16(Conv32Params p, cudaStream_t stream) {
  using G = C16Geom<K, S>;
  constexpr int smem = (CIT * G::IH * G::PITCH + CIT * K * K * 16) * (int)sizeof(float);
  static DeviceOnce configured;
  if (configured.pending()) {
    if (smem > 48 * 1024) ATDN_CUDA(cudaFuncSetAttribute(conv16_kernel<K, S, CIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.done();
  }
  p.tiles_x = ceil_div(p.OW, kC16TW);
  dim3 grid(p.tiles_x * ceil_div(p.OH, kC16TH), 1, p.B);
  conv16_kernel<K, S, CIT><<<grid, kC16Threads, smem, stream>>>(p);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// TMA-fed variant with the weights in the constant bank (the per-pair hot layers of the CLVO encoder).
//
// ncu on conv16_kernel (B200, 27 pairs, 3x3/1 at 188x616): 20K instructions per thread for 9.2K FFMAs, issue slots
// 52% busy, 42% of the stall samples on the first use of the staged global loads (long scoreboard) -- the per-element
// global->shared staging (index arithmetic, bounds checks, exposed L2 latency) costs more than the weights.  Here
//   * the input tile arrives by TMA: one cp.async.bulk.tensor box {68 columns, IH rows, CIT channels} of the NCHW
//     fp32 input per stage (zero fill outside the image = the conv zero padding, negative coordinates included),
//     double-buffered over channel chunks behind an mbarrier pair.  No thread executes a staging instruction;
//     every thread reads its (3 S + K)-value row segment with aligned 128-bit shared loads.  Needs a row pitch
//     that is a multiple of 4 floats (16-byte TMA strides): the encoder pads its intermediate maps
//     (154 -> 156, 77 -> 80, 39 -> 40 columns);
//   * the whole filter ([ci][ky][kx][16 co], <= 9 KiB) is staged once per CTA while the first tile is in flight
//     and read as warp-wide broadcasts (4 x LDS.128 per 64 FFMAs);
//   * AFFINE (the stem): x * in_scale[c] + in_shift[c] is applied to the in-bounds elements of the landed tile.
// CTA = 128 threads = 16 x 8, tile = 64 x 8 output pixels, thread = 4 consecutive pixels x 16 channels.
// ------------------------------------------------------------------------------------------------
template <int K, int S, int PAD>
struct C16tGeom {
  static constexpr int IH = (kC16TH - 1) * S + K;
  // TMA boxes of an fp32 NCHW map must START on a 16-byte boundary (x coordinate a multiple of 4: verified with
  // tools/tma_f32_test.cu, a box at x = -1 faults), so the box begins OFF columns left of the first tap
  static constexpr int OFF = (4 - PAD % 4) % 4;
  static constexpr int NV = 3 * S + K;                        // input values per thread and filter row
  static constexpr int NL = (OFF + NV + 3) / 4;               // 128-bit shared loads per thread and filter row
  static constexpr int ROW = (OFF + 63 * S + K + 3) / 4 * 4;  // floats per staged row
  static_assert(15 * 4 * S + 4 * NL <= ROW, "a thread's row segment must stay inside the staged row");
  static_assert(ROW <= 256, "TMA box limit");
};

template <int K, int S, int PAD, int CIN, int CIT, bool AFFINE, bool SKIP>
__global__ void __launch_bounds__(kC16Threads) conv16t_kernel(const __grid_constant__ CUtensorMap tmx,
                                                              const __grid_constant__ Conv32Params p) {
  using G = C16tGeom<K, S, PAD>;
  static_assert(CIN % CIT == 0, "input channels are staged in whole chunks");
  constexpr int NCH = CIN / CIT;
  constexpr int BOX_FLOATS = CIT * G::IH * G::ROW;            // one TMA box: [CIT][IH][ROW]
  constexpr int STAGE_FLOATS = (BOX_FLOATS + 31) / 32 * 32;   // TMA destinations are 128-byte aligned
  constexpr int NSTAGE = NCH > 1 ? 2 : 1;
  extern __shared__ uint8_t smem_raw16[];
  __shared__ __align__(8) uint64_t full[2];
  float* s_in = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw16) + 127) & ~uintptr_t(127));
  float* s_w = s_in + NSTAGE * STAGE_FLOATS;        // [CIN][K*K][16]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int tile_x = blockIdx.x % p.tiles_x, tile_y = blockIdx.x / p.tiles_x;
  const int b = blockIdx.z;
  const int ox0 = tile_x * kC16TW + tx * 4, oy = tile_y * kC16TH + ty;
  const int bx0 = tile_x * kC16TW * S - PAD - G::OFF, iy0 = tile_y * kC16TH * S - PAD;   // bx0 is a multiple of 4

  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int chunk) {
    uint64_t* bar = &full[chunk % NSTAGE];
    mbar_arrive_expect_tx(bar, BOX_FLOATS * 4);
    tma_load_4d(s_in + (chunk % NSTAGE) * STAGE_FLOATS, &tmx, bar, bx0, iy0, chunk * CIT, b);
  };
  // TMA issue follows the pattern of the tensor-core kernels: a warp-uniform branch, one elected lane
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp == 0) {
    if (elect_one_sync()) {
      tma_prefetch_desc(&tmx);
      issue(0);
    }
    __syncwarp();
  }
  for (int i = threadIdx.x; i < CIN * K * K * 16; i += kC16Threads) {   // [co][ci][tap] (PyTorch) -> [ci][tap][co]
    const int co = i & 15, t = (i >> 4) % (K * K), ci = i / (16 * K * K);
    s_w[i] = __ldg(p.w + (co * CIN + ci) * (K * K) + t);
  }
  __syncthreads();

  float acc[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[i][c] = 0.0f;

#pragma unroll 1
  for (int chunk = 0; chunk < NCH; ++chunk) {
    if (warp == 0 && chunk + 1 < NCH) {            // the other stage was released by the barrier at the end of the loop
      if (elect_one_sync()) issue(chunk + 1);
      __syncwarp();
    }
    mbar_wait(&full[chunk % NSTAGE], static_cast<uint32_t>((chunk / NSTAGE) & 1));
    float* st = s_in + (chunk % NSTAGE) * STAGE_FLOATS;
    if constexpr (AFFINE) {
      for (int e = threadIdx.x; e < BOX_FLOATS; e += kC16Threads) {
        const int j = e % G::ROW;
        const int r = e / G::ROW;
        const int ry = r % G::IH, c = r / G::IH;
        const int xx = bx0 + j, yy = iy0 + ry;
        if (xx >= 0 && xx < p.W && yy >= 0 && yy < p.H) {
          const int ci = chunk * CIT + c;
          st[e] = st[e] * __ldg(p.in_scale + ci) + __ldg(p.in_shift + ci);
        }
      }
      __syncthreads();
    }
#pragma unroll 1
    for (int c = 0; c < CIT; ++c) {
      const float4* wc = reinterpret_cast<const float4*>(s_w + (chunk * CIT + c) * (K * K * 16));   // warp-wide broadcasts
#pragma unroll K == 3 ? 3 : 1
      for (int ky = 0; ky < K; ++ky) {
        const float* row = st + (c * G::IH + ty * S + ky) * G::ROW + tx * 4 * S;
        float buf[4 * G::NL];
#pragma unroll
        for (int q = 0; q < G::NL; ++q) {
          const float4 f = *reinterpret_cast<const float4*>(row + q * 4);
          buf[4 * q] = f.x; buf[4 * q + 1] = f.y; buf[4 * q + 2] = f.z; buf[4 * q + 3] = f.w;
        }
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          float w[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 f = wc[(ky * K + kx) * 4 + q];
            w[4 * q] = f.x; w[4 * q + 1] = f.y; w[4 * q + 2] = f.z; w[4 * q + 3] = f.w;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float v = buf[G::OFF + i * S + kx];
#pragma unroll
            for (int co = 0; co < 16; ++co) acc[i][co] = fmaf(v, w[co], acc[i][co]);
          }
        }
      }
    }
    if (chunk + 1 < NCH) __syncthreads();          // all reads of this stage are done before it is refilled
  }
  // epilogue of a Conv block (layers/conv.py:36-37: bn(mish(conv + bias))), SKIP: ResidualConv tail bn2(mish(. + skip))
  if (oy >= p.OH) return;
  const bool v4 = (p.yp & 3) == 0 && ox0 + 4 <= p.yp;
#pragma unroll
  for (int co = 0; co < 16; ++co) {
    const long long o = ((static_cast<long long>(b) * 16 + co) * p.OH + oy) * p.yp + ox0;
    const float bias = __ldg(p.bias + co);
    const float s1 = __ldg(p.bn_scale + co), h1 = __ldg(p.bn_shift + co);
    float s2 = 1.0f, h2 = 0.0f;
    float v[4], sk[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if constexpr (SKIP) {
      s2 = __ldg(p.bn2_scale + co);
      h2 = __ldg(p.bn2_shift + co);
      if (v4) {
        const float4 f = *reinterpret_cast<const float4*>(p.skip + o);
        sk[0] = f.x; sk[1] = f.y; sk[2] = f.z; sk[3] = f.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (ox0 + i < p.OW) sk[i] = p.skip[o + i];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t = mish_fast(acc[i][co] + bias) * s1 + h1;
      if constexpr (SKIP) t = mish_fast(t + sk[i]) * s2 + h2;
      v[i] = t;
    }
    if (v4) {
      *reinterpret_cast<float4*>(p.y + o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (ox0 + i < p.OW) p.y[o + i] = v[i];
    }
  }
}

template <int K, int S, int PAD, int CIN, int CIT, bool AFFINE, bool SKIP>
static int launch_conv16t(Conv32Params p, cudaStream_t stream) {
  using G = C16tGeom<K, S, PAD>;
  constexpr int NSTAGE = CIN / CIT > 1 ? 2 : 1;
  constexpr int STAGE_FLOATS = (CIT * G::IH * G::ROW + 31) / 32 * 32;
  constexpr int smem = (NSTAGE * STAGE_FLOATS + CIN * K * K * 16) * (int)sizeof(float) + 128;
  static DeviceOnce configured;
  if (configured.pending()) {
    if (smem > 48 * 1024)
      ATDN_CUDA(cudaFuncSetAttribute(conv16t_kernel<K, S, PAD, CIN, CIT, AFFINE, SKIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.done();
  }
  CUtensorMap tmx;
  {
    const int64_t dims[4] = {p.W, p.H, CIN, p.B};
    const int64_t str[3] = {p.xp, (int64_t)p.H * p.xp, (int64_t)CIN * p.H * p.xp};
    const uint32_t box[4] = {(uint32_t)G::ROW, (uint32_t)G::IH, (uint32_t)CIT, 1};
    const uint32_t ones[4] = {1, 1, 1, 1};
    if (int e = make_map(&tmx, 4, CU_TENSOR_MAP_SWIZZLE_NONE, p.x, dims, str, box, ones, "conv16 input")) return e;
  }
  p.tiles_x = ceil_div(p.OW, kC16TW);
  dim3 grid(p.tiles_x * ceil_div(p.OH, kC16TH), 1, p.B);
  conv16t_kernel<K, S, PAD, CIN, CIT, AFFINE, SKIP><<<grid, kC16Threads, smem, stream>>>(tmx, p);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// ResidualConv skip path (layers/conv.py:86): 1x1 stride-2 convolution 16 -> 16 channels, bias only.  The generic
// kernel above stages a 31 x 31 input tile per 16 x 16 outputs and took 0.3 ms per 27 pairs at 188x616 (as long as
// the 3x3/2 convolution next to it); this one is a pure streaming kernel: thread = 2 adjacent output pixels,
// 16 strided input loads each, 256 weights in shared memory, coalesced 8-byte stores per output channel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) skip16_kernel(Conv32Params p) {
  __shared__ float s_w[16 * 16];                       // [ci][co]
  __shared__ float s_b[16];
  s_w[threadIdx.x] = __ldg(p.w + (threadIdx.x & 15) * 16 + (threadIdx.x >> 4));
  if (threadIdx.x < 16) s_b[threadIdx.x] = p.bias ? __ldg(p.bias + threadIdx.x) : 0.0f;
  __syncthreads();
  const int ow2 = (p.OW + 1) >> 1;                     // pixel pairs per output row
  const long long total = static_cast<long long>(p.B) * p.OH * ow2;
  for (long long t = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; t < total; t += static_cast<long long>(gridDim.x) * 256) {
    const int xp2 = static_cast<int>(t % ow2);
    const long long r = t / ow2;
    const int oy = static_cast<int>(r % p.OH), b = static_cast<int>(r / p.OH);
    const int ox = 2 * xp2;
    const bool two = ox + 1 < p.OW;
    const float* xin = p.x + (static_cast<long long>(b) * 16 * p.H + 2 * oy) * p.xp + 2 * ox;
    float v0[16], v1[16];
#pragma unroll
    for (int ci = 0; ci < 16; ++ci) {
      const float* px = xin + static_cast<long long>(ci) * p.H * p.xp;
      v0[ci] = __ldg(px);
      v1[ci] = two ? __ldg(px + 2) : 0.0f;
    }
    float* yo = p.y + (static_cast<long long>(b) * 16 * p.OH + oy) * p.yp + ox;
#pragma unroll
    for (int co = 0; co < 16; ++co) {
      float a0 = s_b[co], a1 = s_b[co];
#pragma unroll
      for (int ci = 0; ci < 16; ++ci) {
        const float w = s_w[ci * 16 + co];
        a0 = fmaf(v0[ci], w, a0);
        a1 = fmaf(v1[ci], w, a1);
      }
      float* dst = yo + static_cast<long long>(co) * p.OH * p.yp;
      if (two && (p.yp & 1) == 0) {
        *reinterpret_cast<float2*>(dst) = make_float2(a0, a1);
      } else {
        dst[0] = a0;
        if (two) dst[1] = a1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Linear (+Mish) and LSTM cell: one warp per output row, looping over the batch
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) linear32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ y, int B,
                                                       int IN, int OUT, int act) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= OUT) return;
  const float* wr = w + static_cast<long long>(row) * IN;
  for (int b = 0; b < B; ++b) {
    const float* xb = x + static_cast<long long>(b) * IN;
    float acc = 0.0f;
    for (int i = lane; i < IN; i += 32) acc += __ldg(wr + i) * xb[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float v = acc + (bias ? bias[row] : 0.0f);
      if (act == 1) v = mishf(v);
      y[static_cast<long long>(b) * OUT + row] = v;
    }
  }
}

__global__ void __launch_bounds__(256) lstm_gates_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                         const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                                         const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                                         float* __restrict__ gates, int B, int IN, int HID) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= 4 * HID) return;
  const float* wi = w_ih + static_cast<long long>(row) * IN;
  const float* wh = w_hh + static_cast<long long>(row) * HID;
  for (int b = 0; b < B; ++b) {
    float a1 = 0.0f, a2 = 0.0f;
    for (int i = lane; i < IN; i += 32) a1 += __ldg(wi + i) * x[static_cast<long long>(b) * IN + i];
    for (int i = lane; i < HID; i += 32) a2 += __ldg(wh + i) * h[static_cast<long long>(b) * HID + i];
    for (int o = 16; o > 0; o >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    // torch: (x W_ih^T + b_ih) + (h W_hh^T + b_hh)
    if (lane == 0) gates[static_cast<long long>(b) * 4 * HID + row] = (a1 + b_ih[row]) + (a2 + b_hh[row]);
  }
}

__global__ void lstm_pointwise_kernel(const float* __restrict__ gates, float* __restrict__ h, float* __restrict__ c,
                                      int B, int HID) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * HID) return;
  const int b = idx / HID, j = idx - b * HID;
  const float* g = gates + static_cast<long long>(b) * 4 * HID;
  const float ig = 1.0f / (1.0f + expf(-g[j]));
  const float fg = 1.0f / (1.0f + expf(-g[HID + j]));
  const float gg = tanhf(g[2 * HID + j]);
  const float og = 1.0f / (1.0f + expf(-g[3 * HID + j]));
  const float cn = fg * c[idx] + ig * gg;
  c[idx] = cn;
  h[idx] = og * tanhf(cn);
}

// ------------------------------------------------------------------------------------------------
// Keyframe search: streaming squared-L2 per row (one CTA per keyframe), then first arg-min
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kf_dist_kernel(const float* __restrict__ emb, const float* __restrict__ code,
                                                      float* __restrict__ dist, int dim) {
  __shared__ float red[8];
  const float4* e = reinterpret_cast<const float4*>(emb + static_cast<long long>(blockIdx.x) * dim);
  const float4* q = reinterpret_cast<const float4*>(code);
  float acc = 0.0f;
  for (int i = threadIdx.x; i < dim / 4; i += 256) {
    const float4 a = __ldg(e + i), b = __ldg(q + i);
    const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
    acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < 8; ++i) t += red[i];
    dist[blockIdx.x] = sqrtf(t);
  }
}

// torch.argmin semantics (neural_slam.py:383): the FIRST minimum; a NaN distance counts as the minimum (ATen propagates NaN),
// so a corrupt embedding yields a valid index (the NaN's) instead of the sentinel.
__device__ __forceinline__ bool kf_better(float v, long long i, float bv, long long bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn || bn) return vn && (!bn || i < bi);
  return v < bv || (v == bv && i < bi);
}

__global__ void __launch_bounds__(1024) kf_argmin_kernel(const float* __restrict__ dist, long long n,
                                                         int* __restrict__ index) {
  __shared__ float sv[1024];
  __shared__ long long si[1024];
  float best = INFINITY;
  long long bi = 0x7fffffffffffffffLL;
  for (long long i = threadIdx.x; i < n; i += 1024) {
    const float v = dist[i];
    if (kf_better(v, i, best, bi)) { best = v; bi = i; }
  }
  sv[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const float v = sv[threadIdx.x + s];
      const long long j = si[threadIdx.x + s];
      if (kf_better(v, j, sv[threadIdx.x], si[threadIdx.x])) {
        sv[threadIdx.x] = v;
        si[threadIdx.x] = j;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *index = n > 0 ? static_cast<int>(si[0]) : -1;
}

}  // namespace atdn

using namespace atdn;

extern "C" int atdn_conv32(const atdn_conv32_desc* d, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(d && d->x && d->y && d->w, ATDN_ERR_ARG, "atdn_conv32: null argument");
  ATDN_REQUIRE(d->k >= 1 && d->k <= 7 && d->stride >= 1 && d->stride <= 3, ATDN_ERR_UNSUP, "atdn_conv32: k=%d stride=%d", d->k, d->stride);
  ATDN_REQUIRE((d->in_scale == nullptr) == (d->in_shift == nullptr) && (d->bn_scale == nullptr) == (d->bn_shift == nullptr) &&
                   (d->bn2_scale == nullptr) == (d->bn2_shift == nullptr), ATDN_ERR_ARG,
               "atdn_conv32: scale/shift must come in pairs");
  Conv32Params p;
  p.x = d->x; p.w = d->w; p.bias = d->bias; p.in_scale = d->in_scale; p.in_shift = d->in_shift; p.skip = d->skip;
  p.bn_scale = d->bn_scale; p.bn_shift = d->bn_shift; p.bn2_scale = d->bn2_scale; p.bn2_shift = d->bn2_shift; p.y = d->y;
  p.B = d->batch; p.Cin = d->cin; p.Cout = d->cout; p.H = d->in_h; p.W = d->in_w; p.K = d->k; p.stride = d->stride;
  p.pad = d->pad; p.mish = d->mish;
  p.OH = (d->in_h + 2 * d->pad - d->k) / d->stride + 1;
  p.OW = (d->in_w + 2 * d->pad - d->k) / d->stride + 1;
  p.xp = d->x_pitch > 0 ? d->x_pitch : d->in_w;
  p.yp = d->y_pitch > 0 ? d->y_pitch : p.OW;
  ATDN_REQUIRE(p.xp >= d->in_w && p.yp >= p.OW, ATDN_ERR_ARG, "atdn_conv32: row pitch smaller than the row");
  ATDN_REQUIRE(p.OH >= 1 && p.OW >= 1, ATDN_ERR_ARG, "atdn_conv32: empty output");
  if (d->cout == 16 && d->cin == 16 && d->k == 1 && d->stride == 2 && d->pad == 0 && !d->mish && !d->in_scale && !d->bn_scale &&
      !d->skip && (reinterpret_cast<uintptr_t>(d->y) & 7u) == 0) {   // ResidualConv skip path
    const long long total = static_cast<long long>(p.B) * p.OH * ((p.OW + 1) / 2);
    const long long blocks = (total + 255) / 256;
    skip16_kernel<<<static_cast<unsigned>(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    ATDN_CUDA(cudaGetLastError());
    return 0;
  }
  if (d->cout == 16 && d->cin <= 16 && p.OW >= 24) {   // CLVO encoder layers: register-tiled kernels
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // 16-byte rows + the Conv-block epilogue (bias, mish, bn; skip with bn2): TMA-fed tiles
    const bool block = d->bias && d->mish && d->bn_scale && (!d->skip || d->bn2_scale);
    if (block && p.xp % 4 == 0 && aligned16(d->x)) {
      const bool k3 = d->k == 3 && d->pad == 1 && d->cin == 16 && !d->in_scale;
      if (k3 && d->stride == 1 && !d->skip) return launch_conv16t<3, 1, 1, 16, 8, false, false>(p, st);
      if (k3 && d->stride == 1 && d->skip) return launch_conv16t<3, 1, 1, 16, 8, false, true>(p, st);
      if (k3 && d->stride == 2 && !d->skip) return launch_conv16t<3, 2, 1, 16, 2, false, false>(p, st);
      if (k3 && d->stride == 2 && d->skip) return launch_conv16t<3, 2, 1, 16, 2, false, true>(p, st);
      if (d->k == 7 && d->stride == 2 && d->pad == 3 && d->cin == 2 && d->in_scale && !d->skip)
        return launch_conv16t<7, 2, 3, 2, 2, true, false>(p, st);
    }
    if (d->k == 3 && d->stride == 1) return launch_conv16<3, 1, 8>(p, st);
    if (d->k == 3 && d->stride == 2) return launch_conv16<3, 2, 4>(p, st);
    if (d->k == 7 && d->stride == 2 && d->cin <= 2) return launch_conv16<7, 2, 2>(p, st);
  }
  p.tiles_x = ceil_div(p.OW, kTile);
  p.in_tile = (kTile - 1) * d->stride + d->k;
  const int smem = (kCiT * p.in_tile * p.in_tile + kCiT * d->k * d->k * kCoT) * (int)sizeof(float);
  static int max_smem = 48 * 1024;
  if (smem > max_smem) {
    ATDN_CUDA(cudaFuncSetAttribute(conv32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    max_smem = smem;
  }
  dim3 grid(p.tiles_x * ceil_div(p.OH, kTile), ceil_div(p.Cout, kCoT), p.B);
  conv32_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(p);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_linear32(const float* x, const float* w, const float* bias, float* y, int32_t batch, int32_t in_f,
                             int32_t out_f, int32_t act, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(x && w && y && batch >= 1 && in_f >= 1 && out_f >= 1, ATDN_ERR_ARG, "atdn_linear32: bad arguments");
  linear32_kernel<<<ceil_div(out_f, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, y, batch, in_f, out_f, act);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_lstm_cell(const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                              float* h, float* c, float* gates, int32_t batch, int32_t in_f, int32_t hidden, void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(x && w_ih && w_hh && b_ih && b_hh && h && c && gates, ATDN_ERR_ARG, "atdn_lstm_cell: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  lstm_gates_kernel<<<ceil_div(4 * hidden, 8), 256, 0, s>>>(x, h, w_ih, w_hh, b_ih, b_hh, gates, batch, in_f, hidden);
  ATDN_CUDA(cudaGetLastError());
  lstm_pointwise_kernel<<<ceil_div(batch * hidden, 256), 256, 0, s>>>(gates, h, c, batch, hidden);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int atdn_keyframe_search(const float* emb, const float* code, float* dist, int32_t* index, int64_t num, int32_t dim,
                                    void* stream) {
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(emb && code && dist && index && num >= 1, ATDN_ERR_ARG, "atdn_keyframe_search: bad arguments");
  ATDN_REQUIRE(dim % 4 == 0 && aligned16(emb) && aligned16(code), ATDN_ERR_ALIGN, "atdn_keyframe_search: dim %% 4 and 16-byte alignment required");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  kf_dist_kernel<<<static_cast<unsigned>(num), 256, 0, s>>>(emb, code, dist, dim);
  ATDN_CUDA(cudaGetLastError());
  kf_argmin_kernel<<<1, 1024, 0, s>>>(dist, num, index);
  ATDN_CUDA(cudaGetLastError());
  return 0;
}
