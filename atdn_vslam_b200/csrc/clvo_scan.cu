// Persistent recurrent scan of the CLVO pose network (atdn_vslam/odometry/network.py:137-140):
//   for t in 0..T-1:  (h1, c1) = LSTMCell1(feat[t], (h1, c1));  x2 = mish(W_ll h1 + b_ll);  (h2, c2) = LSTMCell2(x2, (h2, c2))
// The chain is strictly sequential (3 dependent mat-vecs per step).  Launching it as ~11 kernels per step
// costs 3000 launches for a 270-pair sequence (25 ms); here ONE cooperative kernel runs all T steps:
//   * 128 CTAs, CTA c owns hidden units 4c..4c+3 of both LSTMs and rows 4c..4c+3 of the linear layer; its
//     slices of W_hh1 (16 x 512), W_ll (4 x 512) and [W_ih2 | W_hh2] (16 x 1024) stay in shared memory (104 KB)
//     for the whole scan, the cell states stay in registers;
//   * the input-side term of LSTM1, feat[t] W_ih1^T + b_ih1, does not depend on the recurrence and is
//     precomputed for all T as one batched product (the caller passes it as `p1`);
//   * two grid barriers per step (after h1, after x2) on a global counter; h1 / h2 of every step are kept
//     ([T, B, 512]) so that the buffers double as the t-1 / t double buffer and h2 feeds the regressor heads
//     as one batched product after the scan.
// Gate order i, f, g, o and the association (x W_ih^T + b_ih) + (h W_hh^T + b_hh) follow torch.nn.LSTMCell.
#include <math.h>
#include <string.h>

#include "common.h"

namespace atdn {

constexpr int kScanCtas = 128;
constexpr int kScanThreads = 256;
constexpr int kScanHid = 512;
constexpr int kScanMaxBatch = 32;
constexpr int kScanSmemFloats = 16 * 512 + 4 * 512 + 16 * 1024;

struct ScanParams {
  const float* p1;        // [T, B, 2048] = feat W_ih1^T + b_ih1
  const float *w_hh1, *b_hh1, *w_ll, *b_ll, *w_ih2, *w_hh2, *b_ih2, *b_hh2;
  float *h1_0, *c1, *h2_0, *c2;      // [B, 512] states: read at t = 0, c written back at the end (h from h*_all[T-1])
  float *h1_all, *h2_all;            // [T, B, 512]
  float* x2;                         // [B, 512] scratch
  unsigned int* counter;             // zero-initialised grid-barrier counter
  int T, B;
};

__device__ __forceinline__ float mish_scan(float x) {
  const float sp = x > 20.0f ? x : log1pf(expf(x));
  return x * tanhf(sp);
}
__device__ __forceinline__ float sigmoid_scan(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int spins = 0;
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= target) break;
      if (++spins > (1u << 28)) __trap();   // a lost CTA must surface as a trapped launch, not a hung GPU
    }
  }
  __syncthreads();
}

// dot of a shared-memory weight row with a global (L2) vector written by other CTAs: ld.global.cg, never L1
__device__ __forceinline__ float dot512(const float* __restrict__ w, const float* v, int lane) {
  float acc = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(w + (i * 32 + lane) * 4);
    const float4 b = __ldcg(reinterpret_cast<const float4*>(v) + i * 32 + lane);
    acc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
  return acc;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kScanThreads, 1) clvo_scan_kernel(ScanParams p) {
  extern __shared__ float sw[];
  float* w1 = sw;                    // [16][512]  row r = gate * 4 + u  <->  W_hh1[gate * 512 + 4c + u]
  float* wl = sw + 16 * 512;         // [4][512]
  float* w2 = wl + 4 * 512;          // [16][1024] = [W_ih2 row | W_hh2 row]
  __shared__ float gates[kScanMaxBatch][16];
  const int c = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = p.B;

  for (int i = threadIdx.x; i < 16 * 512; i += kScanThreads) {
    const int r = i >> 9, k = i & 511;
    const int grow = (r >> 2) * kScanHid + 4 * c + (r & 3);
    w1[i] = p.w_hh1[static_cast<long long>(grow) * 512 + k];
    w2[r * 1024 + k] = p.w_ih2[static_cast<long long>(grow) * 512 + k];
    w2[r * 1024 + 512 + k] = p.w_hh2[static_cast<long long>(grow) * 512 + k];
  }
  for (int i = threadIdx.x; i < 4 * 512; i += kScanThreads) wl[i] = p.w_ll[static_cast<long long>(4 * c + (i >> 9)) * 512 + (i & 511)];

  // cell states of this CTA's 4 units: thread (b * 4 + u) owns c1[b][4c+u], c2[b][4c+u]
  const bool owner = threadIdx.x < 4 * B;
  const int ob = threadIdx.x >> 2, ou = threadIdx.x & 3;
  float c1 = 0.0f, c2 = 0.0f;
  if (owner) {
    c1 = p.c1[ob * kScanHid + 4 * c + ou];
    c2 = p.c2[ob * kScanHid + 4 * c + ou];
  }
  __syncthreads();
  unsigned int target = 0;

  for (int t = 0; t < p.T; ++t) {
    const float* h1_prev = t ? p.h1_all + static_cast<long long>(t - 1) * B * kScanHid : p.h1_0;
    const float* h2_prev = t ? p.h2_all + static_cast<long long>(t - 1) * B * kScanHid : p.h2_0;
    float* h1_cur = p.h1_all + static_cast<long long>(t) * B * kScanHid;
    float* h2_cur = p.h2_all + static_cast<long long>(t) * B * kScanHid;

    // ---- A: LSTM1 gates of this CTA's units: warp w computes rows 2w, 2w+1
    for (int b = 0; b < B; ++b) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = 2 * warp + rr;
        const float s = warp_sum(dot512(w1 + r * 512, h1_prev + b * kScanHid, lane));
        if (lane == 0) {
          const int grow = (r >> 2) * kScanHid + 4 * c + (r & 3);
          gates[b][r] = p.p1[(static_cast<long long>(t) * B + b) * 2048 + grow] + (s + p.b_hh1[grow]);
        }
      }
    }
    __syncthreads();
    if (owner) {
      const float ig = sigmoid_scan(gates[ob][ou]), fg = sigmoid_scan(gates[ob][4 + ou]);
      const float gg = tanhf(gates[ob][8 + ou]), og = sigmoid_scan(gates[ob][12 + ou]);
      c1 = fg * c1 + ig * gg;
      h1_cur[ob * kScanHid + 4 * c + ou] = og * tanhf(c1);
    }
    target += kScanCtas;
    grid_barrier(p.counter, target);

    // ---- B: x2 rows 4c..4c+3 = mish(W_ll h1 + b_ll)
    if (warp < 4) {
      for (int b = 0; b < B; ++b) {
        const float s = warp_sum(dot512(wl + warp * 512, h1_cur + b * kScanHid, lane));
        if (lane == 0) p.x2[b * kScanHid + 4 * c + warp] = mish_scan(s + p.b_ll[4 * c + warp]);
      }
    }
    target += kScanCtas;
    grid_barrier(p.counter, target);

    // ---- C: LSTM2
    for (int b = 0; b < B; ++b) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = 2 * warp + rr;
        const float a1 = warp_sum(dot512(w2 + r * 1024, p.x2 + b * kScanHid, lane));
        const float a2 = warp_sum(dot512(w2 + r * 1024 + 512, h2_prev + b * kScanHid, lane));
        if (lane == 0) {
          const int grow = (r >> 2) * kScanHid + 4 * c + (r & 3);
          gates[b][r] = (a1 + p.b_ih2[grow]) + (a2 + p.b_hh2[grow]);
        }
      }
    }
    __syncthreads();
    if (owner) {
      const float ig = sigmoid_scan(gates[ob][ou]), fg = sigmoid_scan(gates[ob][4 + ou]);
      const float gg = tanhf(gates[ob][8 + ou]), og = sigmoid_scan(gates[ob][12 + ou]);
      c2 = fg * c2 + ig * gg;
      h2_cur[ob * kScanHid + 4 * c + ou] = og * tanhf(c2);
    }
    __syncthreads();   // `gates` is rewritten by phase A of the next step
  }
  if (owner) {
    p.c1[ob * kScanHid + 4 * c + ou] = c1;
    p.c2[ob * kScanHid + 4 * c + ou] = c2;
  }
}

}  // namespace atdn

using namespace atdn;

extern "C" int atdn_clvo_lstm_scan(const float* p1, const float* w_hh1, const float* b_hh1, const float* w_ll,
                                   const float* b_ll, const float* w_ih2, const float* w_hh2, const float* b_ih2,
                                   const float* b_hh2, const float* h1_0, float* c1, const float* h2_0, float* c2,
                                   float* h1_all, float* h2_all, float* x2_scratch, uint32_t* counter, int32_t steps,
                                   int32_t batch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = require_sm100()) return e;
  ATDN_REQUIRE(p1 && w_hh1 && b_hh1 && w_ll && b_ll && w_ih2 && w_hh2 && b_ih2 && b_hh2 && h1_0 && c1 && h2_0 && c2 && h1_all &&
                   h2_all && x2_scratch && counter, ATDN_ERR_ARG, "atdn_clvo_lstm_scan: null argument");
  ATDN_REQUIRE(steps >= 1 && batch >= 1 && batch <= kScanMaxBatch, ATDN_ERR_ARG, "atdn_clvo_lstm_scan: steps %d, batch %d (max %d)", steps, batch, kScanMaxBatch);
  ATDN_REQUIRE(aligned16(h1_0) && aligned16(h2_0) && aligned16(h1_all) && aligned16(h2_all) && aligned16(x2_scratch), ATDN_ERR_ALIGN,
               "atdn_clvo_lstm_scan: state buffers must be 16-byte aligned");
  int dev = 0, sms = 0, coop = 0;
  ATDN_CUDA(cudaGetDevice(&dev));
  ATDN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ATDN_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  ATDN_REQUIRE(coop && sms >= kScanCtas, ATDN_ERR_UNSUP, "atdn_clvo_lstm_scan: needs cooperative launch and >= %d SMs", kScanCtas);
  ScanParams p;
  p.p1 = p1; p.w_hh1 = w_hh1; p.b_hh1 = b_hh1; p.w_ll = w_ll; p.b_ll = b_ll; p.w_ih2 = w_ih2; p.w_hh2 = w_hh2;
  p.b_ih2 = b_ih2; p.b_hh2 = b_hh2;
  p.h1_0 = const_cast<float*>(h1_0); p.c1 = c1; p.h2_0 = const_cast<float*>(h2_0); p.c2 = c2;
  p.h1_all = h1_all; p.h2_all = h2_all; p.x2 = x2_scratch; p.counter = counter; p.T = steps; p.B = batch;
  const int smem = kScanSmemFloats * (int)sizeof(float);
  static DeviceOnce configured;
  if (configured.pending()) {
    ATDN_CUDA(cudaFuncSetAttribute(clvo_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.done();
  }
  ATDN_CUDA(cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream));
  void* args[] = {&p};
  ATDN_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(clvo_scan_kernel), dim3(kScanCtas), dim3(kScanThreads), args, smem, stream));
  return 0;
}
