// Microbenchmark: FP32 FMA issue rate on sm_100a for the operand forms the CLVO conv kernels can use.
//   mode 0: FFMA  R, R, R, R          (three register operands)
//   mode 1: FFMA2 R, R, R, R          (packed pair, fma.rn.f32x2)
//   mode 2: FFMA  R, R, c[0][imm], R  (weight from the constant bank = kernel parameter)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/fma_rate tools/fma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
struct W { float w[16]; };
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, const __grid_constant__ W wt, float seed) {
  float a[8], acc[16];
  float2 acc2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = i;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc2[i] = make_float2(i, i + 1);
  float b[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) b[i] = out[i] + seed;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(a[i & 7], b[i], acc[i]);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc2[i] = __ffma2_rn(make_float2(a[i], a[i]), make_float2(b[2 * i], b[2 * i + 1]), acc2[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(a[i & 7], wt.w[i], acc[i]);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc2[i].x + acc2[i].y;
  if (s == 12345.678f) out[threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* buf) {
  W wt;
  for (int i = 0; i < 16; ++i) wt.w[i] = 1.0f + i * 1e-3f;
  const int iters = 1 << 15, blocks = 148 * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(buf, iters, wt, 1e-9f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(buf, iters, wt, 1e-9f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)blocks * 256 * iters * 16;
  printf("%-28s %.3f ms  %.1f TFLOP/s  (%.1f FMA/clk/SM at 1.9 GHz)  %s\n", name, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.9e9,
         cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float* buf;
  cudaMalloc(&buf, 1 << 20);
  cudaMemset(buf, 0, 1 << 20);
  run<0>("FFMA 3-register", buf);
  run<1>("FFMA2 packed", buf);
  run<2>("FFMA constant-bank operand", buf);
  return 0;
}
