// Microbenchmark: HBM write bandwidth as a function of the store pattern of the corr-pyramid epilogue.
// Each CTA owns 128 "query rows" of ROWB bytes (contiguous: one query's level-0 map); 8 warps x 16 rows.
//   mode 0: patch order of the real kernel: for tile (th, tw), for hl < 8: 128 B per row at (th*8+hl)*640 + tw*128
//   mode 1: linear, RUN bytes per row per step (RUN = 128, 256, 512, 1024, 2048): offset = step * RUN
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/write_pattern_bench tools/write_pattern_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ROWB = 47 * 640;   // 30080 B per query
__global__ void __launch_bounds__(256) wpat(float* out, int mode, int run) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  char* base = reinterpret_cast<char*>(out) + (size_t)blockIdx.x * 128 * ROWB;
  const uint4 val = make_uint4(lane, warp, blockIdx.x, 7);
  if (mode == 0) {
    for (int t = 0; t < 30; ++t) {
      const int th = t / 5, tw = t % 5;
      for (int hl = 0; hl < 8; ++hl) {
        const int h = th * 8 + hl;
        if (h >= 47) break;
        const size_t off = (size_t)h * 640 + tw * 128 + (lane & 7) * 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = warp * 16 + (lane >> 3) + 4 * i;
          *reinterpret_cast<uint4*>(base + (size_t)r * ROWB + off) = val;
        }
      }
    }
  } else {
    const int lanes_per_row = run / 16 < 32 ? run / 16 : 32;          // lanes covering one row per instruction
    const int rows_per_instr = 32 / lanes_per_row;
    const int instr_per_run = run / (lanes_per_row * 16);
    for (int off0 = 0; off0 + run <= ROWB; off0 += run) {
      for (int rb = 0; rb < 16; rb += rows_per_instr) {
        const int r = warp * 16 + rb + lane / lanes_per_row;
        for (int k = 0; k < instr_per_run; ++k)
          *reinterpret_cast<uint4*>(base + (size_t)r * ROWB + off0 + k * lanes_per_row * 16 + (lane % lanes_per_row) * 16) = val;
      }
    }
  }
}
int main() {
  const int ctas = 57 * 27;
  float* buf;
  const size_t bytes = (size_t)ctas * 128 * ROWB;
  cudaMalloc(&buf, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int cfg[][2] = {{0, 128}, {1, 128}, {1, 256}, {1, 512}, {1, 1024}, {1, 2048}};
  for (auto& c : cfg) {
    wpat<<<ctas, 256>>>(buf, c[0], c[1]);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) wpat<<<ctas, 256>>>(buf, c[0], c[1]);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double wr = c[0] == 0 ? (double)ctas * 128 * 47 * 640 : (double)ctas * 128 * (ROWB / c[1]) * c[1];
    printf("mode %d run %4d: %.3f ms  %.0f GB/s  (%s)\n", c[0], c[1], ms, wr / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
