// Microbenchmarks behind the round-2 pyramid layout decision:
//  1. HBM WRITE bandwidth of the corr-pyramid store pattern: every CTA owns 128 queries and emits, tile after tile, one
//     CHUNK-byte piece per query.  query-major [q][tile][CHUNK] (pieces 30*CHUNK bytes apart) vs tile-major
//     [tile][q][CHUNK] (the 128 pieces of a CTA and tile are one contiguous 128*CHUNK-byte run).
//  2. HBM READ cost of gathering isolated 32-byte sectors / 64-byte sector pairs / 128-byte lines at random places, under
//     cudaLimitMaxL2FetchGranularity = 32 / 64 / 128 (does the L2 fetch more than the sector that was asked for?).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/mem_pattern_bench tools/mem_pattern_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int TILES = 30;

// 256 threads = 8 warps; warp w owns queries 16w..16w+15 of the CTA's 128; one instruction stores 32 lanes x 16 B = 512 B
template <int CHUNK>
__global__ void __launch_bounds__(256) wr_pattern(char* out, int tile_major, long long nq) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q0 = (long long)blockIdx.x * 128 + warp * 16;
  const uint4 val = make_uint4(lane, warp, blockIdx.x, 7);
  constexpr int LPQ = CHUNK / 16 < 32 ? CHUNK / 16 : 32;   // lanes per query piece
  constexpr int QPI = 32 / LPQ;                            // queries per instruction
  constexpr int IPQ = CHUNK / (LPQ * 16);                  // instructions per piece
  for (int t = 0; t < TILES; ++t) {
    for (int qb = 0; qb < 16; qb += QPI) {
      const long long q = q0 + qb + lane / LPQ;
      char* dst = tile_major ? out + ((long long)t * nq + q) * CHUNK : out + (q * TILES + t) * CHUNK;
#pragma unroll
      for (int k = 0; k < IPQ; ++k) *reinterpret_cast<uint4*>(dst + k * LPQ * 16 + (lane % LPQ) * 16) = val;
    }
  }
}

// every thread reads GRAN bytes at a pseudo-random GRAN-aligned place; sum kept alive
template <int GRAN>
__global__ void __launch_bounds__(256) rd_gather(const char* in, unsigned long long units, unsigned long long* sink, int per_thread) {
  unsigned long long x = (blockIdx.x * 256ull + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
  unsigned int acc = 0;
  for (int i = 0; i < per_thread; ++i) {
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    const unsigned long long u = (x * 0x2545F4914F6CDD1Dull) % units;
    const uint4* p = reinterpret_cast<const uint4*>(in + u * GRAN);
#pragma unroll
    for (int k = 0; k < GRAN / 16; ++k) { const uint4 v = __ldg(p + k); acc += v.x ^ v.y ^ v.z ^ v.w; }
  }
  if (acc == 0x12345678u) *sink = acc;
}

template <class F>
static float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  const int ctas = 57 * 54;
  const long long nq = (long long)ctas * 128;
  char* buf;
  const size_t bytes = (size_t)nq * TILES * 512;   // 6 GB
  if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(buf, 1, bytes);
  for (int tm = 0; tm < 2; ++tm) {
    float ms = time_ms([&] { wr_pattern<512><<<ctas, 256>>>(buf, tm, nq); });
    printf("write CHUNK 512 %s: %.3f ms %.0f GB/s\n", tm ? "tile-major [tile][q]" : "query-major [q][tile]", ms, (double)nq * TILES * 512 / ms / 1e6);
    ms = time_ms([&] { wr_pattern<128><<<ctas, 256>>>(buf, tm, nq); });
    printf("write CHUNK 128 %s: %.3f ms %.0f GB/s\n", tm ? "tile-major [tile][q]" : "query-major [q][tile]", ms, (double)nq * TILES * 128 / ms / 1e6);
    ms = time_ms([&] { wr_pattern<32><<<ctas, 256>>>(buf, tm, nq); });
    printf("write CHUNK  32 %s: %.3f ms %.0f GB/s\n", tm ? "tile-major [tile][q]" : "query-major [q][tile]", ms, (double)nq * TILES * 32 / ms / 1e6);
  }
  unsigned long long* sink;
  cudaMalloc(&sink, 8);
  const int grans[3] = {32, 64, 128};
  for (int gi = 0; gi < 3; ++gi) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, grans[gi]);
    size_t got = 0;
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    const int blocks = 148 * 64, per = 64;
    const double n = (double)blocks * 256 * per;
    float ms = time_ms([&] { rd_gather<32><<<blocks, 256>>>(buf, bytes / 32, sink, per); });
    printf("L2 fetch granularity limit %zu (%s): gather 32 B: %.3f ms  %.0f M units/s  %.0f GB/s useful\n", got, cudaGetErrorString(e), ms, n / ms / 1e3, n * 32 / ms / 1e6);
    ms = time_ms([&] { rd_gather<64><<<blocks, 256>>>(buf, bytes / 64, sink, per); });
    printf("L2 fetch granularity limit %zu: gather 64 B: %.3f ms  %.0f M units/s  %.0f GB/s useful\n", got, ms, n / ms / 1e3, n * 64 / ms / 1e6);
    ms = time_ms([&] { rd_gather<128><<<blocks, 256>>>(buf, bytes / 128, sink, per); });
    printf("L2 fetch granularity limit %zu: gather 128 B: %.3f ms  %.0f M units/s  %.0f GB/s useful\n", got, ms, n / ms / 1e3, n * 128 / ms / 1e6);
  }
  return 0;
}
