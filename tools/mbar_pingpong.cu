// Microbenchmark: mbarrier producer/consumer ring latency on sm_100a (platform characterisation).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void binit(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void barrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ bool btry(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
  return ok;
}
__device__ __forceinline__ bool btest(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
  return ok;
}
template <int MODE>
__device__ __forceinline__ void bwait(uint64_t* b, uint32_t par) {
  if (MODE == 0) { while (!btry(b, par)) {} }
  else { while (!btest(b, par)) {} }
}
template <int STAGES, int MODE>
__global__ void pingpong_fence(int iters, int prod_warp, int cons_warp, long long* out) {
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { binit(&full[s], 1); binit(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (warp == prod_warp && lane == 0) {
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) { bwait<0>(&empty[st], ph ^ 1); barrive(&full[st]); if (++st == STAGES) { st = 0; ph ^= 1; } }
  } else if (warp == cons_warp && lane == 0) {
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      bwait<0>(&full[st], ph);
      if (MODE == 1) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (MODE == 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&empty[st])) : "memory");
      else barrive(&empty[st]);
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
template <int STAGES, int MODE>
__global__ void pingpong(int iters, int prod_warp, int cons_warp, long long* out) {
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { binit(&full[s], 1); binit(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (warp == prod_warp && lane == 0) {
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) { bwait<MODE>(&empty[st], ph ^ 1); barrive(&full[st]); if (++st == STAGES) { st = 0; ph ^= 1; } }
  } else if (warp == cons_warp && lane == 0) {
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) { bwait<MODE>(&full[st], ph); barrive(&empty[st]); if (++st == STAGES) { st = 0; ph ^= 1; } }
    out[blockIdx.x] = clock64() - t0;
  }
}

// Ping-pong with the rest of the CTA behaving like the GEMM kernel's idle roles:
//   BG = 1: warps 0-3 sleep in a named hardware barrier that is released at the end
//   BG = 2: BG 1 + thread 0 polls a never-completing mbarrier with try_wait + nanosleep(40) (the accumulator wait)
//   BG = 3: all 128 threads of warps 0-3 poll the never-completing mbarrier with try_wait (no sleep)
//   BG = 4: thread 0 polls it with try_wait only (no sleep), the others sleep in the named barrier
//   lanes 1-31 of the producer / consumer warps wait at the final __syncthreads like in the GEMM kernel.
template <int STAGES, int BG>
__global__ void pingpong_bg(int iters, int prod_warp, int cons_warp, long long* out) {
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], never;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { binit(&full[s], 1); binit(&empty[s], 1); }
    binit(&never, 1);
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (warp == prod_warp) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < iters; ++i) { bwait<0>(&empty[st], ph ^ 1); barrive(&full[st]); if (++st == STAGES) { st = 0; ph ^= 1; } }
    }
  } else if (warp == cons_warp) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < iters; ++i) { bwait<0>(&full[st], ph); barrive(&empty[st]); if (++st == STAGES) { st = 0; ph ^= 1; } }
      out[blockIdx.x] = clock64() - t0;
      barrive(&never);
    }
  } else {
    if (BG == 2 || BG == 4) {
      if (threadIdx.x == 0) { while (!btry(&never, 0)) { if (BG == 2) __nanosleep(40); } }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    } else if (BG == 3) {
      while (!btry(&never, 0)) {}
    } else if (BG == 1) {
      if (threadIdx.x == 0) { while (!btest(&never, 0)) { __nanosleep(2000); } }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  __syncthreads();
}
int main() {
  long long* d; cudaMalloc(&d, 1024 * sizeof(long long));
  long long h[4];
  const int iters = 2000;
  auto run = [&](const char* name, auto kern, int pw, int cw, int threads, int blocks) {
    kern<<<blocks, threads>>>(iters, pw, cw, d); cudaDeviceSynchronize();
    kern<<<blocks, threads>>>(iters, pw, cw, d); cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(long long), cudaMemcpyDeviceToHost);
    printf("%-44s prod_warp=%d cons_warp=%d blocks=%d: %.1f cycles/iter (%s)\n", name, pw, cw, blocks, (double)h[0] / iters, cudaGetErrorString(cudaGetLastError()));
  };
  run("try_wait stages=3", pingpong<3, 0>, 4, 5, 192, 1);
  run("try_wait stages=3", pingpong<3, 0>, 0, 1, 64, 1);
  run("try_wait stages=3 same SMSP", pingpong<3, 0>, 0, 4, 192, 1);
  run("test_wait stages=3", pingpong<3, 1>, 4, 5, 192, 1);
  run("try_wait stages=1", pingpong<1, 0>, 4, 5, 192, 1);
  run("test_wait stages=1", pingpong<1, 1>, 4, 5, 192, 1);
  run("try_wait stages=6", pingpong<6, 0>, 4, 5, 192, 1);
  run("try_wait + tcgen05.fence::after", pingpong_fence<3, 1>, 4, 5, 192, 1);
  run("try_wait, consumer arrives via tcgen05.commit", pingpong_fence<3, 2>, 4, 5, 192, 1);
  run("try_wait stages=3 296 blocks", pingpong<3, 0>, 4, 5, 192, 296);
  run("test_wait stages=3 296 blocks", pingpong<3, 1>, 4, 5, 192, 296);
  run("bg1 named-barrier sleepers", pingpong_bg<3, 1>, 4, 5, 192, 1);
  run("bg2 + thread0 try_wait+nanosleep", pingpong_bg<3, 2>, 4, 5, 192, 1);
  run("bg3 128 threads try_wait", pingpong_bg<3, 3>, 4, 5, 192, 1);
  run("bg4 thread0 try_wait no sleep", pingpong_bg<3, 4>, 4, 5, 192, 1);
  run("bg2 296 blocks", pingpong_bg<3, 2>, 4, 5, 192, 296);
  run("bg3 296 blocks", pingpong_bg<3, 3>, 4, 5, 192, 296);
  return 0;
}
