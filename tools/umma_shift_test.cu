// Platform characterisation for the halo-reuse implicit GEMM: does a SWIZZLE_128B K-major UMMA shared-memory
// descriptor accept (a) a start address that is a multiple of 128 B but not of 1024 B (a row shift inside the
// swizzle atom) and (b) a stride-byte-offset (8-row group pitch) that is not a multiple of 1024 B, when the data
// were written by TMA with the absolute-address 128B swizzle?  Prints one line per (shift, SBO, base_offset mode).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../atdn_vslam_b200/csrc umma_shift_test.cu -o umma_shift_test
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_ptx.cuh"

using namespace atdn;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kARows = 384;   // rows of A staged in shared memory (48 KiB)
constexpr int kBRows = 128;

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7u) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128) shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                    int shift_rows, int sbo_bytes, int base_mode, float* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, done;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kARows * 128;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_smem, 128);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, (kARows + kBRows) * 128);
    // A in two boxes of 192 rows (box dims are limited to 256)
    tma_load_4d(sA, &tmA, &bar, 0, 0, 0, 0);
    tma_load_4d(sA + 192 * 128, &tmA, &bar, 0, 192, 0, 0);
    tma_load_4d(sB, &tmB, &bar, 0, 0, 0, 0);
    mbar_wait(&bar, 0);
    tcgen05_fence_after();
    const uint32_t a_addr = smem_u32(sA) + shift_rows * 128;
    const uint32_t bo = base_mode ? ((a_addr >> 7) & 7u) : 0u;
    const uint64_t a_desc = make_desc(a_addr, sbo_bytes, bo);
    const uint64_t b_desc = make_desc(smem_u32(sB), 1024, 0);
    for (int k = 0; k < 4; ++k) umma_f16(tmem_base, a_desc + 2u * k, b_desc + 2u * k, make_idesc_f16(128, 128), k > 0);
    umma_commit(&done);
  }
  mbar_wait(&done, 0);
  tcgen05_fence_after();
  const int row = threadIdx.x;
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[row * 128 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) != cudaSuccess || !fnp) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fnp);
  std::vector<__half> hA(kARows * 64), hB(kBRows * 64);
  std::vector<float> fA(kARows * 64), fB(kBRows * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 7 - 3); hA[i] = __float2half(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 5 - 2); hB[i] = __float2half(fB[i]); }
  __half *dA, *dB;
  float* dO;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dO, 128 * 128 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  {
    cuuint64_t gd[4] = {64, (cuuint64_t)kARows, 1, 1}, gs[3] = {128, (cuuint64_t)kARows * 128, (cuuint64_t)kARows * 128};
    cuuint32_t bx[4] = {64, 192, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dA, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A %d\n", (int)r); return 1; }
    cuuint64_t gdb[4] = {64, (cuuint64_t)kBRows, 1, 1}, gsb[3] = {128, (cuuint64_t)kBRows * 128, (cuuint64_t)kBRows * 128};
    cuuint32_t bxb[4] = {64, 128, 1, 1};
    r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dB, gdb, gsb, bxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B %d\n", (int)r); return 1; }
  }
  const int smem = (kARows + kBRows) * 128 + 1024;
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hO(128 * 128);
  const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 16, 19};
  const int sbos[] = {1024, 1280, 1536, 2048, 2560};
  for (int base_mode = 0; base_mode < 2; ++base_mode)
    for (int sbo : sbos)
      for (int shift : shifts) {
        cudaMemset(dO, 0xff, 128 * 128 * 4);
        shift_kernel<<<1, 128, smem>>>(tmA, tmB, shift, sbo, base_mode, dO);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("shift=%d sbo=%d base_mode=%d: CUDA error %s\n", shift, sbo, base_mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, bad_rows = 0;
        for (int i = 0; i < 128; ++i) {
          const int src = shift + (i / 8) * (sbo / 128) + (i % 8);
          int row_bad = 0;
          for (int n = 0; n < 128; ++n) {
            float acc = 0.f;
            if (src < kARows) for (int k = 0; k < 64; ++k) acc += fA[src * 64 + k] * fB[n * 64 + k];
            if (acc != hO[i * 128 + n]) ++row_bad;
          }
          bad += row_bad;
          bad_rows += row_bad > 0;
        }
        printf("%s shift=%2d sbo=%4d base_offset=%s: mismatches=%d rows_bad=%d\n", bad ? "FAIL" : "PASS", shift, sbo,
               base_mode ? "(addr>>7)&7" : "0", bad, bad_rows);
      }
  return 0;
}
