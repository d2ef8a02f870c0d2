// Microbenchmark: throughput of TMA stores of [32 rows][128 B] boxes (the epilogue store unit of the corr-pyramid and
// attention kernels) as a function of WHERE the 32 rows land in global memory:
//   mode 0: rows ROWSTRIDE bytes apart (one row per query: 15 KiB apart in the pyramid, 14.6 KiB in P)
//   mode 1: the 32 rows of a box contiguous (4 KiB run)
// 8 warps per CTA, one CTA per SM, every warp owns two 4 KiB staging slots and issues stores back to back
// (cp.async.bulk.wait_group.read 1 before reusing a slot) -- no MMA, no epilogue arithmetic.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/tma_store_bench tools/tma_store_bench.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(256, 1) store_bench(const __grid_constant__ CUtensorMap tm, int boxes_per_warp, int rows_total, int contiguous) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* slot = smem + warp * 8192;
  for (int i = lane; i < 8192 / 4; i += 32) reinterpret_cast<uint32_t*>(slot)[i] = i * 2654435761u + warp;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const long long gw = (long long)blockIdx.x * 8 + warp;     // global warp id
  if (lane == 0) {
    for (int b = 0; b < boxes_per_warp; ++b) {
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      // tensor: dim0 = 64 halves (128 B), dim1 = column block (strided rows: tile index; contiguous: unused 1), dim2 = row
      const long long box = gw * boxes_per_warp + b;
      int c1, c2;
      if (contiguous) { c1 = 0; c2 = (int)((box * 32) % rows_total); }
      else { c1 = (int)(box % 120); c2 = (int)(((box / 120) * 32) % rows_total); }
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(slot + (b & 1) * 4096)), "r"(0), "r"(c1), "r"(c2) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// the same boxes LOADED (the A-operand stream of the P.V kernel): every warp keeps 4 boxes in flight on 4 mbarriers
__global__ void __launch_bounds__(256, 1) load_bench(const __grid_constant__ CUtensorMap tm, int boxes_per_warp, int rows_total, int contiguous) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[8][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* slot = smem + warp * 16384;
  if (lane == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[warp][i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const long long gw = (long long)blockIdx.x * 8 + warp;
  if (lane == 0) {
    for (int b = 0; b < boxes_per_warp + 4; ++b) {
      const int s4 = b & 3;
      if (b >= 4) {   // wait for the load issued 4 boxes ago into this slot
        const uint32_t parity = ((b - 4) >> 2) & 1;
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[warp][s4])), "r"(parity) : "memory");
      }
      if (b < boxes_per_warp) {
        const long long box = gw * boxes_per_warp + b;
        int c1, c2;
        if (contiguous) { c1 = 0; c2 = (int)((box * 32) % rows_total); }
        else { c1 = (int)(box % 120); c2 = (int)(((box / 120) * 32) % rows_total); }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[warp][s4])), "r"(4096) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(slot + s4 * 4096)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bars[warp][s4])), "r"(0), "r"(c1), "r"(c2) : "memory");
      }
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn encode = reinterpret_cast<EncodeFn>(fnp);
  const size_t bytes = 6ull << 30;
  char* buf;
  cudaMalloc(&buf, bytes);
  cudaFuncSetAttribute(store_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int contiguous = 0; contiguous < 2; ++contiguous) {
    // strided: [rows][120 column blocks][64 halves] = 15360-byte rows, a box takes one 128-byte piece of 32 consecutive rows;
    // contiguous: [rows][64 halves]
    const cuuint64_t rows = contiguous ? bytes / 128 : bytes / 15360;
    cuuint64_t dims[3] = {64, contiguous ? 1ull : 120ull, rows};
    cuuint64_t strides[2] = {128, contiguous ? 128ull : 15360ull};
    cuuint32_t box[3] = {64, 1, 32}, es[3] = {1, 1, 1};
    CUtensorMap tm;
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int boxes = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    store_bench<<<sms, 256, 65536 + 1024>>>(tm, boxes, (int)rows, contiguous);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    store_bench<<<sms, 256, 65536 + 1024>>>(tm, boxes, (int)rows, contiguous);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double total = (double)sms * 8 * boxes * 4096;
    printf("STORE %s rows: %.3f ms  %.0f GB/s  (%.1f B/clk/SM at 1.9 GHz)  %s\n", contiguous ? "contiguous" : "15 KiB-strided", ms, total / ms / 1e6,
           total / sms / (ms * 1e-3) / 1.9e9, cudaGetErrorString(cudaGetLastError()));
    cudaFuncSetAttribute(load_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
    load_bench<<<sms, 256, 131072 + 1024>>>(tm, boxes, (int)rows, contiguous);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    load_bench<<<sms, 256, 131072 + 1024>>>(tm, boxes, (int)rows, contiguous);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("LOAD  %s rows: %.3f ms  %.0f GB/s  (%.1f B/clk/SM at 1.9 GHz)  %s\n", contiguous ? "contiguous" : "15 KiB-strided", ms, total / ms / 1e6,
           total / sms / (ms * 1e-3) / 1.9e9, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
