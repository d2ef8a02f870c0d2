// Host-side helpers shared by the tensor-core translation units: TMA tensor-map construction.
#pragma once
#include <string.h>

#include "common.h"

namespace atdn {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 4-D tensor map, dims innermost first, strides in ELEMENTS for dims 1..3 (byte strides must be multiples of
// 16), zero OOB fill (loads) / clipping (stores).  esize = bytes per element (1: bytes / e4m3, 2: fp16, 4: fp32).
inline int make_map(CUtensorMap* m, int esize, CUtensorMapSwizzle swz, const void* ptr, const int64_t dims[4],
                    const int64_t strides[3], const uint32_t box[4], const uint32_t estr[4], const char* what) {
  EncodeTiledFn fn = get_encode_fn();
  ATDN_REQUIRE(fn != nullptr, ATDN_ERR_ARCH, "cuTensorMapEncodeTiled is not available from the driver");
  ATDN_REQUIRE(ptr != nullptr && aligned16(ptr), ATDN_ERR_ALIGN, "%s: pointer must be non-null and 16-byte aligned", what);
  ATDN_REQUIRE(esize == 1 || esize == 2 || esize == 4, ATDN_ERR_ARG, "%s: element size %d", what, esize);
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < 4; ++i) {
    ATDN_REQUIRE(dims[i] >= 1, ATDN_ERR_ARG, "%s: dims[%d] = %lld", what, i, (long long)dims[i]);
    gd[i] = (cuuint64_t)dims[i];
    bx[i] = box[i];
    es[i] = estr[i];
  }
  for (int i = 0; i < 3; ++i) {
    ATDN_REQUIRE(strides[i] > 0 && (strides[i] * esize) % 16 == 0, ATDN_ERR_ALIGN,
                 "%s: strides[%d] = %lld elements is not a positive multiple of 16 bytes", what, i, (long long)strides[i]);
    gs[i] = (cuuint64_t)strides[i] * (cuuint64_t)esize;
  }
  CUresult r = fn(m, esize == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                  const_cast<void*>(ptr), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ATDN_REQUIRE(r == CUDA_SUCCESS, (int)r, "%s: cuTensorMapEncodeTiled failed with CUresult %d", what, (int)r);
  return 0;
}

// fp16 tensor, 128B swizzle (the operand layout of every tcgen05 kernel here)
inline int make_map_f16(CUtensorMap* m, const void* ptr, const int64_t dims[4], const int64_t strides[3],
                        const uint32_t box[4], const uint32_t estr[4], const char* what) {
  return make_map(m, 2, CU_TENSOR_MAP_SWIZZLE_128B, ptr, dims, strides, box, estr, what);
}


// Launch with programmatic stream serialization when ATDN_PDL is on (tc_ptx.cuh: pdl_launch_dependents / pdl_wait); the
// kernel must call pdl_wait() before its first global access.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = env_switches().pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// CTA-pair / persistent conv kernel entry (tc_conv.cu); called by atdn_tc_gemm when desc->mt > 0
int launch_conv_halo(const atdn_tc_desc* d, cudaStream_t stream);

}  // namespace atdn
