// Fused epilogues shared by the tensor-core kernels (tc_gemm.cu, tc_conv.cu).  One call handles 32
// consecutive accumulator columns of ONE accumulator row (= one output pixel / GEMM row) held by one thread
// after tcgen05.ld.32x32b.x32.  Every global load of the chunk (bias, residual, recurrent state) is issued
// before the first store: the output may alias the inputs as far as the compiler knows, so interleaving
// loads and stores serialises one L2 round trip per 8 columns (measured: 15K cycles per 128x128 tile).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/atdn_b200.h"
#include "tc_ptx.cuh"

namespace atdn {

struct EpiParams {
  int n_valid, flags;
  float alpha;
  const float* bias;
  void* out;
  long long out_pitch, out_ch_off;
  const __half* resid;
  long long resid_pitch, resid_ch_off;
  float* h32;
  float* z32;
  __half* rh16;
  const float* aux32;
  const float* gamma;
  int img_w, img_h;   // FLOW: output image size (pixel coordinates of `pix`)
  uint8_t* out8;      // STORE16 of a ROWS GEMM: second copy of the output as two e4m3 planes, hi = e4m3(y), lo = e4m3(y - hi):
  int out8_rows;      //   [batch][2][out8_rows][out_pitch] bytes (out8_rows = GEMM rows per batch element)
};

// 1/(1+e^-x) and tanh through MUFU.EX2 + MUFU.RCP: ~1e-6 relative, far below the fp16 operand rounding (5e-4)
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) {
  const float xc = fminf(fmaxf(x, -15.0f), 15.0f);
  return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * xc));
}

__device__ __forceinline__ uint4 pack8_f16(const float* y) {
  __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
  __half2 h2 = __floats2half2_rn(y[4], y[5]), h3 = __floats2half2_rn(y[6], y[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2);
  u.w = *reinterpret_cast<uint32_t*>(&h3);
  return u;
}
__device__ __forceinline__ void unpack8_f16(const uint4& u, float* y) {
  float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  float2 c = __half22float2(*reinterpret_cast<const __half2*>(&u.z));
  float2 d = __half22float2(*reinterpret_cast<const __half2*>(&u.w));
  y[0] = a.x; y[1] = a.y; y[2] = b.x; y[3] = b.y; y[4] = c.x; y[5] = c.y; y[6] = d.x; y[7] = d.y;
}

// The fp32 recurrent state (h32: GRU hidden state master copy, z32: update gate) lives in a TILED layout so that the
// thread-per-accumulator-row epilogues access it coalesced: a warp of the halo kernel owns 32 pixels of one 16 x 8
// sub-tile (r = q * 32 + lane, pixel = (r >> 3, r & 7)), and with the natural [pixel][128] layout every one of its
// float4 accesses touched 32 different 512-byte rows (the GRU_Q epilogue ran 46K cycles per tile against 25K for the
// MMAs: L1 transaction bound).  Layout: float4 index = ((sub-tile * 4 + q) * 32 + c4) * 32 + lane for channels
// 4*c4 .. 4*c4+3, sub-tile = (batch * ceil(H/16) + h/16) * ceil(W/8) + w/8: one warp access = 512 contiguous bytes.
// Size: batch * ceil(H/16)*16 * ceil(W/8)*8 * 128 floats.
__device__ __forceinline__ long long state_index(int b, int h, int w, int img_h, int img_w) {
  const int th_n = (img_h + 15) >> 4, tw_n = (img_w + 7) >> 3;
  const int r = ((h & 15) << 3) | (w & 7);
  return ((((static_cast<long long>(b) * th_n + (h >> 4)) * tw_n + (w >> 3)) * 4 + (r >> 5)) * 32) * 32 + (r & 31);
}

// fp16 store of `ng` complete 8-column groups followed by `tail` (< 8) single columns
__device__ __forceinline__ void store_row_f16(__half* dst, const float (&y)[32], int ng, int tail) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    if (g < ng) reinterpret_cast<uint4*>(dst)[g] = pack8_f16(&y[g * 8]);
  if (tail > 0) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g == ng) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < tail) dst[g * 8 + j] = __float2half_rn(y[g * 8 + j]);
      }
  }
}

// 32 channels (c0 .. c0+31) of one pixel of a tiled state buffer, fp32 or fp16 storage
__device__ __forceinline__ void load_tiled(const void* base, bool half, long long sidx, int c0, float4 (&o)[8]) {
  if (half) {
    const uint2* s = reinterpret_cast<const uint2*>(base) + sidx + (c0 >> 2) * 32;
    uint2 u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = s[i * 32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u[i].x));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u[i].y));
      o[i] = make_float4(a.x, a.y, b.x, b.y);
    }
  } else {
    const float4* s = reinterpret_cast<const float4*>(base) + sidx + (c0 >> 2) * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = s[i * 32];
  }
}
__device__ __forceinline__ void store_tiled(void* base, bool half, long long sidx, int c0, const float (&y)[32]) {
  if (half) {
    uint2* d = reinterpret_cast<uint2*>(base) + sidx + (c0 >> 2) * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __half2 a = __floats2half2_rn(y[4 * i], y[4 * i + 1]), b = __floats2half2_rn(y[4 * i + 2], y[4 * i + 3]);
      d[i * 32] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
  } else {
    float4* d = reinterpret_cast<float4*>(base) + sidx + (c0 >> 2) * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i * 32] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
  }
}

// y[0..31] += the stored pre-activation term of channels c0 .. c0+31 (tiled layout, fp32 or fp16)
__device__ __forceinline__ void add_pre_term(const EpiParams& p, long long elem_off, long long sidx, int c0, float (&y)[32]) {
  if (p.flags & ATDN_F_PRE16) {
    const uint2* pre = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.aux32) + elem_off) + sidx + (c0 >> 2) * 32;
    uint2 pv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pv[i] = __ldg(pre + i * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&pv[i].x));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&pv[i].y));
      y[4 * i] += a.x; y[4 * i + 1] += a.y; y[4 * i + 2] += b.x; y[4 * i + 3] += b.y;
    }
  } else {
    const float4* pre = reinterpret_cast<const float4*>(p.aux32 + elem_off) + sidx + (c0 >> 2) * 32;
    float4 pv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pv[i] = __ldg(pre + i * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i) { y[4 * i] += pv[i].x; y[4 * i + 1] += pv[i].y; y[4 * i + 2] += pv[i].z; y[4 * i + 3] += pv[i].w; }
  }
}

// sidx: state_index() of the pixel (4-channel-group units), or < 0 to derive it from pix = (b * img_h + h) * img_w + w
// st_row (STORE16, GRU_Q, r half of GRU_ZR): shared-memory address of this thread's 64-byte row in a 64B-swizzled
// staging box (the caller hands the box to a TMA store), or 0 to store to global memory directly; st_swz = (row >> 1) & 3.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& p, bool valid, long long pix, int n,
                                               const uint32_t (&v)[32], long long sidx = -1, uint32_t st_row = 0,
                                               uint32_t st_swz = 0) {
  if (n >= p.n_valid) return;
  if constexpr (EPI == ATDN_EPI_STORE16 || EPI == ATDN_EPI_STORE32 || EPI == ATDN_EPI_GRU_ZR || EPI == ATDN_EPI_GRU_Q) {
    if (sidx < 0 && valid && ((EPI != ATDN_EPI_STORE16 && EPI != ATDN_EPI_STORE32) || (p.flags & (ATDN_F_TANH_LO | ATDN_F_TILED32)))) {
      const int w = static_cast<int>(pix % p.img_w);
      const long long t = pix / p.img_w;
      sidx = state_index(static_cast<int>(t / p.img_h), static_cast<int>(t % p.img_h), w, p.img_h, p.img_w);
    }
  }
  const int nv = min(32, p.n_valid - n);   // valid columns of this chunk
  const int ng = nv >> 3, tail = nv & 7;
  float y[32];
  if (p.bias) {   // bias arrays are padded to a multiple of 64 floats
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = __ldg(b4 + i);
      y[4 * i + 0] = p.alpha * (__uint_as_float(v[4 * i + 0]) + b.x);
      y[4 * i + 1] = p.alpha * (__uint_as_float(v[4 * i + 1]) + b.y);
      y[4 * i + 2] = p.alpha * (__uint_as_float(v[4 * i + 2]) + b.z);
      y[4 * i + 3] = p.alpha * (__uint_as_float(v[4 * i + 3]) + b.w);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = p.alpha * __uint_as_float(v[j]);
  }
  if (!valid) return;

  if constexpr (EPI == ATDN_EPI_STORE16) {
    uint4 r4[4];
    const bool has_resid = (p.flags & ATDN_F_RESID) != 0;
    if (has_resid) {
      const uint4* rp = reinterpret_cast<const uint4*>(p.resid + pix * p.resid_pitch + p.resid_ch_off + n);
#pragma unroll
      for (int g = 0; g < 4; ++g) r4[g] = (g < ng) ? rp[g] : make_uint4(0, 0, 0, 0);
    }
    float2 fl = make_float2(0.f, 0.f);
    if (p.flags & ATDN_F_FLOWTAIL) fl = *reinterpret_cast<const float2*>(p.aux32 + pix * 2);
    if (p.flags & ATDN_F_TANH_LO) {
      if (n < 128) {
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = tanh_fast(y[j]);
        store_tiled(p.h32, (p.flags & ATDN_F_H16) != 0, sidx, n, y);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.0f);
      }
    } else if (p.flags & ATDN_F_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.0f);
    }
    if (p.flags & ATDN_F_FLOWTAIL) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (n + j == p.n_valid - 2) y[j] = fl.x;
        if (n + j == p.n_valid - 1) y[j] = fl.y;
      }
    }
    if (has_resid) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float r[8];
        unpack8_f16(r4[g], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[g * 8 + j] = fmaxf(r[j] + y[g * 8 + j], 0.0f);
      }
    }
    if (p.out8 != nullptr && n + 32 <= p.out_pitch) {   // columns past n_valid are zero accumulators: the pad stays finite
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t a = pack2_e4m3(y[4 * i], y[4 * i + 1]), b = pack2_e4m3(y[4 * i + 2], y[4 * i + 3]);
        const float2 fa = unpack2_e4m3(a), fb = unpack2_e4m3(b);
        hi[i] = a | (b << 16);
        lo[i] = pack2_e4m3(y[4 * i] - fa.x, y[4 * i + 1] - fa.y) | (pack2_e4m3(y[4 * i + 2] - fb.x, y[4 * i + 3] - fb.y) << 16);
      }
      const long long bidx = pix / p.out8_rows;
      uint8_t* d8 = p.out8 + (pix + bidx * p.out8_rows) * p.out_pitch + n;     // one 32-byte sector per plane and row
      st_global_v8(d8, hi[0], hi[1], hi[2], hi[3], hi[4], hi[5], hi[6], hi[7]);
      st_global_v8(d8 + static_cast<long long>(p.out8_rows) * p.out_pitch, lo[0], lo[1], lo[2], lo[3], lo[4], lo[5], lo[6], lo[7]);
    }
    if (st_row != 0) {
#pragma unroll
      for (uint32_t g = 0; g < 4; ++g) {
        const uint4 u = pack8_f16(&y[g * 8]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + ((g ^ st_swz) << 4)), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                     : "memory");
      }
    } else {
      store_row_f16(reinterpret_cast<__half*>(p.out) + pix * p.out_pitch + p.out_ch_off + n, y, ng, tail);
    }
  } else if constexpr (EPI == ATDN_EPI_STORE32) {
    if ((p.flags & ATDN_F_TILED32) && (p.flags & ATDN_F_PRE16)) {   // same tiled index space, fp16 storage
      uint2* t = reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + static_cast<long long>(n >> 7) * p.out_pitch) + sidx +
                 ((n & 127) >> 2) * 32;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __half2 a = __floats2half2_rn(y[4 * i], y[4 * i + 1]), b = __floats2half2_rn(y[4 * i + 2], y[4 * i + 3]);
        t[i * 32] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
      }
      return;
    }
    if (p.flags & ATDN_F_TILED32) {   // tiled recurrent-state layout, 128 channels per buffer (n_valid is a multiple of 32 here)
      float4* t = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<long long>(n >> 7) * p.out_pitch) + sidx +
                  ((n & 127) >> 2) * 32;
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i * 32] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
      return;
    }
    float* dst = reinterpret_cast<float*>(p.out) + pix * p.out_pitch + p.out_ch_off + n;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (g < ng) {
        reinterpret_cast<float4*>(dst)[2 * g] = make_float4(y[g * 8], y[g * 8 + 1], y[g * 8 + 2], y[g * 8 + 3]);
        reinterpret_cast<float4*>(dst)[2 * g + 1] = make_float4(y[g * 8 + 4], y[g * 8 + 5], y[g * 8 + 6], y[g * 8 + 7]);
      } else if (g == ng) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < tail) dst[g * 8 + j] = y[g * 8 + j];
      }
    }
  } else if constexpr (EPI == ATDN_EPI_GRU_ZR) {
    if (p.aux32) add_pre_term(p, n < 128 ? 0 : p.resid_pitch, sidx, n & 127, y);   // context part of the conv, once per pair (bias included)
    if (n < 128) {
      if (p.flags & ATDN_F_Z16) {
        uint2* z = reinterpret_cast<uint2*>(p.z32) + sidx + (n >> 2) * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const __half2 a = __floats2half2_rn(sigmoid_fast(y[4 * i]), sigmoid_fast(y[4 * i + 1]));
          const __half2 b = __floats2half2_rn(sigmoid_fast(y[4 * i + 2]), sigmoid_fast(y[4 * i + 3]));
          z[i * 32] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
        }
      } else {
        float4* z = reinterpret_cast<float4*>(p.z32) + sidx + (n >> 2) * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          z[i * 32] = make_float4(sigmoid_fast(y[4 * i]), sigmoid_fast(y[4 * i + 1]), sigmoid_fast(y[4 * i + 2]), sigmoid_fast(y[4 * i + 3]));
      }
    } else {
      float4 hv[8];
      load_tiled(p.h32, (p.flags & ATDN_F_H16) != 0, sidx, n - 128, hv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        y[4 * i + 0] = sigmoid_fast(y[4 * i + 0]) * hv[i].x;
        y[4 * i + 1] = sigmoid_fast(y[4 * i + 1]) * hv[i].y;
        y[4 * i + 2] = sigmoid_fast(y[4 * i + 2]) * hv[i].z;
        y[4 * i + 3] = sigmoid_fast(y[4 * i + 3]) * hv[i].w;
      }
      if (st_row != 0) {
#pragma unroll
        for (uint32_t g = 0; g < 4; ++g) {
          const uint4 u = pack8_f16(&y[g * 8]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + ((g ^ st_swz) << 4)), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                       : "memory");
        }
      } else {
        uint4* dst = reinterpret_cast<uint4*>(p.rh16 + pix * 128 + (n - 128));
#pragma unroll
        for (int g = 0; g < 4; ++g) dst[g] = pack8_f16(&y[g * 8]);
      }
    }
  } else if constexpr (EPI == ATDN_EPI_GRU_Q) {
    if (p.aux32) add_pre_term(p, 0, sidx, n, y);
    float4 hv[8], zv[8];
    load_tiled(p.h32, (p.flags & ATDN_F_H16) != 0, sidx, n, hv);
    load_tiled(p.z32, (p.flags & ATDN_F_Z16) != 0, sidx, n, zv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      y[4 * i + 0] = (1.0f - zv[i].x) * hv[i].x + zv[i].x * tanh_fast(y[4 * i + 0]);
      y[4 * i + 1] = (1.0f - zv[i].y) * hv[i].y + zv[i].y * tanh_fast(y[4 * i + 1]);
      y[4 * i + 2] = (1.0f - zv[i].z) * hv[i].z + zv[i].z * tanh_fast(y[4 * i + 2]);
      y[4 * i + 3] = (1.0f - zv[i].w) * hv[i].w + zv[i].w * tanh_fast(y[4 * i + 3]);
    }
    store_tiled(p.h32, (p.flags & ATDN_F_H16) != 0, sidx, n, y);
    if (st_row != 0) {
#pragma unroll
      for (uint32_t g = 0; g < 4; ++g) {
        const uint4 u = pack8_f16(&y[g * 8]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + ((g ^ st_swz) << 4)), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                     : "memory");
      }
    } else {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + pix * p.out_pitch + p.out_ch_off + n);
#pragma unroll
      for (int g = 0; g < 4; ++g) dst[g] = pack8_f16(&y[g * 8]);
    }
  } else if constexpr (EPI == ATDN_EPI_FLOW) {
    if (n == 0) {
      float2* c1 = reinterpret_cast<float2*>(p.h32) + pix;
      const float2 c = *c1;
      const int xw = static_cast<int>(pix % p.img_w);
      const int yh = static_cast<int>((pix / p.img_w) % p.img_h);
      const float cxn = c.x + y[0], cyn = c.y + y[1];
      *c1 = make_float2(cxn, cyn);
      reinterpret_cast<float2*>(p.z32)[pix] = make_float2(cxn - static_cast<float>(xw), cyn - static_cast<float>(yh));
    }
  } else if constexpr (EPI == ATDN_EPI_PV) {
    const float scale = p.aux32[pix] * __ldg(p.gamma);
    const uint4* rp = reinterpret_cast<const uint4*>(p.resid + pix * p.resid_pitch + p.resid_ch_off + n);
    uint4 r4[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) r4[g] = (g < ng) ? rp[g] : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float r[8];
      unpack8_f16(r4[g], r);
#pragma unroll
      for (int j = 0; j < 8; ++j) y[g * 8 + j] = r[j] + scale * y[g * 8 + j];
    }
    store_row_f16(reinterpret_cast<__half*>(p.out) + pix * p.out_pitch + p.out_ch_off + n, y, ng, tail);
  }
}

}  // namespace atdn
