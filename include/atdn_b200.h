/*
 * atdn_b200.h -- C ABI of libatdn_b200.so: hand-written sm_100a kernels for the ATDN vSLAM odometry
 * front end (GMA flow -> CLVO pose -> keyframe search).
 *
 * The reference (MILAB-IIT-CV/ATDN_vSLAM) is pure Python/PyTorch and has NO FFI of its own
 * (SURVEY.md section 2.2); every entry point below therefore cites the reference *Python* call site
 * whose ATen/cuDNN/cuBLAS dispatch it replaces.  GMA.whl!/ = inside GMA-1.0.0-py3-none-any.whl.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *  - the caller owns all memory (inputs, outputs, workspaces); the library never allocates device
 *    memory and keeps no reference after the call returns;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises;
 *  - return value: 0 = ok, < 0 = argument/shape/alignment error detected before launch,
 *    > 0 = cudaError_t / CUresult of the failed runtime call.  atdn_last_error() returns a
 *    thread-local, NUL-terminated description of the last non-zero return;
 *  - there is no CPU fallback: on a device that is not sm_100 every compute call returns
 *    ATDN_ERR_ARCH.
 *  - activation tensors are NHWC fp16 ("pixel rows"); `pitch` = elements between consecutive
 *    pixels; channel slices of a wider buffer are addressed with (pointer + channel offset, pitch).
 */
#ifndef ATDN_B200_H_
#define ATDN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATDN_ERR_ARG   (-1)
#define ATDN_ERR_ALIGN (-2)
#define ATDN_ERR_ARCH  (-3)
#define ATDN_ERR_UNSUP (-4)

const char* atdn_last_error(void);
int atdn_version(void);
/* 0 if `device` is an sm_100 part, ATDN_ERR_ARCH otherwise (also primes the driver entry points). */
int atdn_check_device(int device);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core implicit GEMM (tcgen05.mma kind::f16, fp32 accumulation in TMEM, operands staged by
 * TMA with 128B swizzle).  D[m, n] = sum_k A[m, k] * B[n, k].
 *
 * One kernel family serves every dense contraction on the path:
 *   convolutions   GMA.whl!/GMA/core/extractor.py:47-55,173-181 (encoders), update.py:76-84 (motion
 *                  encoder), :48-63 (SepConvGRU), :7-15,120-123 (flow / mask heads), gma.py:59 (to_qk),
 *                  gma.py:105 (to_v)                                      -- replaces cuDNN conv
 *   corr volume    GMA.whl!/GMA/core/corr.py:55-63 + pyramid :28-30      -- replaces cuBLAS SGEMM + avg_pool2d
 *   attention      GMA.whl!/GMA/core/gma.py:72 (q k^T), :107 (attn v)     -- replaces cuBLAS batched GEMM
 *
 * Operand addressing ("modes"), all fp16, innermost dimension first:
 *   ROWS : dims {K, rows, 1, batch}            one tile = 128 (A) or `bn` (B) consecutive rows
 *   PATCH: dims {C, W, H, batch} (NHWC image)  one A tile = 8 x 16 output pixels, one K step per
 *          (tap, 64-channel chunk); out-of-image taps are zero-filled by TMA (= conv zero padding).
 *          For B (corr volume only) one tile = 8 x 32 target pixels.
 * strides[] are in ELEMENTS for dims 1..3 and must be multiples of 8 (16 bytes); pointers must be
 * 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
enum { ATDN_MODE_ROWS = 0, ATDN_MODE_PATCH = 1 };

enum {
  ATDN_EPI_STORE16 = 0, /* out16[pix, ch_off+n] = act(alpha*acc + bias[n]) (+ optional residual / flow tail / tanh split) */
  ATDN_EPI_STORE32 = 1, /* out32[pix, ch_off+n] = alpha*acc + alpha*bias[n]        (fp32)                                */
  ATDN_EPI_CORR    = 2, /* corr pyramid: level0 = alpha*acc, levels 1..3 = 2x2 means, fp32 (corr.py:16-30)              */
  ATDN_EPI_GRU_ZR  = 3, /* n<128: z32 = sigmoid(acc+b); n>=128: rh16 = sigmoid(acc+b) * h32   (update.py:51-53,58-60)   */
  ATDN_EPI_GRU_Q   = 4, /* q = tanh(acc+b); h32 = (1-z)*h32 + z*q; out16 = h32               (update.py:53-54,60-61)   */
  ATDN_EPI_PV      = 5, /* out16 = resid16 + gamma * acc * row_scale[pix]                    (gma.py:107-113)           */
  ATDN_EPI_FLOW    = 6  /* halo kernel only, n_valid = 2 (weights zero-padded to 32 rows): delta = acc + bias;
                           coords1[pix] += delta; flow[pix] = coords1[pix] - (x, y); coords1 = h32, flow = z32
                           (update.py:14 flow_head.conv2 + network.py:111,116)                                           */
};

enum {
  ATDN_F_RELU      = 1,  /* STORE16: relu after bias                                                      */
  ATDN_F_RESID     = 2,  /* STORE16: y = relu(resid16[pix, n] + y)   (extractor.py:55)                     */
  ATDN_F_FLOWTAIL  = 4,  /* STORE16: columns n >= n_valid-2 take aux32[pix*2 + (n - (n_valid-2))] (update.py:84) */
  ATDN_F_TANH_LO   = 8,  /* STORE16: n < 128 -> tanh (also written to h32), n >= 128 -> relu (network.py:95-97) */
  ATDN_F_B_BATCHED = 16, /* B operand has a batch dimension (attention GEMMs, corr volume)               */
  ATDN_F_A_SHARED  = 32, /* ROWS A is shared by all batches (weights as the A operand: transposed output); batch = b_dims[3] */
  ATDN_F_PAIR      = 64, /* CTA-pair kernel (tcgen05 cta_group::2): 256 x bn tiles, each CTA stages bn/2 B rows; bn up to 256 */
  ATDN_F_TILED32   = 256, /* STORE32: fp32 output in the tiled recurrent-state layout (see h32 below), 128 channels per buffer:
                            channel n goes to buffer n / 128 at out + (n / 128) * out_pitch floats (out_pitch = floats per buffer).
                            Used once per pair for the part of the SepConvGRU convolutions that only sees the context
                            features (constant over the refinement iterations); bias included                          */
  ATDN_F_PRE16     = 512, /* the pre-activation term is fp16: with ATDN_F_TILED32 the buffers are written as fp16 (same tiled
                            index space, out_pitch in elements); with GRU_ZR / GRU_Q aux32 points to fp16 (resid_pitch in
                            elements).  Halves the per-iteration epilogue traffic of the context term               */
  ATDN_F_Z16       = 1024, /* GRU_ZR / GRU_Q: z32 points to fp16 (same tiled index space): the update gate lies in (0, 1), its fp16
                             rounding (2.4e-4 absolute) is below the fp16 rounding of the hidden state it blends        */
  ATDN_F_H16       = 2048, /* h32 points to fp16 (same tiled index space): the hidden state has no fp32 master copy, as in the
                             reference's own fp16-autocast path (GRU_ZR / GRU_Q / STORE16|TANH_LO)                      */
  ATDN_F_A_TILED   = 4096, /* ROWS A (single-CTA kernel) is stored in blocks of 32 rows x 64 columns, [batch][ceil(rows/32)][cols/64][32][64]
                              (the layout atdn_attn_probs writes with p_tiled != 0: its 32 x 64 store boxes and the 128 x 64 operand
                              boxes read here are contiguous 4 KiB runs); a_dims = {cols, rows, 1, batch}, a_strides are ignored,
                              cols must be a multiple of 64 with every column block fully written (zeros past the true extent); also on CTA pairs */
  ATDN_F_A_MIXED   = 8192, /* with ATDN_F_PAIR | ATDN_F_A_TILED, ATDN_EPI_PV, bn = 128 or 64: the A blocks of 256 rows x 64 columns are fp16 or e4m3
                              per block as atdn_attn_probs (block_hot != NULL) + atdn_attn_harmonize left them (a_hot = the pair bitmap);
                              an e4m3 block is multiplied with the e4m3 copy of B (b8, two planes hi + lo, written by a STORE16 GEMM
                              with out8) by two tcgen05.mma.kind::f8f6f4 per 32 columns, into the same fp32 accumulator as the fp16 blocks */
  ATDN_F_STATS     = 128, /* STORE16 on the halo kernel with mt = 4, bn = 64 (n_valid = 64): per-channel partial sums of
                            (acc + bias) and its square over the in-image pixels each epilogue warp sees in one tile go to
                            aux32 as [batch, parts, 64, 2] fp32, parts = ceil(H/16) * ceil(W/(32*cl)) * cl * 4 (cl = 2 with
                            ATDN_F_PAIR): the instance-norm statistics pass without re-reading the tensor; reduce them with
                            atdn_inorm_finalize                                                                         */
};

typedef struct atdn_tc_desc {
  int32_t bn;             /* N tile: 64, 96, 128, 192 or 256 (256 only for ATDN_EPI_CORR)           */
  int32_t epi;            /* ATDN_EPI_*                                                             */
  int32_t flags;          /* ATDN_F_*                                                               */
  int32_t a_mode;         /* ATDN_MODE_*                                                            */
  int32_t b_mode;         /* ATDN_MODE_ROWS, or ATDN_MODE_PATCH for ATDN_EPI_CORR                   */
  int32_t n_valid;        /* number of valid output columns (Cout / N)                              */
  /* convolution geometry (PATCH A): out_h x out_w output pixels per image                          */
  int32_t out_h, out_w;
  int32_t taps_h, taps_w, pad_h, pad_w, stride;   /* stride 1 or 2                                  */
  int32_t a_split_chunk;  /* 64-channel chunks [0, split) of each tap come from `a`, the rest from
                             `a2` (concatenated inputs without a copy); 0 = `a` only                */
  const void* a;  int64_t a_dims[4];  int64_t a_strides[3];
  const void* a2; int64_t a2_dims[4]; int64_t a2_strides[3];
  const void* b;  int64_t b_dims[4];  int64_t b_strides[3];
  /* epilogue */
  float alpha;            /* scale applied to the accumulator                                       */
  const float* bias;      /* [n_valid] fp32 or NULL                                                 */
  void* out;              /* fp16 (STORE16, GRU_Q, PV) / fp32 (STORE32) / level-0 fp32 (CORR)       */
  int64_t out_pitch;      /* elements per pixel row                                                 */
  int64_t out_ch_off;
  const void* resid16;    /* STORE16|RESID, PV: fp16 [pix, resid_pitch] at resid_ch_off             */
  int64_t resid_pitch, resid_ch_off;
  /* fp32 recurrent state, 128 channels per pixel, in the TILED layout the epilogue warps access coalesced:
   *   float index = ((((b*ceil(H/16) + h/16)*ceil(W/8) + w/8)*4 + r/32)*32 + c/4)*128 + (r%32)*4 + c%4,
   *   r = (h%16)*8 + w%8  (H, W = out_h, out_w); size batch*ceil(H/16)*16*ceil(W/8)*8*128 floats.              */
  float* h32;             /* GRU_*: hidden state master copy; STORE16|TANH_LO: written (FLOW: coords1)   */
  float* z32;             /* GRU_ZR writes / GRU_Q reads the update gate            (FLOW: flow)         */
  void* rh16;             /* GRU_ZR: fp16 [pix,128] = r*h                                           */
  const float* aux32;     /* FLOWTAIL: fp32 flow [pix,2]; PV: row_scale [pix]; STATS: partial sums (written);
                             GRU_ZR / GRU_Q: optional pre-activation term in the tiled layout (ATDN_F_TILED32 output), added
                             to acc + bias; GRU_ZR reads the r half at aux32 + resid_pitch floats                 */
  const float* gamma;     /* PV: pointer to the scalar Aggregate.gamma                              */
  /* CORR: pyramid levels 1..3 (fp32) and their row pitches; level l is [batch*rows, H_l, pitch_l]  */
  float* lvl[3];
  int32_t lvl_pitch[4];   /* pitch of levels 0..3 in elements (multiples of 4)                      */
  int32_t corr_h, corr_w; /* target grid H8 x W8                                                    */
  /* mt > 0 selects the persistent halo-reuse convolution kernel (PATCH A, stride 1): one CTA owns
   * `mt` side-by-side sub-tiles of 16 x 8 output pixels, stages one (16+kh-1) x (8*mt+kw-1) input box per
   * 64-channel chunk and reads every filter tap as a shifted view of it; mt * bn <= 256 (two TMEM
   * accumulator buffers).  Instances: (mt, bn) = (1,256) (1,192) (1,128) (2,128) (2,96) (2,64) (4,64) (4,32);
   * epilogues STORE16, STORE32, GRU_ZR, GRU_Q, FLOW (bn 32).  mt = 0: one 8 x 16 pixel tile per CTA.       */
  int32_t mt;
  /* mixed-precision attention operands (ATDN_F_A_MIXED; all NULL otherwise) */
  void* out8;             /* STORE16 of a ROWS GEMM (mt = 0): also write the output as two e4m3 planes, hi = e4m3(y) and
                             lo = e4m3(y - hi): bytes [batch][2][rows][out_pitch]; columns [n_valid, out_pitch) of a started 32-column
                             chunk are written as zeros, later chunks are untouched (allocate zeroed)                  */
  const void* b8;         /* ATDN_F_A_MIXED: e4m3 planes of B, bytes [batch][2][b_dims[1]][b_strides[0]] (an out8 buffer)       */
  const uint8_t* a_hot;   /* ATDN_F_A_MIXED: [batch][ceil(rows/256)][cols/64], 1 = fp16 block, 0 = e4m3 block (atdn_attn_harmonize) */
} atdn_tc_desc;

int atdn_tc_gemm(const atdn_tc_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Correlation volume + pyramid -- GMA.whl!/GMA/core/corr.py:16-30 (CorrBlock.__init__) and :55-63
 * (CorrBlock.corr); replaces cuBLAS SGEMM + 3 x avg_pool2d.  fmap1 / fmap2: NHWC fp16 [batch, h8, w8, fmap_pitch]
 * (channels = 256).  lvl[l]: fp32 [batch * h8 * w8, H_l, lvl_pitch[l]] with H_l = h8 >> l, W_l = w8 >> l
 * (floor), lvl_pitch[l] >= W_l and a multiple of 4:
 *   lvl[0][q, y, x]   = alpha * <fmap1[q], fmap2[y, x]>          (alpha = 1 / sqrt(channels))
 *   lvl[l+1][q, y, x] = mean of the 2x2 block of lvl[l]           (hierarchical, fp32)
 * One persistent-style kernel: tcgen05 MMAs into two TMEM accumulators, fp32 boxes staged in shared memory
 * and written with TMA stores (full 128-byte runs per query row); columns [W_l, ceil4(W_l)) of a row may be
 * overwritten with pad values (16-byte store granularity).
 * half_levels = 0: all four levels fp32 (the reference's corr_pyramid, bit-for-bit layout of its values).
 * half_levels = 4: all levels are stored as fp16 in the STRIP layout: the target map is cut into tiles of 8 x 32 texels
 *   (tiles_w = ceil(w8/32), tiles = ceil(h8/8) * tiles_w) and lvl[l] = [batch*n, tiles, chunk_l], chunk = 256 / 64 / 16;
 *   lvl_pitch[l] must be the chunk size (256, 64, 16, 4).  Texel (y, x) of level l lives in tile
 *   (y >> (3-l)) * tiles_w + (x >> (5-l)) at chunk offset
 *     level 0: ((x >> 3) & 3) * 64 + (y & 7) * 8 + (x & 7)      [strip 4][row 8][col 8]
 *     level 1: ((x >> 3) & 1) * 32 + (y & 3) * 8 + (x & 7)      [strip pair 2][row 4][col 8]
 *     level 2:                      (y & 1) * 8 + (x & 7)       [row 2][col 8]
 *   and level 3 is lvl[3] = [batch*n, ceil(h8/8) * tiles_w3, 4] with tiles_w3 = tiles_w rounded up to even: texel (y, x) at
 *   (y * tiles_w3 + (x >> 2)) * 4 + (x & 3); the pad tile column is never written and must be zero (allocate it zeroed).
 *   Equivalently, at every level the half offset inside a query's maps is rowpart(y) + (x >> 3) * (64 >> l) + (x & 7) with
 *   rowpart(y) = (y >> (3-l)) * tiles_w * chunk_l + (y & ((8 >> l) - 1)) * 8: every 8-texel row piece is 16-byte aligned, and
 *   a 64-byte HBM fetch granule is a 4 x 8-texel block at levels 0 / 1 (the L2 of the B200 fetches 64-byte granules).
 *   Texels of a tile that lie outside the level's map (H_l x W_l) are written as ZEROS, so a reader can stage windows
 *   without masking.  Each level is pooled from the un-rounded fp32 values of the level below and rounded once.  The kernel
 *   is bound by its HBM stores, so fp16 halves its time (and the lookup's read traffic); the end-to-end flow moves by
 *   8e-4 px mean (tools/fp16_pyramid_sensitivity.py), inside the 1e-2 px bar.
 * ---------------------------------------------------------------------------------------------- */
int atdn_corr_pyramid(const void* fmap1, const void* fmap2, int64_t fmap_pitch, int32_t channels,
                      void* const lvl[4], const int32_t lvl_pitch[4], int32_t half_levels, int32_t batch,
                      int32_t h8, int32_t w8, float alpha, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused attention probabilities -- GMA.whl!/GMA/core/gma.py:66-73 (q k^T * scale, softmax over keys);
 * replaces cuBLAS batched GEMM + softmax.  qk16: fp16 [batch, n, qk_pitch] with q in channels 0..127 and k
 * in channels 128..255 (the to_qk 1x1 conv output).  The logits never reach HBM: every key tile goes
 * through tcgen05 twice (row maxima, then probabilities).
 *   p16[b, i, j]  = exp((q_i . k_j - max_j q_i . k_j) * scale)   fp16, un-normalised, j < n; columns [n, ceil8(n)) are
 *                   written as zeros (16-byte store granularity), the rest of the row pad is untouched
 *   inv_sum[b, i] = 1 / sum_j p16[b, i, j]                       fp32
 * p_tiled != 0: p16 is written in blocks of 32 rows x 64 columns, [batch][ceil(n/32)][p_pitch/64][32][64] (p_pitch a multiple
 *   of 64 >= n; columns [n, p_pitch) of the last written block are zeros): every store box is one contiguous 4 KiB run
 *   instead of 32 rows p_pitch apart (strided rows cost the TMA store path ~27%, tools/tma_store_bench.cu), and the P.V GEMM
 *   reads it with ATDN_F_A_TILED.
 * block_hot != NULL (needs p_tiled and p_pitch = ceil64(n)): MIXED storage.  Every value is stored scaled by 256 (inv_sum
 *   follows, so consumers see no difference), and each sub-block of 32 rows x 64 columns is either fp16 as above or e4m3:
 *   [32][64] BYTES in the first 2 KiB of its 4 KiB slot.  A sub-block stays fp16 ("hot", block_hot[b][i/32][j/64] = 1) when
 *   for one of its rows sqrt(sum_block p^2) > hot_energy * sum_row p, i.e. when the e4m3 rounding of the block could move
 *   that row's aggregate by more than ~hot_energy / 16 relative.  The row sums of the criterion are ESTIMATES from the first
 *   pass (online softmax over every 4th column, floored by the row maximum: exact for flat rows, within ~4x for rows carried
 *   by a few keys); inv_sum comes from the ROUNDED stored values as before.  block_hot: [batch][ceil(n/32)][p_pitch/64].
 * ---------------------------------------------------------------------------------------------- */
int atdn_attn_probs(const void* qk16, int64_t qk_pitch, void* p16, int64_t p_pitch, int32_t p_tiled, float* inv_sum,
                    int32_t batch, int32_t n, float scale, uint8_t* block_hot, float hot_energy, void* stream);

/* Second step of the mixed storage, for the CTA-pair P.V kernel (ATDN_F_PAIR | ATDN_F_A_TILED | ATDN_F_A_MIXED) whose MMAs span
 * 256 rows: pair_hot[b][i/256][j/64] = OR of the eight sub-block flags; where they disagree the e4m3 sub-blocks are rewritten
 * in place as fp16 (identical values: their rounding already happened).  P.V then reads ~(1 + hot fraction) / 2 of the fp16
 * bytes.  p16 / p_pitch / block_hot as passed to atdn_attn_probs; pair_hot: [batch][ceil(n/256)][p_pitch/64]. */
int atdn_attn_harmonize(void* p16, int64_t p_pitch, const uint8_t* block_hot, uint8_t* pair_hot, int32_t batch, int32_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Correlation lookup -- GMA.whl!/GMA/core/corr.py:32-53 + utils/utils.py:59-73 (grid_sample).
 * coords: fp32 [B, H8, W8, 2] (x, y); levels and half_levels (0 or 4) as written by atdn_corr_pyramid;
 * the fp16 path stages the windows with cp.async (zero fill outside the padded maps) and blends the 10x10 window
 * separably with one fractional offset per level (grid_sample's per-tap coordinate round trip, an ulp-level
 * perturbation, is dropped);
 * out: fp16 [B*H8*W8, out_pitch], channel = level*81 + a*9 + b  (a offsets x, b offsets y); out32 (optional): the
 * un-rounded fp32 results [B*H8*W8, 324].
 * ---------------------------------------------------------------------------------------------- */
int atdn_corr_lookup(const void* const lvl[4], const int32_t lvl_pitch[4], int32_t half_levels, const float* coords,
                     void* out16, int64_t out_pitch, float* out32_or_null,
                     int32_t batch, int32_t h8, int32_t w8, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Element-wise / data-movement kernels of the flow net
 * ---------------------------------------------------------------------------------------------- */
/* Caller-side resize of the SLAM loop -- neural_slam.py:197-199 (TF.resize: bilinear, antialias=True) -- replacing
 * torch.nn.functional.interpolate(..., antialias=True).  src: [planes, in_h, in_w] fp32 or uint8 (src_is_u8), dst: [planes, out_h,
 * out_w] fp32.  ATen's separable triangle filter: per axis scale = in / out, support = max(scale, 1), center = scale * (o + 0.5),
 * taps lo = max(int(center - support + 0.5), 0) .. min(int(center + support + 0.5), in), weight max(0, 1 - |(i - center + 0.5) /
 * max(scale, 1)|) normalised to sum 1; rows are filtered first.  Equal sizes on an axis give the identity on that axis.
 * Down-scaling by more than 5x is refused (ATDN_ERR_UNSUP). */
int atdn_resize_aa(const void* src, int32_t src_is_u8, float* dst, int32_t planes, int32_t in_h, int32_t in_w, int32_t out_h, int32_t out_w,
                   void* stream);
/* The two 7x7 convolutions on thin inputs (3-channel image, 2-channel flow) run on atdn_tc_gemm after their
 * HORIZONTAL taps have been folded into channels; the vertical taps stay implicit in the convolution.
 *
 * network.py:75-76 (2*(image/255)-1) + extractor.py:173 (conv1 7x7 stride 2 pad 3): image fp32 NCHW [B,3,H,W]
 * in 0..255 -> x16 NHWC fp16 [B, H/2, W/2, 48],
 *   x16[b, y2, ox, (ry*3 + c)*8 + xx] = 2*(image[b, c, 2*y2+ry, 2*(ox-2)+xx]/255) - 1   (0 outside the image),
 * to be convolved 4x1 (top padding 2) with W4[o, (ry*3+c)*8+xx, ai] = w[o, c, 2*ai+ry-1, xx-1] (0 outside 7x7). */
int atdn_stem_pack(const float* image, void* x16, int32_t batch, int32_t h, int32_t w, void* stream);
/* update.py:79 (convf1 7x7 pad 3 on the 2-channel flow): flow fp32 [B,H8,W8,2] -> x16 NHWC fp16 [B,H8,W8,16],
 *   x16[b, y, x, dx*2 + c] = flow[b, y, x+dx-3, c] (0 outside; channels 14, 15 = 0),
 * to be convolved 7x1 (padding 3) with W7[o, dx*2+c, dy] = w[o, c, dy, dx].                                  */
int atdn_flow_pack(const float* flow, void* x16, int32_t batch, int32_t h8, int32_t w8, void* stream);
/* nn.InstanceNorm2d (extractor.py:28-32,127) on NHWC fp16: per (image, channel) mean / rstd over H*W
 * into stats fp32 [B, C, 2]; two-stage, deterministic.  scratch: fp32 [B * parts * C * 2].          */
int atdn_inorm_stats(const void* x16, int64_t pitch, int32_t batch, int32_t hw, int32_t c,
                     float* scratch, int32_t parts, float* stats, void* stream);
/* second stage alone: scratch fp32 [batch, parts, c, 2] partial (sum, sum of squares) -> stats [batch, c, 2] = (mean, rstd);
 * the partials may come from atdn_tc_gemm epilogues (ATDN_F_STATS).  Fixed reduction order: deterministic.          */
int atdn_inorm_finalize(const float* scratch, int32_t parts, int32_t batch, int32_t c, int32_t hw, float* stats, void* stream);
/* y = relu((x - mean) * rstd); if resid16: y = relu(resid16 + y) (extractor.py:47-55).  In place ok. */
int atdn_inorm_apply(const void* x16, int64_t pitch, const float* stats, const void* resid16, int64_t resid_pitch,
                     void* y16, int64_t y_pitch, int32_t batch, int32_t hw, int32_t c, int32_t relu, void* stream);
/* update.py:14 flow_head.conv2 (3x3, 256 -> 2) + network.py:116 (coords1 += delta; flow = coords1 - coords0) from the per-tap
 * partial products of a 256 -> 18 1x1 convolution on atdn_tc_gemm (STORE32):
 * d32[pix, tap*2 + co] = sum_c w[co, c, tap] x[pix, c], tap = dy*3 + dx;  delta[p, co] = bias[co] + sum over the taps whose
 * neighbour p + (dy-1, dx-1) lies inside the image of d32[that neighbour, tap*2 + co].  d32: fp32 [B*H8*W8, pitch].        */
int atdn_flow_head_gather(const float* d32, int64_t pitch, const float* bias, float* coords1, float* flow,
                          int32_t batch, int32_t h8, int32_t w8, void* stream);
/* network.py:59-70 convex upsampling: mask [pix, mask_pitch >= 576] (already x0.25), fp32 or -- mask_is_half != 0 -- fp16
 * (the reference's autocast path hands upsample_flow an fp16 mask too; +4e-5 px, tools/fp16_mask_sensitivity.py),
 * flow fp32 [B,H8,W8,2] -> flow_up fp32 NCHW [B,2,8*H8,8*W8]; flow_lo NCHW [B,2,H8,W8] is written when non-NULL.          */
int atdn_convex_upsample(const void* mask, int32_t mask_is_half, int64_t mask_pitch, const float* flow, float* flow_up,
                         float* flow_lo_or_null, int32_t batch, int32_t h8, int32_t w8, void* stream);
/* coords_grid (utils.py:76-79): coords[b,y,x,:] = (x, y) (+ flow_init NCHW [B,2,H8,W8] when non-NULL) */
int atdn_coords_init(float* coords1, float* flow, const float* flow_init_or_null, int32_t batch, int32_t h8,
                     int32_t w8, void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32 CUDA-core layers of the small-channel networks (CLVO encoder, MappingVAE encoder)
 * atdn_vslam/layers/conv.py:36-37 (Conv = bn(mish(conv))), :83-90 (ResidualConv),
 * atdn_vslam/odometry/network.py:63-73,131-134, atdn_vslam/localization/network.py:29-45,57-72.
 *
 * y = post2( post( conv(pre(x)) + bias ) ),  NCHW fp32.
 *   pre  : x * in_scale[c] + in_shift[c] applied to in-bounds inputs (flow/RGB normalisation and the
 *          depthwise 1x1 of encoder_CNN.0), or NULL
 *   post : mish then affine bn (scale/shift per channel), each optional            (Conv block)
 *   post2: when skip != NULL: v = bn2(mish(v + skip))                               (ResidualConv tail)
 * ---------------------------------------------------------------------------------------------- */
typedef struct atdn_conv32_desc {
  const float* x; float* y;
  const float* w;         /* [cout, cin, k, k] (PyTorch layout)                                    */
  const float* bias;      /* [cout] or NULL                                                         */
  const float* in_scale;  const float* in_shift;   /* [cin] or NULL                                 */
  const float* bn_scale;  const float* bn_shift;   /* [cout] folded eval-mode batch norm, or NULL   */
  const float* skip;      /* [B, cout, oh, ow] or NULL: enables post2                               */
  const float* bn2_scale; const float* bn2_shift;  /* [cout] batch norm of post2, or NULL           */
  int32_t batch, cin, cout, in_h, in_w, k, stride, pad, mish;
  int32_t x_pitch, y_pitch; /* row pitch in elements of x and of y / skip; 0 = dense (in_w / out_w).  The 16-channel
                               CLVO layers take their input tiles by TMA when x is 16-byte aligned with x_pitch % 4 == 0 */
} atdn_conv32_desc;
int atdn_conv32(const atdn_conv32_desc* desc, void* stream);

/* y[b, o] = act(sum_i w[o, i] x[b, i] + bias[o]); act: 0 none, 1 mish (layers/linear.py:35-42)     */
int atdn_linear32(const float* x, const float* w, const float* bias, float* y, int32_t batch, int32_t in_f,
                  int32_t out_f, int32_t act, void* stream);
/* torch.nn.LSTMCell (odometry/network.py:137,139): gate order i,f,g,o; h, c updated in place.
 * gates_scratch: fp32 [batch, 4*hidden].                                                            */
int atdn_lstm_cell(const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                   float* h, float* c, float* gates_scratch, int32_t batch, int32_t in_f, int32_t hidden,
                   void* stream);

/* Persistent recurrent scan of the pose network (odometry/network.py:137-140) over `steps` time steps in ONE
 * cooperative kernel (128 CTAs, weights resident in shared memory, two grid barriers per step):
 *   (h1,c1) = LSTMCell1(feat[t], (h1,c1)); x2 = mish(W_ll h1 + b_ll); (h2,c2) = LSTMCell2(x2, (h2,c2)).
 * p1 = feat W_ih1^T + b_ih1 for all steps, fp32 [steps, batch, 2048] (one atdn_linear32 call); hidden = 512.
 * h1_0 / h2_0: initial hidden states [batch,512] (read only); c1 / c2: cell states [batch,512], updated in
 * place; h1_all / h2_all: fp32 [steps, batch, 512] hidden states of every step (h*_all[steps-1] is the new
 * state; h2_all feeds the regressor heads).  x2_scratch: fp32 [batch,512]; counter: one uint32 (the call
 * zeroes it on the stream).  batch <= 32.                                                                     */
int atdn_clvo_lstm_scan(const float* p1, const float* w_hh1, const float* b_hh1, const float* w_ll,
                        const float* b_ll, const float* w_ih2, const float* w_hh2, const float* b_ih2,
                        const float* b_hh2, const float* h1_0, float* c1, const float* h2_0, float* c2,
                        float* h1_all, float* h2_all, float* x2_scratch, uint32_t* counter, int32_t steps,
                        int32_t batch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Keyframe search -- atdn_vslam/slam_framework/neural_slam.py:373-384:
 * dist[k] = || emb[k, :] - code ||_2 ; *index = first arg-min.  emb fp32 [K, dim] (row pitch = dim).
 * ---------------------------------------------------------------------------------------------- */
int atdn_keyframe_search(const float* emb, const float* code, float* dist, int32_t* index, int64_t num,
                         int32_t dim, void* stream);

/* ------------------------------------------------------------------------------------------------
 * HOST function (all pointers are HOST pointers, no device work): pose chaining and the keyframe rule over the
 * relative poses of `num` consecutive pairs -- atdn_vslam/utils/transforms.py:54-119 (euler2matrix 'yxz',
 * matrix2euler, transform) and slam_framework/neural_slam.py:204-215 (current_pose @= T), :288-302
 * (__decide_keyframe: propagation @= T; keyframe when ||euler(propagation)|| > rot_threshold or
 * ||t(propagation)|| > tr_threshold, then propagation = I).  fp32, operation order of the reference formulas.
 * rot, tr: [num, 3] Euler angles / translations; cos_sin_or_null: optional [num, 6] = (cos(rot), sin(rot)) computed
 * by the caller's math library (then the poses are bit-identical to that library's chain), else libm is used.
 * poses: [num+1, 4, 4] row-major (poses[0] = I); is_key: [num+1] (is_key[0] = 1).
 * ---------------------------------------------------------------------------------------------- */
int atdn_pose_chain(const float* rot, const float* cos_sin_or_null, const float* tr, int64_t num,
                    float rot_threshold_rad, float tr_threshold, float* poses, int32_t* is_key);

#ifdef __cplusplus
}
#endif
#endif /* ATDN_B200_H_ */
