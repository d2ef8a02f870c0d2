"""Tensor-level wrappers around the C ABI (one Python call = one kernel launch on the current stream).

Activations are NHWC fp16 buffers ``[B, H, W, pitch]``; a :class:`View` names a channel slice of such
a buffer so that concatenations (``torch.cat`` in the reference) are free: producers write into
slices of a wider buffer and consumers read the slices through their own TMA tensor maps.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib as L


def _profiled(fn):
    """When bench.py installs L.PROFILER, time this op's launches with CUDA events under the op's name."""
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **kw):
        if L.PROFILER is None:
            return fn(*a, **kw)
        with L.PROFILER(fn.__name__, 0.0, 0.0):
            return fn(*a, **kw)
    return wrapper


class View:
    """Channels [c0, c0+c) of an NHWC buffer ``t`` of shape [B, H, W, pitch]."""

    def __init__(self, t, c0=0, c=None):
        assert t.dim() == 4 and t.is_contiguous()
        self.t, self.c0 = t, c0
        self.c = (t.shape[3] - c0) if c is None else c
        assert 0 <= c0 and c0 + self.c <= t.shape[3]

    B = property(lambda s: s.t.shape[0])
    H = property(lambda s: s.t.shape[1])
    W = property(lambda s: s.t.shape[2])
    pitch = property(lambda s: s.t.shape[3])

    def ptr(self):
        return L.ptr(self.t, self.c0)


def pad_bias(b, n=None):
    """fp32 bias padded with zeros to a multiple of 64 entries (epilogues read whole 8-groups)."""
    n = b.numel() if n is None else n
    out = torch.zeros((n + 63) // 64 * 64, dtype=torch.float32, device=b.device)
    out[: b.numel()] = b.float()
    return out


def pack_conv_weight(w):
    """[Cout, Cin, kh, kw] fp32 -> K-major fp16 [Cout, taps * ceil(Cin/64)*64], K = (tap, channel)."""
    cout, cin, kh, kw = w.shape
    cpad = (cin + 63) // 64 * 64
    wp = torch.zeros(cout, kh * kw, cpad, dtype=torch.float32, device=w.device)
    wp[:, :, :cin] = w.float().permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    return wp.reshape(cout, kh * kw * cpad).half().contiguous()


def pack_rows_weight(w2d):
    """[Cout, K] fp32 -> fp16 [Cout, ceil(K/64)*64] (zero padded)."""
    cout, k = w2d.shape
    kp = (k + 63) // 64 * 64
    wp = torch.zeros(cout, kp, dtype=torch.float32, device=w2d.device)
    wp[:, :k] = w2d.float()
    return wp.half().contiguous()


def _fill_weight(d, wp):
    cout, ktot = wp.shape
    d.b = L.ptr(wp)
    L._set(d.b_dims, (ktot, cout, 1, 1))
    L._set(d.b_strides, (ktot, ktot * cout, ktot * cout))


def _fill_out(d, out, n_valid, alpha, bias):
    d.n_valid = n_valid
    d.alpha = alpha
    d.bias = L.ptr(bias)
    if out is not None:
        d.out = out.ptr()
        d.out_pitch = out.pitch
        d.out_ch_off = 0


def conv_tc(a, wp, bias, out, *, cout, taps=(1, 1), pad=(0, 0), stride=1, bn=128, epi=L.EPI_STORE16, flags=0,
            alpha=1.0, a2=None, resid=None, h32=None, z32=None, rh16=None, aux32=None, gamma=None, mt=0, stamps=None,
            out_hw=None, out_pitch=None, aux_half_offset=None):
    """Convolution as implicit GEMM on tcgen05.  ``a`` (and optional ``a2``, concatenated after it)
    are NHWC fp16 Views of the INPUT image; ``out`` is a View of the output buffer."""
    d = L.TcDesc()
    d.bn, d.epi, d.flags, d.a_mode, d.b_mode = bn, epi, flags, L.MODE_PATCH, L.MODE_ROWS
    d.mt = mt
    if stamps is not None:
        d.lvl[2] = stamps.data_ptr()
    kh, kw = taps
    oh = (a.H + 2 * pad[0] - kh) // stride + 1
    ow = (a.W + 2 * pad[1] - kw) // stride + 1
    if out_hw is not None:   # asymmetric padding: `pad` is the leading pad, the trailing one follows from the output size
        oh, ow = out_hw
    d.out_h, d.out_w = oh, ow
    d.taps_h, d.taps_w, d.pad_h, d.pad_w, d.stride = kh, kw, pad[0], pad[1], stride
    d.a = a.ptr()
    L._set(d.a_dims, (a.c, a.W, a.H, a.B))
    L._set(d.a_strides, (a.pitch, a.W * a.pitch, a.H * a.W * a.pitch))
    if a2 is not None:
        assert a.c % 64 == 0 and (a2.B, a2.H, a2.W) == (a.B, a.H, a.W)
        d.a_split_chunk = a.c // 64
        d.a2 = a2.ptr()
        L._set(d.a2_dims, (a2.c, a2.W, a2.H, a2.B))
        L._set(d.a2_strides, (a2.pitch, a2.W * a2.pitch, a2.H * a2.W * a2.pitch))
    _fill_weight(d, wp)
    _fill_out(d, out, cout, alpha, bias)
    if resid is not None:
        d.resid16, d.resid_pitch, d.resid_ch_off = resid.ptr(), resid.pitch, 0
    d.h32, d.z32, d.rh16, d.aux32, d.gamma = L.ptr(h32), L.ptr(z32), L.ptr(rh16), L.ptr(aux32), L.ptr(gamma)
    if out_pitch is not None:       # F_TILED32: floats per 128-channel buffer
        d.out_pitch = out_pitch
    if aux_half_offset is not None:  # GRU_ZR with a pre-activation term: offset (floats) of the r half
        d.resid_pitch = aux_half_offset
    L.tc_gemm(d)
    return oh, ow


def gemm_rows(a, a_k, a_rows, a_pitch, batch, b, b_rows, b_pitch, out_ptr, out_pitch, *, n_valid, a_bstride=None,
              b_bstride=None, bn=128, epi=L.EPI_STORE16, flags=0, alpha=1.0, bias=None, resid_ptr=None,
              resid_pitch=0, aux32=None, gamma=None, b_k=None, b8=None, a_hot=None):
    """Batched D[b] = A[b] (a_rows x a_k) . B[b]^T (b_rows x a_k); pointers are ctypes void pointers,
    pitches in elements.  ``b_bstride=None`` shares B across the batch.  ``b_k``: true K extent of B when A's column
    count is padded (F_A_TILED: A is read in whole 64-column blocks, B is zero-filled past its extent).  ``b8`` / ``a_hot``:
    the e4m3 planes of B and the block bitmap of A for F_A_MIXED."""
    d = L.TcDesc()
    d.bn, d.epi, d.a_mode, d.b_mode = bn, epi, L.MODE_ROWS, L.MODE_ROWS
    d.flags = flags | (L.F_B_BATCHED if b_bstride is not None else 0)
    a_bstride = a_rows * a_pitch if a_bstride is None else a_bstride
    d.a = a
    L._set(d.a_dims, (a_k, a_rows, 1, batch))
    L._set(d.a_strides, (a_pitch, a_bstride, a_bstride))
    d.b = b
    nb = batch if b_bstride is not None else 1
    bs = b_bstride if b_bstride is not None else b_rows * b_pitch
    L._set(d.b_dims, (a_k if b_k is None else b_k, b_rows, 1, nb))
    L._set(d.b_strides, (b_pitch, bs, bs))
    d.n_valid, d.alpha, d.bias = n_valid, alpha, L.ptr(bias)
    d.out, d.out_pitch, d.out_ch_off = out_ptr, out_pitch, 0
    if resid_ptr is not None:
        d.resid16, d.resid_pitch, d.resid_ch_off = resid_ptr, resid_pitch, 0
    d.aux32, d.gamma = L.ptr(aux32), L.ptr(gamma)
    d.b8, d.a_hot = L.ptr(b8), L.ptr(a_hot)
    L.tc_gemm(d)


# ----------------------------------------------------------------------------------------------
# fp32 recurrent state (GRU hidden state master copy h32, update gate z32): tiled layout of tc_epilogue.cuh
# ----------------------------------------------------------------------------------------------
def state_alloc(b, h8, w8, device):
    """Uninitialised fp32 state buffer for a [b, h8, w8, 128] map in the epilogues' tiled layout:
    [b, ceil(h8/16), ceil(w8/8), 4 (32-pixel quarter), 32 (channel group), 32 (pixel), 4 (channel)]."""
    return torch.empty(b, (h8 + 15) // 16, (w8 + 7) // 8, 4, 32, 32, 4, dtype=torch.float32, device=device)


def state_from_nhwc(x):
    """[B, H, W, 128] fp32 -> tiled state buffer (pad pixels zero)."""
    b, h, w, c = x.shape
    assert c == 128
    th, tw = (h + 15) // 16, (w + 7) // 8
    xp = torch.zeros(b, th * 16, tw * 8, 128, dtype=torch.float32, device=x.device)
    xp[:, :h, :w] = x.float()
    # (b, th, q, rh_lo, tw, rw, c4, ci) -> (b, th, tw, q, c4, rh_lo, rw, ci)
    t = xp.view(b, th, 4, 4, tw, 8, 32, 4).permute(0, 1, 4, 2, 6, 3, 5, 7).contiguous()
    return t.view(b, th, tw, 4, 32, 32, 4)


def state_to_nhwc(t, h, w):
    """Tiled state buffer -> [B, h, w, 128] fp32."""
    b, th, tw = t.shape[:3]
    x = t.float().view(b, th, tw, 4, 32, 4, 8, 4).permute(0, 1, 3, 5, 2, 6, 4, 7).reshape(b, th * 16, tw * 8, 128)
    return x[:, :h, :w].contiguous()


# ----------------------------------------------------------------------------------------------
# correlation pyramid
# ----------------------------------------------------------------------------------------------
def pyramid_shapes(h8, w8):
    """[(H_l, W_l, pitch_l)] for the 4 levels.  The row pitch rounds W_l up so that every box the pyramid kernel
    stores (32 / 16 / 8 / 4 floats per row at levels 0..3) starts on a 128 / 64 / 32 / 32-byte boundary: level-0
    rows are whole cache lines and no store leaves a partially written 32-byte sector behind."""
    out, h, w = [], h8, w8
    for l in range(4):
        q = (32, 16, 8, 8)[l]
        out.append((h, w, (w + q - 1) // q * q))
        h, w = h // 2, w // 2
    return out


def alloc_pyramid(batch, h8, w8, device, half_levels=0):
    """The four pyramid levels.  ``half_levels=0``: the reference's fp32 pyramid, row-major [B*N, H_l, pitch_l]
    (``CorrBlock`` drop-in).  ``half_levels=4`` (the sequence pipeline): fp16 in the STRIP layout of
    ``atdn_corr_pyramid`` (include/atdn_b200.h): level l = [B*N, tiles, chunk_l] with tiles = ceil(h8/8) * ceil(w8/32)
    8 x 32-texel target tiles and, inside a tile's chunk,
      level 0: [strip 4][row 8][col 8]   level 1: [strip pair 2][row 4][col 8]   level 2: [row 2][col 8]
    level 3: [B*N, ceil(h8/8) * tiles_w3, 4] (one row of 4 texels per tile, tile columns padded to an even count; the pad
    tiles are never written and must stay zero, hence ``zeros``).  A 64-byte DRAM fetch granule of the lookup is a
    4 x 8-texel block at levels 0 / 1, and every row piece of 8 texels is 16-byte aligned at every level."""
    n = h8 * w8
    if half_levels:
        assert half_levels == 4
        th, tw = (h8 + 7) // 8, (w8 + 31) // 32
        lv = [torch.empty(batch * n, th * tw, 256 >> (2 * l), dtype=torch.float16, device=device) for l in range(3)]
        lv.append(torch.zeros(batch * n, th * ((tw + 1) // 2 * 2), 4, dtype=torch.float16, device=device))
        return lv
    return [torch.empty(batch * n, h, p, dtype=torch.float32, device=device) for (h, w, p) in pyramid_shapes(h8, w8)]


def pyramid_untile(levels, h8, w8, padded=False):
    """Strip-layout fp16 pyramid -> list of row-major fp32 [B*N, H_l, W_l] tensors (tests / inspection);
    ``padded``: keep the whole tile grid (the texels outside the maps, which the kernel writes as zeros)."""
    th, tw = (h8 + 7) // 8, (w8 + 31) // 32
    nq = levels[0].shape[0]
    out = []
    # level 0: [q, th, tw, strip 4, row 8, col 8] -> [q, th*8, tw*32]
    x = levels[0].float().view(nq, th, tw, 4, 8, 8).permute(0, 1, 4, 2, 3, 5).reshape(nq, th * 8, tw * 32)
    out.append(x if padded else x[:, :h8, :w8].contiguous())
    # level 1: [q, th, tw, pair 2, row 4, col 8] -> [q, th*4, tw*16]
    x = levels[1].float().view(nq, th, tw, 2, 4, 8).permute(0, 1, 4, 2, 3, 5).reshape(nq, th * 4, tw * 16)
    out.append(x if padded else x[:, : h8 >> 1, : w8 >> 1].contiguous())
    # level 2: [q, th, tw, row 2, col 8] -> [q, th*2, tw*8]
    x = levels[2].float().view(nq, th, tw, 2, 8).permute(0, 1, 3, 2, 4).reshape(nq, th * 2, tw * 8)
    out.append(x if padded else x[:, : h8 >> 2, : w8 >> 2].contiguous())
    # level 3: [q, th, tw3, 4] -> [q, th, tw3*4]
    x = levels[3].float().view(nq, th, -1)
    out.append(x if padded else x[:, : h8 >> 3, : w8 >> 3].contiguous())
    return out


def _half_levels(levels):
    return sum(1 for t in levels if t.dtype == torch.float16)


def corr_pyramid_build(fmap1, fmap2, levels, legacy=False, pair=False, alpha=None):
    """fmap1/fmap2: Views [B,H8,W8,256] fp16 -> the 4 pyramid levels (corr.py:16-30, 55-63).  ``alpha``: scale of the
    volume, default 1/sqrt(C) (corr.py:62); the flow net passes 1.0 after folding 2^-2 into each feature map."""
    b, h8, w8 = fmap1.B, fmap1.H, fmap1.W
    n = h8 * w8
    if not legacy:
        assert fmap1.pitch == fmap2.pitch and fmap1.c == fmap2.c
        lv = (C.c_void_p * 4)(*[t.data_ptr() for t in levels])
        lp = (C.c_int32 * 4)(*[t.shape[2] for t in levels])

        def go():
            L.check(L.load().atdn_corr_pyramid(fmap1.ptr(), fmap2.ptr(), C.c_int64(fmap1.pitch), fmap1.c, lv, lp, _half_levels(levels), b, h8, w8,
                                               C.c_float(1.0 / math.sqrt(fmap1.c) if alpha is None else alpha), L.stream_ptr()), "atdn_corr_pyramid")
        if L.PROFILER is not None:
            # algorithmic: 2*N*N*C flop; bytes = the four levels written + both feature maps read once
            nbytes = sum(float(levels[l].element_size()) * b * n * (h8 >> l) * (w8 >> l) for l in range(4)) + 2.0 * b * n * fmap1.c * 2
            with L.PROFILER("corr_pyramid", 2.0 * b * n * n * fmap1.c, nbytes):
                go()
            return
        go()
        return
    d = L.TcDesc()
    d.bn, d.epi, d.a_mode, d.b_mode = 256, L.EPI_CORR, L.MODE_ROWS, L.MODE_PATCH
    d.flags = L.F_B_BATCHED | (L.F_PAIR if pair else 0)
    d.a = fmap1.ptr()
    L._set(d.a_dims, (fmap1.c, n, 1, b))
    L._set(d.a_strides, (fmap1.pitch, n * fmap1.pitch, n * fmap1.pitch))
    d.b = fmap2.ptr()
    L._set(d.b_dims, (fmap2.c, w8, h8, b))
    L._set(d.b_strides, (fmap2.pitch, w8 * fmap2.pitch, n * fmap2.pitch))
    d.n_valid = n
    d.alpha = 1.0 / math.sqrt(fmap1.c)
    d.out, d.out_pitch, d.out_ch_off = L.ptr(levels[0]), 8, 0
    for i in range(3):
        d.lvl[i] = levels[i + 1].data_ptr()
    for i in range(4):
        d.lvl_pitch[i] = levels[i].shape[2]
    d.corr_h, d.corr_w = h8, w8
    L.tc_gemm(d)


def corr_lookup(levels, coords, out16=None, out32=None):
    """coords fp32 [B,H8,W8,2] -> out16 View [B,H8,W8,>=324] fp16 and/or out32 [B*H8*W8,324] fp32."""
    b, h8, w8, _ = coords.shape
    lv = (C.c_void_p * 4)(*[t.data_ptr() for t in levels])
    lp = (C.c_int32 * 4)(*[t.shape[2] for t in levels])
    hl = _half_levels(levels)

    def go():
        L.check(L.load().atdn_corr_lookup(lv, lp, hl, L.ptr(coords), out16.ptr() if out16 is not None else None,
                                          C.c_int64(out16.pitch if out16 is not None else 0), L.ptr(out32),
                                          b, h8, w8, L.stream_ptr()), "atdn_corr_lookup")
    if L.PROFILER is not None:
        # algorithmic bytes: <= 4 levels x 10x10 texels read + coords + 324 outputs written per query
        q = b * h8 * w8
        nbytes = q * (100 * sum(t.element_size() for t in levels) + 8 + 324 * (2 if out16 is not None else 4))
        with L.PROFILER("corr_lookup", 0.0, float(nbytes)):
            go()
        return
    go()


# ----------------------------------------------------------------------------------------------
# element-wise
# ----------------------------------------------------------------------------------------------
@_profiled
def stem_pack(image, x16):
    """image fp32 [B,3,H,W] -> x16 [B,H/2,W/2,48]: normalised, horizontal taps of the 7x7/2 stem folded into channels."""
    b, _, h, w = image.shape
    L.check(L.load().atdn_stem_pack(L.ptr(image), L.ptr(x16), b, h, w, L.stream_ptr()), "atdn_stem_pack")


def stem_weight(w):
    """[Cout,3,7,7] -> [Cout,48,4,1] for the 4x1 convolution over the packed stem input (see atdn_stem_pack)."""
    cout = w.shape[0]
    w4 = torch.zeros(cout, 48, 4, 1, dtype=torch.float32, device=w.device)
    for ai in range(4):
        for ry in range(2):
            dy = 2 * ai + ry - 1
            if not 0 <= dy <= 6:
                continue
            for c in range(3):
                ch = (ry * 3 + c) * 8
                w4[:, ch + 1:ch + 8, ai, 0] = w[:, c, dy, :].float()
    return w4


@_profiled
def flow_pack(flow, x16):
    """flow fp32 [B,H8,W8,2] -> x16 [B,H8,W8,16]: horizontal taps of convf1 (7x7) folded into channels."""
    b, h8, w8, _ = flow.shape
    L.check(L.load().atdn_flow_pack(L.ptr(flow), L.ptr(x16), b, h8, w8, L.stream_ptr()), "atdn_flow_pack")


def flow_weight(w):
    """[Cout,2,7,7] -> [Cout,14,7,1] for the 7x1 convolution over the packed flow (see atdn_flow_pack)."""
    return w.float().permute(0, 3, 1, 2).reshape(w.shape[0], 14, 7).unsqueeze(-1).contiguous()   # [o, dx*2+c, dy, 1]


@_profiled
def inorm_stats(x, scratch, parts, stats):
    L.check(L.load().atdn_inorm_stats(x.ptr(), C.c_int64(x.pitch), x.B, x.H * x.W, x.c, L.ptr(scratch), parts,
                                      L.ptr(stats), L.stream_ptr()), "atdn_inorm_stats", 2)


def stats_parts(h, w, pair):
    """Partial-sum slots per image written by a conv launched with F_STATS (halo kernel, mt 4, bn 64)."""
    cl = 2 if pair else 1
    return ((h + 15) // 16) * ((w + 32 * cl - 1) // (32 * cl)) * cl * 4


@_profiled
def inorm_finalize(scratch, parts, stats, batch, c, hw):
    L.check(L.load().atdn_inorm_finalize(L.ptr(scratch), parts, batch, c, hw, L.ptr(stats), L.stream_ptr()), "atdn_inorm_finalize")


@_profiled
def inorm_apply(x, stats, y, resid=None, relu=True):
    L.check(L.load().atdn_inorm_apply(x.ptr(), C.c_int64(x.pitch), L.ptr(stats),
                                      resid.ptr() if resid is not None else None,
                                      C.c_int64(resid.pitch if resid is not None else 0), y.ptr(), C.c_int64(y.pitch),
                                      x.B, x.H * x.W, x.c, int(relu), L.stream_ptr()), "atdn_inorm_apply")


def attn_probs(qk, p16, inv_sum, scale, tiled=False, block_hot=None, hot_energy=1e-2):
    """qk fp16 [B,H8,W8,256] (q | k) -> p16 [B,N,Np] un-normalised probabilities, inv_sum [B*N] (gma.py:66-73).
    ``tiled``: p16 is [B, ceil(N/32), Np/64, 32, 64] (blocks of 32 rows x 64 columns, see atdn_attn_probs).
    ``block_hot`` uint8 [B, ceil(N/128), Np/64]: mixed fp16 / e4m3 storage, the bitmap is written (1 = fp16 block)."""
    b, h8, w8, pitch = qk.shape
    n = h8 * w8

    def go():
        L.check(L.load().atdn_attn_probs(L.ptr(qk), C.c_int64(pitch), L.ptr(p16), C.c_int64(p16.shape[-1] if not tiled else p16.shape[2] * 64),
                                         1 if tiled else 0, L.ptr(inv_sum), b, n, C.c_float(scale), L.ptr(block_hot), C.c_float(hot_energy),
                                         L.stream_ptr()), "atdn_attn_probs")
    if L.PROFILER is not None:
        # algorithmic: one q.k^T (2*N*N*128 flop) and N*N fp16 probabilities written per image
        with L.PROFILER("attn_probs", 2.0 * b * n * n * 128, 2.0 * b * n * n):
            go()
        return
    go()


@_profiled
def resize_aa(frames, size):
    """frames [T,C,H,W] float32 or uint8 (CUDA, contiguous) -> [T,C,size[0],size[1]] float32, antialiased bilinear
    (TF.resize of neural_slam.py:197-199)."""
    L.require_cuda(frames)
    if frames.dtype not in (torch.float32, torch.uint8):
        frames = frames.float()
    frames = frames.contiguous()
    t, c, h, w = frames.shape
    out = torch.empty(t, c, size[0], size[1], dtype=torch.float32, device=frames.device)
    L.check(L.load().atdn_resize_aa(L.ptr(frames), 1 if frames.dtype == torch.uint8 else 0, L.ptr(out), t * c, h, w, size[0], size[1],
                                    L.stream_ptr()), "atdn_resize_aa")
    return out


@_profiled
def attn_harmonize(p16, block_hot, pair_hot, n):
    """Mixed-precision P for the CTA-pair P.V kernel: 256-row bitmap, e4m3 halves next to fp16 halves rewritten as fp16."""
    L.check(L.load().atdn_attn_harmonize(L.ptr(p16), C.c_int64(p16.shape[2] * 64), L.ptr(block_hot), L.ptr(pair_hot), p16.shape[0], n,
                                         L.stream_ptr()), "atdn_attn_harmonize")


@_profiled
def flow_head_gather(d32, bias, coords1, flow):
    """d32 fp32 [B,H8,W8,pitch] per-tap partial products of flow_head.conv2 -> coords1 += delta, flow = coords1 - grid."""
    b, h8, w8, pitch = d32.shape
    L.check(L.load().atdn_flow_head_gather(L.ptr(d32), C.c_int64(pitch), L.ptr(bias), L.ptr(coords1), L.ptr(flow), b, h8, w8,
                                           L.stream_ptr()), "atdn_flow_head_gather")


@_profiled
def convex_upsample(mask, flow, flow_up, flow_lo=None):
    """mask [pix, >=576] fp32 or fp16"""
    b, h8, w8, _ = flow.shape
    L.check(L.load().atdn_convex_upsample(L.ptr(mask), 1 if mask.dtype == torch.float16 else 0, C.c_int64(mask.shape[-1]), L.ptr(flow), L.ptr(flow_up),
                                          L.ptr(flow_lo), b, h8, w8, L.stream_ptr()), "atdn_convex_upsample")


@_profiled
def coords_init(coords1, flow, flow_init=None):
    b, h8, w8, _ = coords1.shape
    L.check(L.load().atdn_coords_init(L.ptr(coords1), L.ptr(flow), L.ptr(flow_init), b, h8, w8, L.stream_ptr()),
            "atdn_coords_init")


# ----------------------------------------------------------------------------------------------
# fp32 small nets
# ----------------------------------------------------------------------------------------------
def nchw_pitch(t):
    """Row pitch of an NCHW map that is dense except for padded rows (a [..., :w] view of a wider buffer)."""
    bb, cc, hh, ww = t.shape
    # strides of size-1 dimensions carry no information: take the pitch from the first dimension that has one
    pt = t.stride(2) if hh > 1 else t.stride(1) if cc > 1 else t.stride(0) if bb > 1 else ww
    assert pt >= ww and (ww == 1 or t.stride(3) == 1) and (cc == 1 or t.stride(1) == hh * pt) and \
        (bb == 1 or t.stride(0) == cc * hh * pt), "unsupported strides"
    return pt


@_profiled
def conv32(x, w, bias, y, *, stride=1, pad=0, mish=False, in_scale=None, in_shift=None, skip=None, bn_scale=None,
           bn_shift=None, bn2_scale=None, bn2_shift=None):
    d = L.Conv32Desc()
    b, cin, h, wd = x.shape
    cout, _, k, _ = w.shape

    pitch = nchw_pitch
    d.x, d.y, d.w, d.bias = L.ptr(x), L.ptr(y), L.ptr(w), L.ptr(bias)
    d.in_scale, d.in_shift, d.skip = L.ptr(in_scale), L.ptr(in_shift), L.ptr(skip)
    d.bn_scale, d.bn_shift = L.ptr(bn_scale), L.ptr(bn_shift)
    d.bn2_scale, d.bn2_shift = L.ptr(bn2_scale), L.ptr(bn2_shift)
    d.batch, d.cin, d.cout, d.in_h, d.in_w, d.k, d.stride, d.pad, d.mish = b, cin, cout, h, wd, k, stride, pad, int(mish)
    d.x_pitch, d.y_pitch = pitch(x), pitch(y)
    if skip is not None:
        assert skip.shape == y.shape and pitch(skip) == d.y_pitch, "skip must share the layout of y"
    L.check(L.load().atdn_conv32(C.byref(d), L.stream_ptr()), "atdn_conv32")


@_profiled
def linear32(x, w, bias, y, act=0):
    L.check(L.load().atdn_linear32(L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(y), x.shape[0], w.shape[1], w.shape[0], act,
                                   L.stream_ptr()), "atdn_linear32")


@_profiled
def lstm_cell(x, w_ih, w_hh, b_ih, b_hh, h, c, gates):
    L.check(L.load().atdn_lstm_cell(L.ptr(x), L.ptr(w_ih), L.ptr(w_hh), L.ptr(b_ih), L.ptr(b_hh), L.ptr(h), L.ptr(c),
                                    L.ptr(gates), x.shape[0], w_ih.shape[1], h.shape[1], L.stream_ptr()), "atdn_lstm_cell", 2)


@_profiled
def clvo_lstm_scan(p1, lstm1, ll, lstm2, h1_0, c1, h2_0, c2, h1_all, h2_all, x2, counter):
    """Persistent LSTM scan (odometry/network.py:137-140); lstm* = (w_ih, w_hh, b_ih, b_hh), ll = (w, b)."""
    steps, batch = h1_all.shape[0], h1_all.shape[1]
    L.check(L.load().atdn_clvo_lstm_scan(L.ptr(p1), L.ptr(lstm1[1]), L.ptr(lstm1[3]), L.ptr(ll[0]), L.ptr(ll[1]),
                                         L.ptr(lstm2[0]), L.ptr(lstm2[1]), L.ptr(lstm2[2]), L.ptr(lstm2[3]),
                                         L.ptr(h1_0), L.ptr(c1), L.ptr(h2_0), L.ptr(c2), L.ptr(h1_all), L.ptr(h2_all),
                                         L.ptr(x2), L.ptr(counter), steps, batch, L.stream_ptr()), "atdn_clvo_lstm_scan")


@_profiled
def keyframe_search(emb, code, dist, index):
    L.check(L.load().atdn_keyframe_search(L.ptr(emb), L.ptr(code), L.ptr(dist), L.ptr(index), C.c_int64(emb.shape[0]),
                                          emb.shape[1], L.stream_ptr()), "atdn_keyframe_search", 2)
