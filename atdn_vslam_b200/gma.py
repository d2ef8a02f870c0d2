"""Drop-in for the reference GMA flow network (``GMA.whl!/GMA/core/network.py:26-129``).

``RAFTGMA(args).forward(image1, image2, iters=12, flow_init=None, upsample=True, test_mode=False)``
keeps the reference signature, parameter names (so the reference state dict loads unchanged, with or
without the ``module.`` prefix of ``DataParallel``) and return values.  All arithmetic runs in
hand-written sm_100a kernels through the C ABI; PyTorch only owns the device buffers and the stream.

Data layout in HBM (per batch of B pairs, 1/8-resolution grid H8 x W8, N = H8*W8):
  activations      NHWC fp16, channel pitch padded to 8
  HX  [B,H8,W8,512] fp16 : [0:128] GRU hidden state h | [128:256] context inp | [256:384] motion
                          features | [384:512] globally aggregated motion features  (the reference's
                          torch.cat([net, inp, mf, mfg]) materialised once, never copied)
  h32, z32         hidden state and update gate in the tiled layout of csrc/tc_epilogue.cuh (one warp access =
                   contiguous bytes; ops.state_alloc); fp16 by default (ATDN_F_H16 / ATDN_F_Z16), fp32 optional
  corr pyramid     fp16 [B*N, H_l, pitch_l] for l = 0..3 (pitch = W_l rounded up to 32/16/8/8 elements: sector-aligned
                   store boxes), pooled in fp32 and rounded once on store; CorrBlock keeps the reference's fp32 pyramid
  P   [B,N,Np]     fp16 un-normalised attention probabilities exp(s - max), Np = N rounded up to 64
                   (written by the fused q.k^T/softmax kernel: the fp32 logits never reach HBM);
  inv_sum [B*N]    fp32 1/sum(P) applied in the P.V epilogue
  coords1, flow    fp32 [B,H8,W8,2]
"""
from __future__ import annotations

import math
import os

import torch
from torch import nn

from . import _lib as L
from . import ops, schema
from .ops import View


def _build_module_tree(root: nn.Module, sch):
    """Create nested containers so that ``root.state_dict()`` has exactly the schema's names."""
    for name, (shape, kind) in sch.items():
        parts = name.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        if kind in ("bn_mean", "bn_var"):
            mod.register_buffer(parts[-1], torch.zeros(shape) if kind == "bn_mean" else torch.ones(shape))
        elif kind == "bn_count":
            mod.register_buffer(parts[-1], torch.tensor(0, dtype=torch.long))
        elif kind == "index":
            n = shape[0]
            mod.register_buffer(parts[-1], torch.arange(n).view(1, -1) - torch.arange(n).view(-1, 1) + n - 1)
        else:
            if kind == "conv_w":
                fan = shape[1] * shape[2] * shape[3]
                t = torch.randn(shape) / math.sqrt(fan)
            elif kind == "lin_w":
                t = torch.randn(shape) / math.sqrt(shape[1])
            elif kind == "bn_w":
                t = torch.ones(shape)
            elif kind == "emb":
                t = torch.randn(shape)
            else:
                t = torch.zeros(shape)
            mod.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))


def _fold_bn(w, b, sd, name):
    """conv -> eval-mode batch norm folded into the conv (extractor.py:21-26, norm after conv)."""
    s = sd[name + ".weight"].float() / torch.sqrt(sd[name + ".running_var"].float() + 1e-5)
    return w.float() * s.view(-1, 1, 1, 1), (b.float() - sd[name + ".running_mean"].float()) * s + sd[name + ".bias"].float()


# (mt, bn, CTA pair) of the persistent halo-reuse kernel per output width (stride-1 convs); measured on B200 at
# 47x154 x 24 pairs / encoder resolutions (profiles/r01b_halo_conv_timings.log)
_HALO_CFG = {64: (4, 64, True), 96: (2, 96, True), 128: (2, 128, True), 192: (1, 192, True), 256: (1, 256, True),
             576: (1, 192, False)}
for _item in filter(None, os.environ.get("ATDN_HALO_CFG", "").split(";")):      # A/B: "128:1,128,1;192:1,192,0" = cout:mt,bn,pair
    _cout, _cfg = _item.split(":")
    _mt, _bn, _pair = (int(v) for v in _cfg.split(","))
    _HALO_CFG[int(_cout)] = (_mt, _bn, bool(_pair))


FMAP_SCALE = 0.25      # plan.buffer("fmap*") holds fnet(x) * FMAP_SCALE (see _EncoderWeights.out)
_NO_FUSED_STATS = os.environ.get("ATDN_NO_FUSED_STATS") == "1"      # A/B switches (bench only)
_NO_GRU_PRE = os.environ.get("ATDN_NO_GRU_PRE") == "1"
_MASK32 = os.environ.get("ATDN_MASK32") == "1"
_PV_PAIR = os.environ.get("ATDN_PV_PAIR") == "1"      # A/B: P.V on CTA pairs (half the V^T stream per SM)
# attention probabilities in blocks of 32 rows x 64 columns: the store boxes of attn_probs and the operand boxes of P.V are
# contiguous 4 KiB runs instead of 128-byte rows 14.6 KB apart (ATDN_P_ROWMAJOR=1: plain [N, Np] rows)
_P_TILED = os.environ.get("ATDN_P_ROWMAJOR") != "1" and not _PV_PAIR
# Mixed-precision attention probabilities (include/atdn_b200.h, atdn_attn_probs block_hot): 128 x 64 blocks of P whose
# e4m3 rounding cannot move a row's aggregate stay out of the fp16 stream that P.V re-reads 12 times per pair.
# ATDN_P_MIXED=0: every block fp16.  ATDN_P_HOT_ENERGY: the criterion's threshold (tools/fp8_attention_sensitivity.py).
_P_MIXED = os.environ.get("ATDN_P_MIXED", "1") == "1" and _P_TILED
_P_HOT_ENERGY = float(os.environ.get("ATDN_P_HOT_ENERGY", "1e-2"))
_GRU_PRE32 = os.environ.get("ATDN_GRU_PRE32") == "1"
_PRE16 = 0 if _GRU_PRE32 else L.F_PRE16
_Z16 = 0 if os.environ.get("ATDN_GRU_Z32") == "1" else L.F_Z16
# fp16-only hidden state (no fp32 master copy), like the reference's own fp16-autocast path: measured flow EPE
# 7.105e-3 vs 7.107e-3 px with the fp32 master (ATDN_GRU_H32=1 restores it), 1.5 KB/pixel/iteration less GRU traffic
_H16 = 0 if os.environ.get("ATDN_GRU_H32") == "1" else L.F_H16


# Small problems -- the reference's per-frame call (batch 1: 60 tiles of 16 x 8 pixels at 1/8 resolution for 148 SMs): the
# throughput configurations above would occupy 15..60 CTAs, so the work is cut into single-CTA tiles of 128 pixels x 64
# (128) outputs instead: 120 CTAs per layer (profiles/r02d_ncu_launches_batch1_forward.csv: 18..29 us per GRU conv launch).
_HALO_SMALL = {64: (1, 64), 128: (1, 64), 192: (1, 64), 256: (1, 128)}
_SMALL_TILES = 128       # "small" = at most this many 16 x 8-pixel tiles in the whole batch


def _is_small(x):
    return x.B * math.ceil(x.H / 16) * math.ceil(x.W / 8) <= _SMALL_TILES


def _halo(cout, taps=(3, 3), small=False):
    if small and cout in _HALO_SMALL:
        mt, bn = _HALO_SMALL[cout]
        return {"mt": mt, "bn": bn, "flags": 0}
    mt, bn, pair = _HALO_CFG[cout]
    if taps == (1, 1) and cout == 256:
        pair = False
    return {"mt": mt, "bn": bn, "flags": L.F_PAIR if pair else 0}


def _conv_s1(x, c, out, *, cout, taps, flags=0, **kw):
    """Stride-1 convolution on the halo kernel; `c` is a _Conv."""
    cfg = _halo(cout, taps, _is_small(x) and not (flags & L.F_STATS))
    return ops.conv_tc(x, c.wp, c.bias, out, cout=cout, taps=taps, pad=(taps[0] // 2, taps[1] // 2), bn=cfg["bn"],
                       mt=cfg["mt"], flags=flags | cfg["flags"], **kw)


def _pick_bn(cout, m_tiles):
    """Largest N tile that still yields >= ~1 wave of the 148 SMs."""
    for bn in (192, 128, 96, 64):
        if cout % bn == 0 and m_tiles * (cout // bn) >= 140:
            return bn
    for bn in (64, 96, 128, 192):
        if cout % bn == 0:
            return bn
    return 64 if cout <= 64 else 128


class _Conv:
    """Packed weights of one convolution (K-major fp16 for the tcgen05 kernel, padded fp32 bias)."""

    def __init__(self, w, b, taps=None, rows=False):
        self.cout = w.shape[0]
        self.cin = w.shape[1]
        self.taps = (w.shape[2], w.shape[3])
        self.pad = (w.shape[2] // 2, w.shape[3] // 2)
        if rows:   # im2col layers: K = (tap, channel) flattened
            self.wp = ops.pack_rows_weight(w.float().permute(0, 2, 3, 1).reshape(self.cout, -1))
        else:
            self.wp = ops.pack_conv_weight(w)
        self.bias = ops.pad_bias(b, self.cout) if b is not None else None


class _EncoderWeights:
    def __init__(self, sd, p, norm, out_scale=1.0):
        self.norm = norm

        def conv(name, bn_name=None):
            w, b = sd[name + ".weight"], sd[name + ".bias"]
            if norm == "batch" and bn_name is not None:
                w, b = _fold_bn(w, b, sd, bn_name)
            return w, b

        w, b = conv(p + "conv1", p + "norm1")
        self.stem = _Conv(ops.stem_weight(w), b)      # 7x7/2 as a 4x1 conv over the row-packed image (ops.stem_pack)
        self.blocks = []
        for li, (planes, stride) in enumerate(schema.ENCODER_STAGES, start=1):
            for bi in range(2):
                q = f"{p}layer{li}.{bi}."
                st = stride if bi == 0 else 1
                blk = {"stride": st, "planes": planes,
                       "conv1": _Conv(*conv(q + "conv1", q + "norm1")),
                       "conv2": _Conv(*conv(q + "conv2", q + "norm2"))}
                if st != 1:
                    blk["down"] = _Conv(*conv(q + "downsample.0", q + "downsample.1"))
                self.blocks.append(blk)
        # out_scale (a power of two, exact in fp16): the feature net stores fmap / 4, so that <fmap1, fmap2> already carries
        # the 1/sqrt(256) = 2^-4 of corr.py:62 and the pyramid epilogue skips 256 multiplies per query and tile
        self.out = _Conv(sd[p + "conv2.weight"].float() * out_scale, sd[p + "conv2.bias"].float() * out_scale)


class _Packed:
    """All GMA weights in kernel layouts (built once per device from the module's state dict)."""

    def __init__(self, sd):
        self.fnet = _EncoderWeights(sd, "fnet.", "instance", out_scale=FMAP_SCALE)
        self.cnet = _EncoderWeights(sd, "cnet.", "batch")
        u = "update_block."
        g = lambda n: (sd[u + n + ".weight"], sd[u + n + ".bias"])
        self.convc1 = _Conv(*g("encoder.convc1"))
        self.convc2 = _Conv(*g("encoder.convc2"))
        wf, bf = g("encoder.convf1")
        self.convf1 = _Conv(ops.flow_weight(wf), bf)           # 7x7 as a 7x1 conv over the row-packed flow (ops.flow_pack)
        self.convf2 = _Conv(*g("encoder.convf2"))
        w, b = g("encoder.conv")                       # 126 outputs + 2 raw flow channels (update.py:84)
        self.conv = _Conv(w, torch.cat([b.float(), torch.zeros(2, device=b.device)]))
        # SepConvGRU (update.py:48-63) on hx = [h | inp | mf | mfg]: the context features `inp` do not change over the
        # refinement iterations, so their part of every gate convolution (+ bias) is computed once per pair
        # (gru_pre: zr / q weights over input channels 128..255) and the per-iteration convolutions contract only
        # [h | mf | mfg] (384 of 512 channels: 25% fewer MMAs); the epilogues add the stored term.
        self.gru, self.gru_pre = [], []
        rest = lambda w: torch.cat([w[:, :128], w[:, 256:]], 1)
        for n in ("1", "2"):
            wz, bz = g("gru.convz" + n)
            wr, br = g("gru.convr" + n)
            wq, bq = g("gru.convq" + n)
            wzr, bzr = torch.cat([wz, wr], 0), torch.cat([bz, br], 0)
            if _NO_GRU_PRE:
                self.gru.append((_Conv(wzr, bzr), _Conv(wq, bq)))
            else:
                self.gru.append((_Conv(rest(wzr), None), _Conv(rest(wq), None)))
                self.gru_pre.append((_Conv(wzr[:, 128:256], bzr), _Conv(wq[:, 128:256], bq)))
        self.fh1 = _Conv(*g("flow_head.conv1"))
        # flow_head.conv2 (3x3, 256 -> 2) as a 1x1 conv 256 -> 18 (row = tap*2 + co, zero-padded to one 32-column MMA
        # tile) whose per-tap partial products are summed with their shifts by ops.flow_head_gather
        w2, b2 = g("flow_head.conv2")
        w18 = torch.zeros(32, w2.shape[1], 1, 1, dtype=torch.float32, device=w2.device)
        w18[:18, :, 0, 0] = w2.float().permute(2, 3, 0, 1).reshape(18, w2.shape[1])      # [dy, dx, co, c]
        self.fh2 = _Conv(w18, None)
        self.fh2_bias = b2.float().contiguous()
        self.mask0 = _Conv(*g("mask.0"))
        self.mask2 = _Conv(*g("mask.2"))
        self.gamma = sd[u + "aggregator.gamma"].float().contiguous()
        # to_v runs with the weight as the A operand (output stored transposed for the P.V GEMM)
        self.to_v = ops.pack_rows_weight(sd[u + "aggregator.to_v.weight"].float().reshape(128, 128))
        self.to_qk = _Conv(sd["att.to_qk.weight"], None)


class _Plan:
    """Device buffers for a batch of ``b`` pairs at image size h x w (allocated once, reused)."""

    def __init__(self, b, h, w, dev):
        assert h % 8 == 0 and w % 8 == 0, "image sides must be multiples of 8 (use InputPadder as the reference does)"
        self.b, self.h, self.w = b, h, w
        self.h8, self.w8 = h // 8, w // 8
        self.n = n = self.h8 * self.w8
        assert self.h8 >= 16 and self.w8 >= 16, "1/8-resolution grid must be at least 16x16 (4-level pyramid)"
        self.np_ = (n + 63) // 64 * 64
        f16 = lambda *s: torch.empty(*s, dtype=torch.float16, device=dev)
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        self.enc = {}          # encoder scratch keyed by number of images
        h8, w8 = self.h8, self.w8
        self.hx = f16(b, h8, w8, 512)
        # recurrent state in the tiled layout of csrc/tc_epilogue.cuh: hidden state h, update gate z (fp16 by default)
        st = ops.state_alloc(b, h8, w8, dev)
        self.state_numel = st.numel()
        self.h32 = st.half() if _H16 else st
        self.z = torch.empty_like(st, dtype=torch.float16) if _Z16 else torch.empty_like(st)
        self.z32 = self.z                       # (name kept for the fp32 variant and the tests)
        # context part of the GRU gate convolutions per GRU half: [z | r] and q, tiled like h32, fp16 (one rounding of a
        # pre-activation: the same size as the fp16 rounding of the conv operands) unless ATDN_GRU_PRE32=1
        pdt = torch.float32 if _GRU_PRE32 else torch.float16
        self.pre_zr = [torch.empty((2,) + tuple(st.shape), dtype=pdt, device=dev) for _ in range(2)]
        self.pre_q = [torch.empty(tuple(st.shape), dtype=pdt, device=dev) for _ in range(2)]
        self.rh = f16(b, h8, w8, 128)
        self.pyr = ops.alloc_pyramid(b, h8, w8, dev, half_levels=4)
        self.qk = f16(b, h8, w8, 256)
        self.p16 = f16(b, (n + 31) // 32, self.np_ // 64, 32, 64) if _P_TILED else f16(b, n, self.np_)
        self.inv_sum = f32(b * n)
        self.vt = f16(b, 128, self.np_)
        self.mixed = _P_MIXED        # at every batch size: the rounding of a pair must not depend on how pairs are batched
        if self.mixed:
            self.v8 = torch.empty(b, 2, 128, self.np_, dtype=torch.uint8, device=dev)          # e4m3 planes hi, lo of v^T
            self.p_hot = torch.empty(b, (n + 31) // 32, self.np_ // 64, dtype=torch.uint8, device=dev)      # per 32-row sub-block
            self.p_hot2 = torch.empty(b, (n + 255) // 256, self.np_ // 64, dtype=torch.uint8, device=dev)   # per 256-row P.V tile
        self.coords1 = f32(b, h8, w8, 2)
        self.flow = f32(b, h8, w8, 2)
        self.corrfeat = f16(b, h8, w8, 328)
        self.c1 = f16(b, h8, w8, 256)
        self.corflo = f16(b, h8, w8, 256)
        self.fpack = f16(b, h8, w8, 16)
        self.f1 = f16(b, h8, w8, 128)
        self.fh = f16(b, h8, w8, 256)
        self.fh2d = f32(b, h8, w8, 32)      # flow_head.conv2 per-tap partial products [.., tap*2 + co]
        self.mh = f16(b, h8, w8, 256)
        # up-sampling mask: fp16 like the reference's autocast path (+4e-5 px, tools/fp16_mask_sensitivity.py); ATDN_MASK32=1: fp32
        self.mask32 = f32(b * n, 576) if _MASK32 else f16(b * n, 576)

    def buffer(self, name, shape, dtype):
        t = self.__dict__.get("_buf_" + name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(*shape, dtype=dtype, device=self.hx.device)
            self.__dict__["_buf_" + name] = t
        return t

    def encoder_scratch(self, nimg, dev):
        if nimg not in self.enc:
            f16 = lambda *s: torch.empty(*s, dtype=torch.float16, device=dev)
            h2, w2, h4, w4, h8, w8 = self.h // 2, self.w // 2, self.h // 4, self.w // 4, self.h8, self.w8
            self.enc[nimg] = {
                "xpack": f16(nimg, h2, w2, 48),
                "r2": [f16(nimg, h2, w2, 64) for _ in range(4)],
                "r4": [f16(nimg, h4, w4, 96) for _ in range(4)],
                "r8": [f16(nimg, h8, w8, 128) for _ in range(4)],
                "stats": torch.empty(nimg, 128, 2, dtype=torch.float32, device=dev),
                # instance-norm partial sums: stats kernel (<= 296 parts x 128 ch) or conv epilogues (F_STATS, 64 ch)
                "scratch": torch.empty(nimg * max(296 * 128, ops.stats_parts(h2, w2, True) * 64) * 2, dtype=torch.float32, device=dev),
            }
        return self.enc[nimg]


class RAFTGMA(nn.Module):
    """B200-native GMA.  Same constructor contract as the reference (``network.py:27-43``): ``args`` is
    duck-typed (``mixed_precision``, ``num_heads``, ``position_only``, ``position_and_content``,
    ``__contains__``) and gets ``corr_levels``, ``corr_radius`` and ``dropout`` set on it."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.hidden_dim = self.context_dim = 128
        args.corr_levels = 4
        args.corr_radius = 4
        if "dropout" not in self.args:
            self.args.dropout = 0
        if getattr(args, "num_heads", 1) != 1 or getattr(args, "position_only", False) or \
                getattr(args, "position_and_content", False):
            raise NotImplementedError("only the configuration the SLAM uses is built: num_heads=1, content-only "
                                      "attention (atdn_vslam/utils/gma_parameters.py:8-10)")
        _build_module_tree(self, schema.gma_schema())
        self._packed = {}          # kernel-layout weights per device
        self._plans = {}
        self._graphs = {}          # captured forward() graphs per (batch, H, W, iters, device)
        self.generation = 0        # bumped whenever packed weights / plans are dropped: captured CUDA graphs hold raw pointers
        self.capture_forward = os.environ.get("ATDN_NO_FORWARD_GRAPH") != "1"
        # feature-map reuse across consecutive calls (the caller's loop passes the previous frame again as image1,
        # neural_slam.py:199-217): (plan key, previous image2 tensor -- kept alive so that its memory cannot be recycled --, version)
        self.reuse_fmap = os.environ.get("ATDN_NO_FMAP_REUSE") != "1"
        self._last_pair = None

    # -- weight handling --------------------------------------------------------------------------
    def _invalidate(self, plans=False):
        self._packed = {}
        self._graphs = {}
        self._last_pair = None
        if plans:
            self._plans = {}
        self.generation += 1

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._invalidate()
        return super().load_state_dict(schema.strip_module_prefix(state_dict), strict=strict, **kw)

    def _apply(self, fn, *a, **kw):
        self._invalidate(plans=True)
        return super()._apply(fn, *a, **kw)

    def freeze_bn(self):   # network.py:45-48: eval-mode BN is the only mode implemented
        return self

    def _weights(self, dev):
        key = str(dev)
        if key not in self._packed:
            sd = {k: v.detach().to(dev) for k, v in self.state_dict().items()}
            self._packed[key] = _Packed(sd)
        return self._packed[key]

    def _plan(self, b, h, w, dev):
        key = (b, h, w, str(dev))
        if key not in self._plans:
            self._plans[key] = _Plan(b, h, w, dev)
        return self._plans[key]

    # -- encoders -----------------------------------------------------------------------------------
    def _encoder(self, plan, ew, images, out_view, final_flags=0, h32=None, xpack=None):
        """BasicEncoder (extractor.py:165-189) on ``images`` [n,3,H,W] fp32 -> out_view [n,H8,W8,256].  ``xpack``: the
        row-packed normalised input of these images when another encoder call already produced it (both networks see
        2 * x / 255 - 1, network.py:75-76): the context net reuses the feature net's pack of the same frames."""
        n = images.shape[0]
        sc = plan.encoder_scratch(n, images.device)
        inst = ew.norm == "instance"
        relu = 0 if inst else L.F_RELU
        h2, w2 = plan.h // 2, plan.w // 2

        def norm_apply(x, y, resid=None, act=True, fused_parts=0):
            if fused_parts:   # the producing conv left per-tile partial sums in the scratch buffer (F_STATS)
                ops.inorm_finalize(sc["scratch"], fused_parts, sc["stats"], x.B, x.c, x.H * x.W)
            else:
                # ~1K pixels per partial sum: 113 / 28 / 8 CTAs per image at 1/2, 1/4, 1/8 resolution (measured optimum)
                ops.inorm_stats(x, sc["scratch"], max(8, min(296, (x.H * x.W) // 1024)), sc["stats"])
            ops.inorm_apply(x, sc["stats"], y, resid=resid, relu=act)

        # 64-channel layers at 1/2 resolution (70% of the normalised bytes): statistics come out of the conv epilogue
        fuse64 = inst and not _NO_FUSED_STATS
        st64 = {"flags": L.F_STATS, "aux32": sc["scratch"]} if fuse64 else {"flags": 0}
        if xpack is None:
            xpack = sc["xpack"]
            ops.stem_pack(images, xpack)
        r2 = sc["r2"]
        ops.conv_tc(View(xpack), ew.stem.wp, ew.stem.bias, View(r2[0]), cout=64, taps=(4, 1), pad=(2, 0), bn=64, mt=4,
                    flags=relu | st64["flags"], aux32=st64.get("aux32"), out_hw=(h2, w2))   # single CTAs: measured faster than pairs for K = 4 x 48
        x = View(r2[0])
        if inst:
            norm_apply(x, x, fused_parts=ops.stats_parts(h2, w2, False) if fuse64 else 0)
        pools = {64: sc["r2"], 96: sc["r4"], 128: sc["r8"]}
        for blk in ew.blocks:
            planes, st = blk["planes"], blk["stride"]
            free = [t for t in pools[planes] if t is not x.t]
            t1, t2, t3 = View(free[0]), View(free[1]), View(free[2])
            m_tiles = n * math.ceil(t1.H / 8) * math.ceil(t1.W / 16)
            bn = _pick_bn(planes, m_tiles)
            c1, c2 = blk["conv1"], blk["conv2"]
            f64 = fuse64 and planes == 64
            kw64 = {"aux32": sc["scratch"]} if f64 else {}
            parts64 = ops.stats_parts(t1.H, t1.W, _HALO_CFG[64][2]) if f64 else 0
            if st == 1:
                _conv_s1(x, c1, t1, cout=planes, taps=(3, 3), flags=relu | (L.F_STATS if f64 else 0), **kw64)
            else:
                ops.conv_tc(x, c1.wp, c1.bias, t1, cout=planes, taps=(3, 3), pad=(1, 1), stride=st, bn=bn, flags=relu)
            if inst:
                norm_apply(t1, t1, fused_parts=parts64 if st == 1 else 0)
            if st != 1:
                dn = blk["down"]
                ops.conv_tc(x, dn.wp, dn.bias, t3, cout=planes, taps=(1, 1), pad=(0, 0), stride=st, bn=bn)
                if inst:
                    norm_apply(t3, t3, act=False)
                skip = t3
            else:
                skip = x
            if inst:
                _conv_s1(t1, c2, t2, cout=planes, taps=(3, 3), flags=L.F_STATS if f64 else 0, **kw64)
                norm_apply(t2, t2, resid=skip, fused_parts=parts64)
            else:   # y = relu(bn(conv)); out = relu(skip + y) fused in the epilogue
                _conv_s1(t1, c2, t2, cout=planes, taps=(3, 3), flags=L.F_RELU | L.F_RESID, resid=skip)
            x = t2
        _conv_s1(x, ew.out, out_view, cout=256, taps=(1, 1), flags=final_flags, h32=h32)

    # -- forward ------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, image1, image2, iters=12, flow_init=None, upsample=True, test_mode=False):
        """Estimate optical flow between a pair (or batch of pairs) of frames -- network.py:72-129.

        The reference's caller invokes this once per frame at batch 1 (neural_slam.py:202); the ~330 kernel launches of
        one call would then be bound by launch overhead, so ``test_mode`` calls are captured into a CUDA graph per
        (shape, iters) the second time a shape is seen and replayed afterwards (``capture_forward = False`` or
        ATDN_NO_FORWARD_GRAPH=1 keeps every call eager).  Outputs are fresh tensors on every call."""
        L.require_cuda(image1, image2)
        dev = image1.device
        L.check(L.load().atdn_check_device(dev.index if dev.index is not None else torch.cuda.current_device()),
                "atdn_check_device")
        reuse = self._same_as_last_image2(image1)
        if test_mode and self.capture_forward and not torch.cuda.is_current_stream_capturing():
            out = self._forward_graphed(image1, image2, iters, flow_init, reuse)
        else:
            out = self._forward_eager(image1, image2, iters, flow_init, test_mode, reuse)
        self._last_pair = ((tuple(image1.shape), str(dev)), image2, image2._version) if self.reuse_fmap else None
        return out

    def _same_as_last_image2(self, image1):
        """True when ``image1`` is (a view of) the tensor that was ``image2`` of the previous call, unmodified: its feature
        map is still in the plan's buffer, so the feature net only has to run on the new frame (instance norm is per
        image: same values as running both frames)."""
        lp = self._last_pair
        if lp is None or not self.reuse_fmap:
            return False
        key, prev, version = lp
        return (key == (tuple(image1.shape), str(image1.device)) and prev.data_ptr() == image1.data_ptr() and prev.dtype == image1.dtype
                and prev.stride() == image1.stride() and prev._version == version == image1._version)

    def _forward_eager(self, image1, image2, iters, flow_init, test_mode, reuse=False):
        dev = image1.device
        b, _, h, w = image1.shape
        image1 = image1.float().contiguous()
        image2 = image2.float().contiguous()
        wts = self._weights(dev)
        plan = self._plan(b, h, w, dev)
        fmap = plan.buffer("fmap", (2 * b, plan.h8, plan.w8, 256), torch.float16)
        if reuse:
            # fmap[b:] still holds fnet(previous image2) = fnet(image1): move it, encode the new frame only
            fmap[:b].copy_(fmap[b:])
            self._encoder(plan, wts.fnet, image2, View(fmap[b:]))
            xpack = None
        else:
            # feature network on both frames as one batch of 2B (extractor.py:168-171)
            self._encoder(plan, wts.fnet, torch.cat([image1, image2], 0), View(fmap))
            xpack = plan.encoder_scratch(2 * b, dev)["xpack"][:b]
        return self._flow(plan, wts, image1, View(fmap[:b]), View(fmap[b:]), iters, flow_init, test_mode, xpack=xpack)

    def _forward_graphed(self, image1, image2, iters, flow_init, reuse=False):
        key = (tuple(image1.shape), image1.dtype, iters, str(image1.device), None if flow_init is None else tuple(flow_init.shape), reuse)
        g = self._graphs.get(key)
        if g is None:                       # first sighting of this shape: eager (packs weights, sizes the plan)
            self._graphs[key] = "seen"
            return self._forward_eager(image1, image2, iters, flow_init, True, reuse)
        if g == "seen":                     # second sighting: capture
            s1, s2 = image1.clone(), image2.clone()
            sf = None if flow_init is None else flow_init.clone()
            torch.cuda.synchronize(image1.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._forward_eager(s1, s2, iters, sf, True, reuse)
            g = self._graphs[key] = (graph, s1, s2, sf, out)
        graph, s1, s2, sf, out = g
        s1.copy_(image1)
        s2.copy_(image2)
        if sf is not None:
            sf.copy_(flow_init)
        graph.replay()
        return out[0].clone(), out[1].clone()

    @torch.no_grad()
    def forward_frames(self, frames, iters=12, test_mode=True):
        """Flow for the B consecutive pairs (t, t+1) of ``frames`` [B+1,3,H,W]: the feature network
        runs once per FRAME (instance norm is per image, so the result equals B separate ``forward``
        calls) and fmap[t+1] doubles as fmap2 of pair t and fmap1 of pair t+1."""
        L.require_cuda(frames)
        dev = frames.device
        nb, _, h, w = frames.shape
        b = nb - 1
        frames = frames.float().contiguous()
        wts = self._weights(dev)
        plan = self._plan(b, h, w, dev)
        fmap = plan.buffer("fmap_seq", (nb, plan.h8, plan.w8, 256), torch.float16)
        self._encoder(plan, wts.fnet, frames, View(fmap))
        return self._flow(plan, wts, frames[:b], View(fmap[:b]), View(fmap[1:]), iters, None, test_mode,
                          xpack=plan.encoder_scratch(nb, dev)["xpack"][:b])

    def _flow(self, plan, wts, image1, fmap1, fmap2, iters, flow_init, test_mode, xpack=None):
        dev = image1.device
        b, h, w = plan.b, plan.h, plan.w
        h8, w8, n, np_ = plan.h8, plan.w8, plan.n, plan.np_
        m_tiles = b * math.ceil(h8 / 8) * math.ceil(w8 / 16)
        # fp32-accumulated all-pairs correlation pyramid (corr.py:16-30), rounded to fp16 on store
        ops.corr_pyramid_build(fmap1, fmap2, plan.pyr, alpha=1.0 / (math.sqrt(256.0) * FMAP_SCALE * FMAP_SCALE))

        # context network: net = tanh(.) -> HX[0:128] + h32, inp = relu(.) -> HX[128:256]
        hx = plan.hx
        self._encoder(plan, wts.cnet, image1, View(hx, 0, 256), final_flags=L.F_TANH_LO | _H16, h32=plan.h32, xpack=xpack)

        if wts.gru_pre:
            per_buf = plan.state_numel
            for i, (taps, pad) in enumerate((((1, 5), (0, 2)), ((5, 1), (2, 0)))):
                zr, q = wts.gru_pre[i]
                _conv_s1(View(hx, 128, 128), zr, View(plan.pre_zr[i].view(-1, 1, 1, 8)), cout=256, taps=taps, epi=L.EPI_STORE32,
                         flags=L.F_TILED32 | _PRE16, out_pitch=per_buf)
                _conv_s1(View(hx, 128, 128), q, View(plan.pre_q[i].view(-1, 1, 1, 8)), cout=128, taps=taps, epi=L.EPI_STORE32,
                         flags=L.F_TILED32 | _PRE16, out_pitch=per_buf)

        self._attention(plan, wts)

        if flow_init is not None:
            flow_init = flow_init.float().contiguous()
        ops.coords_init(plan.coords1, plan.flow, flow_init)

        preds = []
        flow_up = flow_lo = None
        for itr in range(iters):
            need_up = (not test_mode) or itr == iters - 1
            self._update(plan, wts, m_tiles)
            if need_up:
                c = wts.mask0
                _conv_s1(View(hx, 0, 128), c, View(plan.mh), cout=256, taps=(3, 3), flags=L.F_RELU)
                c = wts.mask2
                d_out = View(plan.mask32.view(b, h8, w8, 576))
                _conv_s1(View(plan.mh), c, d_out, cout=576, taps=(1, 1), epi=L.EPI_STORE32 if _MASK32 else L.EPI_STORE16, alpha=0.25)
                flow_up = torch.empty(b, 2, h, w, dtype=torch.float32, device=dev)
                flow_lo = torch.empty(b, 2, h8, w8, dtype=torch.float32, device=dev)
                ops.convex_upsample(plan.mask32, plan.flow, flow_up, flow_lo)
                preds.append(flow_up)
        if test_mode:
            return flow_lo, flow_up
        return preds

    def _attention(self, plan, wts):
        """Attention.forward (gma.py:54-76) on the context features HX[128:256]: q.k^T * scale -> un-normalised softmax
        numerators P (fp16) + 1 / row sums."""
        _conv_s1(View(plan.hx, 128, 128), wts.to_qk, View(plan.qk), cout=256, taps=(1, 1))
        ops.attn_probs(plan.qk, plan.p16, plan.inv_sum, 128 ** -0.5, tiled=_P_TILED, block_hot=plan.p_hot if plan.mixed else None,
                       hot_energy=_P_HOT_ENERGY)
        if plan.mixed:
            ops.attn_harmonize(plan.p16, plan.p_hot, plan.p_hot2, plan.n)
            if L.PROFILER is not None:    # bench.py's byte accounting of the mixed P.V stream (one host read, profiling only)
                L.PV_HOT_FRACTION = float(plan.p_hot2.float().mean())

    def _aggregate(self, plan, wts):
        """Aggregate.forward (gma.py:102-115) on the motion features HX[256:384] -> HX[384:512]:
        v^T = W_v . mf^T (stored [B,128,Np]); mfg = mf + gamma * (P . v) / rowsum."""
        b, n, np_, hx = plan.b, plan.n, plan.np_, plan.hx
        d = L.TcDesc()
        d.bn, d.epi, d.flags, d.a_mode, d.b_mode = 128, L.EPI_STORE16, L.F_B_BATCHED | L.F_A_SHARED, L.MODE_ROWS, L.MODE_ROWS
        d.a = L.ptr(wts.to_v)
        L._set(d.a_dims, (128, 128, 1, 1))
        L._set(d.a_strides, (128, 128 * 128, 128 * 128))
        d.b = L.ptr(hx, 256)
        L._set(d.b_dims, (128, n, 1, b))
        L._set(d.b_strides, (512, n * 512, n * 512))
        d.n_valid, d.alpha = n, 1.0
        d.out, d.out_pitch = L.ptr(plan.vt), np_
        if plan.mixed:
            d.out8 = L.ptr(plan.v8)
        L.tc_gemm(d)
        small = b * math.ceil(n / 128) <= _SMALL_TILES
        ops.gemm_rows(L.ptr(plan.p16), np_ if _P_TILED else n, n, np_, b, L.ptr(plan.vt), 128, np_, L.ptr(hx, 384), 512, n_valid=128,
                      b_bstride=128 * np_, bn=64 if small else 128, epi=L.EPI_PV, b_k=n,
                      flags=(L.F_PAIR if ((_PV_PAIR and not small) or plan.mixed) else 0) | (L.F_A_TILED if _P_TILED else 0) |
                      (L.F_A_MIXED if plan.mixed else 0),
                      resid_ptr=L.ptr(hx, 256), resid_pitch=512, aux32=plan.inv_sum, gamma=wts.gamma,
                      b8=plan.v8 if plan.mixed else None, a_hot=plan.p_hot2 if plan.mixed else None)

    def _update(self, plan, wts, m_tiles):
        """One refinement iteration: lookup + GMAUpdateBlock (update.py:127-139) + coords update."""
        b, h8, w8, n, np_ = plan.b, plan.h8, plan.w8, plan.n, plan.np_
        hx = plan.hx
        R = L.F_RELU
        ops.corr_lookup(plan.pyr, plan.coords1, out16=View(plan.corrfeat))
        c = wts.convc1
        _conv_s1(View(plan.corrfeat, 0, 324), c, View(plan.c1), cout=256, taps=(1, 1), flags=R)
        c = wts.convc2
        _conv_s1(View(plan.c1), c, View(plan.corflo, 0, 192), cout=192, taps=(3, 3), flags=R)
        ops.flow_pack(plan.flow, plan.fpack)
        c = wts.convf1
        small = _is_small(View(plan.f1))
        ops.conv_tc(View(plan.fpack, 0, 14), c.wp, c.bias, View(plan.f1), cout=128, taps=(7, 1), pad=(3, 0), bn=64 if small else 128,
                    mt=1 if small else 2, flags=R | (0 if small else L.F_PAIR))
        c = wts.convf2
        ops.conv_tc(View(plan.f1), c.wp, c.bias, View(plan.corflo, 192, 64), cout=64, taps=(3, 3), pad=(1, 1), bn=64, mt=1 if small else 2,
                    flags=R | (0 if small else L.F_PAIR))
        c = wts.conv
        _conv_s1(View(plan.corflo), c, View(hx, 256, 128), cout=128, taps=(3, 3), flags=R | L.F_FLOWTAIL, aux32=plan.flow)
        self._aggregate(plan, wts)
        for i, ((zr, q), taps, pad) in enumerate(((wts.gru[0], (1, 5), (0, 2)), (wts.gru[1], (5, 1), (2, 0)))):
            if wts.gru_pre:   # contract [h | mf | mfg] only; the context term comes from plan.pre_*
                _conv_s1(View(hx, 0, 128), zr, None, cout=256, taps=taps, epi=L.EPI_GRU_ZR, a2=View(hx, 256, 256), h32=plan.h32,
                         z32=plan.z, rh16=plan.rh, aux32=plan.pre_zr[i], aux_half_offset=plan.state_numel, flags=_PRE16 | _Z16 | _H16)
                _conv_s1(View(plan.rh), q, View(hx, 0, 128), cout=128, taps=taps, epi=L.EPI_GRU_Q, a2=View(hx, 256, 256),
                         h32=plan.h32, z32=plan.z, aux32=plan.pre_q[i], flags=_PRE16 | _Z16 | _H16)
            else:
                _conv_s1(View(hx), zr, None, cout=256, taps=taps, epi=L.EPI_GRU_ZR, h32=plan.h32, z32=plan.z, rh16=plan.rh,
                         flags=_Z16 | _H16)
                _conv_s1(View(plan.rh), q, View(hx, 0, 128), cout=128, taps=taps, epi=L.EPI_GRU_Q, a2=View(hx, 128, 384),
                         h32=plan.h32, z32=plan.z, flags=_Z16 | _H16)
        c = wts.fh1
        _conv_s1(View(hx, 0, 128), c, View(plan.fh), cout=256, taps=(3, 3), flags=R)
        c = wts.fh2   # flow_head.conv2 + coords update (network.py:111,116): per-tap 1x1 products, then the shifted sum
        ops.conv_tc(View(plan.fh), c.wp, None, View(plan.fh2d), cout=18, taps=(1, 1), pad=(0, 0), bn=32, mt=4, epi=L.EPI_STORE32)
        ops.flow_head_gather(plan.fh2d, wts.fh2_bias, plan.coords1, plan.flow)


class CorrBlock:
    """Drop-in for ``GMA.whl!/GMA/core/corr.py:15-63``: ``CorrBlock(fmap1, fmap2, num_levels=4, radius=4)``
    builds the pyramid on the tensor cores; ``corr_fn(coords) -> [B, 324, H8, W8]`` fp32."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        if num_levels != 4 or radius != 4:
            raise NotImplementedError("the SLAM configuration is num_levels=4, radius=4 (network.py:33-34)")
        L.require_cuda(fmap1, fmap2)
        self.num_levels, self.radius = num_levels, radius
        b, c, h, w = fmap1.shape
        f1 = fmap1.permute(0, 2, 3, 1).contiguous().half()
        f2 = fmap2.permute(0, 2, 3, 1).contiguous().half()
        self.levels = ops.alloc_pyramid(b, h, w, fmap1.device)
        ops.corr_pyramid_build(View(f1), View(f2), self.levels)
        self.shape = (b, h, w)

    @staticmethod
    def corr(fmap1, fmap2):
        """corr.py:55-63: the all-pairs volume alone, [B, H, W, 1, H, W] fp32 = <fmap1, fmap2> / sqrt(C) (level 0 of the
        pyramid the tensor-core kernel builds)."""
        b, c, h, w = fmap1.shape
        blk = CorrBlock(fmap1, fmap2)
        return blk.corr_pyramid[0].reshape(b, h, w, 1, h, w)

    @property
    def corr_pyramid(self):
        """The reference attribute: list of [B*N, 1, H_l, W_l] fp32 tensors."""
        out = []
        for t, (hl, wl, _) in zip(self.levels, ops.pyramid_shapes(self.shape[1], self.shape[2])):
            out.append(t[:, :, :wl].unsqueeze(1))
        return out

    def __call__(self, coords):
        b, h, w = self.shape
        c = coords.float().permute(0, 2, 3, 1).contiguous()
        out32 = torch.empty(b * h * w, 324, dtype=torch.float32, device=coords.device)
        ops.corr_lookup(self.levels, c, out32=out32)
        return out32.view(b, h, w, 324).permute(0, 3, 1, 2).contiguous()
