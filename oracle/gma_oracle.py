"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference GMA flow network.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product path (``atdn_vslam_b200``) never does.

Parity pin: validated in the build container against the reference itself
(``/root/reference/GMA-1.0.0-py3-none-any.whl``, GMA 1.0.0) by ``tests/golden/make_golden.py``;
its outputs are committed under ``tests/golden/`` and re-checked by
``tests/test_oracle_golden.py``.  The reference ships no golden vectors of its own.

Everything is a pure function of a reference-format state dict (plain or ``module.``-prefixed) and
runs in fp32 on CPU -- this is the reference's ``device="cpu"`` path, where
``torch.cuda.amp.autocast`` disables itself.  File:line citations are relative to
``GMA.whl!/GMA/core/``.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def _strip(sd):
    if all(k.startswith("module.") for k in sd):
        return {k[7:]: v for k, v in sd.items()}
    return sd


# ----------------------------------------------------------------------------------------------
# encoders -- extractor.py:6-56 (ResidualBlock), :116-189 (BasicEncoder)
# ----------------------------------------------------------------------------------------------
def _norm(x, sd, name, norm):
    if norm == "instance":      # nn.InstanceNorm2d: no affine, no running stats, eps 1e-5
        return F.instance_norm(x, eps=1e-5)
    if norm == "batch":         # eval-mode batch norm
        return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                            sd[name + ".weight"], sd[name + ".bias"], training=False, eps=1e-5)
    raise ValueError(norm)


def _conv(x, sd, name, stride=1, padding=0):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def residual_block(x, sd, p, norm, stride):
    """extractor.py:47-55"""
    y = F.relu(_norm(_conv(x, sd, p + "conv1", stride, 1), sd, p + "norm1", norm))
    y = F.relu(_norm(_conv(y, sd, p + "conv2", 1, 1), sd, p + "norm2", norm))
    if stride != 1:
        # downsample = Sequential(conv1x1 stride, norm3); batch norm3 lives at downsample.1
        x = _norm(_conv(x, sd, p + "downsample.0", stride, 0), sd, p + "downsample.1", norm)
    return F.relu(x + y)


def basic_encoder(x, sd, p, norm):
    """extractor.py:165-189 (eval mode: dropout inactive)."""
    x = F.relu(_norm(_conv(x, sd, p + "conv1", 2, 3), sd, p + "norm1", norm))
    for li, stride in ((1, 1), (2, 2), (3, 2)):
        x = residual_block(x, sd, f"{p}layer{li}.0.", norm, stride)
        x = residual_block(x, sd, f"{p}layer{li}.1.", norm, 1)
    return _conv(x, sd, p + "conv2")


# ----------------------------------------------------------------------------------------------
# correlation volume, pyramid, lookup -- corr.py:16-63, utils/utils.py:59-73
# ----------------------------------------------------------------------------------------------
def corr_volume(fmap1, fmap2):
    """corr.py:55-63: corr[b,i,j] = sum_c f1[b,c,i] f2[b,c,j] / sqrt(C) -> [B, N, H, W]."""
    b, c, h, w = fmap1.shape
    f1 = fmap1.reshape(b, c, h * w)
    f2 = fmap2.reshape(b, c, h * w)
    corr = torch.matmul(f1.transpose(1, 2), f2) / math.sqrt(c)
    return corr.reshape(b, h * w, h, w)


def corr_pyramid(fmap1, fmap2, num_levels=4):
    """corr.py:16-30: level l+1 = avg_pool2d(level l, 2, stride 2) (floor).  Each level is
    [B, N, Hl, Wl] (the reference flattens B*N into the batch dim, same memory order)."""
    lvl = corr_volume(fmap1, fmap2)
    pyr = [lvl]
    for _ in range(num_levels - 1):
        b, n, h, w = lvl.shape
        h2, w2 = h // 2, w // 2
        v = lvl[:, :, :2 * h2, :2 * w2].reshape(b, n, h2, 2, w2, 2)
        lvl = v.sum(dim=(3, 5)) * 0.25
        pyr.append(lvl)
    return pyr


def corr_lookup(pyramid, coords, radius=4):
    """corr.py:32-53 + utils.py:59-73 restated as an explicit gather.

    For query pixel p with target coordinate (x, y) = coords[b, :, p] and level l, window index
    (a, b) in 0..2r: sample level l at (x / 2^l + (a - r), y / 2^l + (b - r)) -- the *slow* window
    index a offsets x (corr.py:40-46: ``meshgrid(dy, dx)`` is added to (x, y) coordinates).
    Bilinear, align_corners=True (pixel coordinates), zeros outside.  Output channel
    = l*(2r+1)^2 + a*(2r+1) + b, shape [B, L*(2r+1)^2, H, W] fp32.
    """
    bsz, _, h1, w1 = coords.shape
    n = h1 * w1
    k = 2 * radius + 1
    cx = coords[:, 0].reshape(bsz, n, 1, 1)
    cy = coords[:, 1].reshape(bsz, n, 1, 1)
    d = torch.arange(-radius, radius + 1, dtype=coords.dtype, device=coords.device)
    out = []
    for lvl, corr in enumerate(pyramid):
        hl, wl = corr.shape[-2:]
        # grid_sample sees normalised coordinates; reproduce its round trip exactly
        # (utils.py:63-64 then ATen's align_corners=True un-normalisation).
        x = cx / (2 ** lvl) + d.view(1, 1, k, 1)
        y = cy / (2 ** lvl) + d.view(1, 1, 1, k)
        if wl > 1:
            x = ((2 * x / (wl - 1) - 1) + 1) / 2 * (wl - 1)
        if hl > 1:
            y = ((2 * y / (hl - 1) - 1) + 1) / 2 * (hl - 1)
        x = x.expand(bsz, n, k, k)
        y = y.expand(bsz, n, k, k)
        x0 = torch.floor(x)
        y0 = torch.floor(y)
        fx = x - x0
        fy = y - y0
        x0 = x0.long()
        y0 = y0.long()
        flat = corr.reshape(bsz, n, hl * wl)
        acc = torch.zeros(bsz, n, k, k, dtype=corr.dtype, device=corr.device)
        for dy_, dx_, wgt in ((0, 0, (1 - fx) * (1 - fy)), (0, 1, fx * (1 - fy)),
                              (1, 0, (1 - fx) * fy), (1, 1, fx * fy)):
            xi = x0 + dx_
            yi = y0 + dy_
            ok = (xi >= 0) & (xi < wl) & (yi >= 0) & (yi < hl)
            idx = (yi.clamp(0, hl - 1) * wl + xi.clamp(0, wl - 1)).reshape(bsz, n, k * k)
            v = torch.gather(flat, 2, idx).reshape(bsz, n, k, k)
            acc = acc + torch.where(ok, v * wgt, torch.zeros_like(v))
        out.append(acc.reshape(bsz, h1, w1, k * k))
    out = torch.cat(out, dim=-1)
    return out.permute(0, 3, 1, 2).contiguous().float()


def corr_pyramid_aten(fmap1, fmap2, num_levels=4):
    """corr.py:16-30 with the SAME ATen ops the reference calls (``avg_pool2d`` on the [B*N,1,H,W] volume) -- the form
    the CPU baseline is timed on, so that the port is not slower than the reference; equal to :func:`corr_pyramid`
    (checked in tests/test_oracle_golden.py).  Levels are returned as [B, N, Hl, Wl] views."""
    lvl = corr_volume(fmap1, fmap2)
    b, n, h, w = lvl.shape
    flat = lvl.reshape(b * n, 1, h, w)
    pyr = [lvl]
    for _ in range(num_levels - 1):
        flat = F.avg_pool2d(flat, 2, stride=2)
        pyr.append(flat.reshape(b, n, flat.shape[-2], flat.shape[-1]))
    return pyr


def corr_lookup_aten(pyramid, coords, radius=4):
    """corr.py:32-53 + utils.py:59-73 through ``F.grid_sample`` (bilinear, zeros padding, align_corners=True), the op
    the reference itself calls; same result as the explicit gather of :func:`corr_lookup` up to fp32 rounding of the
    normalise / de-normalise round trip.  Used for the timed CPU baseline."""
    bsz, _, h1, w1 = coords.shape
    n = h1 * w1
    k = 2 * radius + 1
    d = torch.arange(-radius, radius + 1, dtype=coords.dtype, device=coords.device)
    centre = coords.permute(0, 2, 3, 1).reshape(bsz * n, 1, 1, 2)
    # window offset [a, b] -> (x + d[a], y + d[b]): the slow window index offsets x (see corr_lookup)
    offs = torch.stack([d.view(k, 1).expand(k, k), d.view(1, k).expand(k, k)], dim=-1).view(1, k, k, 2)
    out = []
    for lvl, corr in enumerate(pyramid):
        hl, wl = corr.shape[-2:]
        pos = centre / (2 ** lvl) + offs
        grid = torch.stack([2 * pos[..., 0] / (wl - 1) - 1, 2 * pos[..., 1] / (hl - 1) - 1], dim=-1)
        smp = F.grid_sample(corr.reshape(bsz * n, 1, hl, wl), grid, mode="bilinear", padding_mode="zeros", align_corners=True)
        out.append(smp.reshape(bsz, h1, w1, k * k))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def coords_grid(batch, ht, wd, device=None):
    """utils.py:76-79: channel 0 = x (column index), channel 1 = y (row index)."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


# ----------------------------------------------------------------------------------------------
# attention / aggregation -- gma.py:54-76, 102-115
# ----------------------------------------------------------------------------------------------
def attention(inp, sd, heads=1):
    """gma.py:54-76 with position_only = position_and_content = False -> [B, heads, N, N]."""
    b, c, h, w = inp.shape
    qk = F.conv2d(inp, sd["att.to_qk.weight"])
    q, k = qk.chunk(2, dim=1)
    dh = q.shape[1] // heads
    q = q.reshape(b, heads, dh, h * w).transpose(2, 3) * (dh ** -0.5)
    k = k.reshape(b, heads, dh, h * w).transpose(2, 3)
    sim = torch.matmul(q, k.transpose(2, 3))
    return sim.softmax(dim=-1)


def aggregate(attn, fmap, sd, heads=1):
    """gma.py:102-115; dim == inner_dim so project is None."""
    b, c, h, w = fmap.shape
    v = F.conv2d(fmap, sd["update_block.aggregator.to_v.weight"])
    dh = v.shape[1] // heads
    v = v.reshape(b, heads, dh, h * w).transpose(2, 3)
    out = torch.matmul(attn, v)                                   # [B, heads, N, dh]
    out = out.transpose(2, 3).reshape(b, heads * dh, h, w)
    return fmap + sd["update_block.aggregator.gamma"] * out


# ----------------------------------------------------------------------------------------------
# update block -- update.py:7-15, 36-63, 66-84, 113-139
# ----------------------------------------------------------------------------------------------
def motion_encoder(flow, corr, sd):
    """update.py:76-84"""
    p = "update_block.encoder."
    cor = F.relu(_conv(corr, sd, p + "convc1"))
    cor = F.relu(_conv(cor, sd, p + "convc2", 1, 1))
    flo = F.relu(_conv(flow, sd, p + "convf1", 1, 3))
    flo = F.relu(_conv(flo, sd, p + "convf2", 1, 1))
    out = F.relu(_conv(torch.cat([cor, flo], dim=1), sd, p + "conv", 1, 1))
    return torch.cat([out, flow], dim=1)


def sep_conv_gru(h, x, sd):
    """update.py:48-63: horizontal (1x5) then vertical (5x1) GRU pass."""
    p = "update_block.gru."
    for n, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([h, x], dim=1)
        z = torch.sigmoid(_conv(hx, sd, p + "convz" + n, 1, pad))
        r = torch.sigmoid(_conv(hx, sd, p + "convr" + n, 1, pad))
        q = torch.tanh(_conv(torch.cat([r * h, x], dim=1), sd, p + "convq" + n, 1, pad))
        h = (1 - z) * h + z * q
    return h


def update_block(net, inp, corr, flow, attn, sd):
    """update.py:127-139 -> (net, mask, delta_flow)"""
    mf = motion_encoder(flow, corr, sd)
    mfg = aggregate(attn, mf, sd)
    net = sep_conv_gru(net, torch.cat([inp, mf, mfg], dim=1), sd)
    p = "update_block."
    delta = _conv(F.relu(_conv(net, sd, p + "flow_head.conv1", 1, 1)), sd, p + "flow_head.conv2", 1, 1)
    mask = 0.25 * _conv(F.relu(_conv(net, sd, p + "mask.0", 1, 1)), sd, p + "mask.2")
    return net, mask, delta


def upsample_flow(flow, mask):
    """network.py:59-70: convex combination over the 3x3 neighbourhood of 8*flow."""
    n, _, h, w = flow.shape
    mask = torch.softmax(mask.reshape(n, 1, 9, 8, 8, h, w), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).reshape(n, 2, 9, 1, 1, h, w)
    up = torch.sum(mask * up, dim=2)                              # [n,2,8,8,h,w]
    return up.permute(0, 1, 4, 2, 5, 3).reshape(n, 2, 8 * h, 8 * w)


# ----------------------------------------------------------------------------------------------
# whole network -- network.py:72-129
# ----------------------------------------------------------------------------------------------
def raftgma_forward(sd, image1, image2, iters=12, flow_init=None, test_mode=True, return_intermediates=False, aten_ops=False,
                    mixed_precision=False):
    """Reference ``RAFTGMA.forward``.  Default: the fp32 path the reference takes with ``device="cpu"`` (O-cpu, the parity
    oracle).  ``test_mode`` returns (coords1 - coords0, flow_up); otherwise the list of per-iteration ``flow_up``.
    ``aten_ops``: pyramid and lookup through ``avg_pool2d`` / ``grid_sample`` like the reference (the timed CPU
    baseline) instead of the explicit restatements (the checker).
    ``mixed_precision``: the reference's CUDA path (O-cuda): the same three
    ``autocast(enabled=args.mixed_precision)`` regions as network.py:85,93,112 (fp16), the fmaps cast back to fp32 before
    the correlation (network.py:88-89), lookup / coords / convex upsampling outside autocast.  Runs on whatever device
    the inputs and ``sd`` live on; used to MEASURE the reference's own fp16-vs-fp32 gap, never as the parity target."""
    sd = _strip(sd)
    dev = image1.device

    def region():
        return torch.autocast(device_type=dev.type, dtype=torch.float16, enabled=mixed_precision)

    with torch.no_grad():
        im1 = (2 * (image1.float() / 255.0) - 1.0).contiguous()
        im2 = (2 * (image2.float() / 255.0) - 1.0).contiguous()
        b = im1.shape[0]
        with region():
            fm = basic_encoder(torch.cat([im1, im2], 0), sd, "fnet.", "instance")
        fmap1, fmap2 = fm[:b].float(), fm[b:].float()
        pyr = (corr_pyramid_aten if aten_ops else corr_pyramid)(fmap1, fmap2)
        lookup = corr_lookup_aten if aten_ops else corr_lookup
        with region():
            cnet = basic_encoder(im1, sd, "cnet.", "batch")
            net, inp = torch.split(cnet, [128, 128], dim=1)
            net = torch.tanh(net)
            inp = torch.relu(inp)
            attn = attention(inp, sd)
        h8, w8 = im1.shape[2] // 8, im1.shape[3] // 8
        coords0 = coords_grid(b, h8, w8, dev)
        coords1 = coords_grid(b, h8, w8, dev)
        if flow_init is not None:
            coords1 = coords1 + flow_init
        preds = []
        inter = {"fmap1": fmap1, "fmap2": fmap2, "net0": net, "inp": inp, "corr": [], "delta": []}
        for _ in range(iters):
            corr = lookup(pyr, coords1)
            flow = coords1 - coords0
            with region():
                net, mask, delta = update_block(net, inp, corr, flow, attn, sd)
            coords1 = coords1 + delta
            flow_up = upsample_flow(coords1 - coords0, mask)
            preds.append(flow_up)
            if return_intermediates:
                inter["corr"].append(corr)
                inter["delta"].append(delta)
        if return_intermediates:
            inter.update(pyramid=pyr, attn=attn, net=net, mask=mask)
            return coords1 - coords0, flow_up, inter
        if test_mode:
            return coords1 - coords0, flow_up
        return preds


def input_pad(h, w):
    """utils.py:8-19 (mode='sintel'): pad so H, W are multiples of 8 -> [left, right, top, bottom]."""
    ph = (((h // 8) + 1) * 8 - h) % 8
    pw = (((w // 8) + 1) * 8 - w) % 8
    return [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]


# ----------------------------------------------------------------------------------------------
# numpy micro-restatement of the lookup for tiny cases (independent of torch indexing)
# ----------------------------------------------------------------------------------------------
def corr_lookup_loops(pyramid, coords, radius=4):
    """Pure-Python loop version of :func:`corr_lookup` (tiny inputs only)."""
    pyr = [p.numpy() for p in pyramid]
    c = coords.numpy()
    bsz, _, h1, w1 = c.shape
    k = 2 * radius + 1
    out = np.zeros((bsz, len(pyr) * k * k, h1, w1), dtype=np.float32)
    for b in range(bsz):
        for py in range(h1):
            for px in range(w1):
                q = py * w1 + px
                for lvl, corr in enumerate(pyr):
                    hl, wl = corr.shape[-2:]
                    for a in range(k):
                        for bb in range(k):
                            x = np.float32(c[b, 0, py, px] / np.float32(2 ** lvl) + np.float32(a - radius))
                            y = np.float32(c[b, 1, py, px] / np.float32(2 ** lvl) + np.float32(bb - radius))
                            x0, y0 = int(np.floor(x)), int(np.floor(y))
                            fx, fy = np.float32(x - x0), np.float32(y - y0)
                            v = np.float32(0)
                            for dy_, dx_, wg in ((0, 0, (1 - fx) * (1 - fy)), (0, 1, fx * (1 - fy)),
                                                 (1, 0, (1 - fx) * fy), (1, 1, fx * fy)):
                                xi, yi = x0 + dx_, y0 + dy_
                                if 0 <= xi < wl and 0 <= yi < hl:
                                    v += np.float32(wg) * corr[b, q, yi, xi]
                            out[b, lvl * k * k + a * k + bb, py, px] = v
    return torch.from_numpy(out)
