"""Seeded synthetic weights and KITTI-shaped frames.

The reference checkpoints are absent (``.MISSING_LARGE_BLOBS``), so parity and benchmarks use
weights drawn here from a CPU ``torch.Generator`` (bit-reproducible for a given torch build)
in the reference's state-dict format.  Hazards handled:
``aggregator.gamma`` is non-zero, batch-norm running statistics are away from (0, 1).
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict

import torch

from . import schema

GMA_SEED, ATDNVO_SEED, VAE_SEED, FRAME_SEED = 0, 1, 2, 1234


def _draw(shape, kind, g, gain):
    if kind == "conv_w":
        fan_in = shape[1] * shape[2] * shape[3]
        return torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
    if kind == "lin_w":
        return torch.randn(shape, generator=g) * (gain / math.sqrt(shape[1]))
    if kind == "bias":
        return torch.randn(shape, generator=g) * 0.05
    if kind == "bn_w":
        return 0.75 + 0.5 * torch.rand(shape, generator=g)
    if kind == "bn_b":
        return torch.randn(shape, generator=g) * 0.1
    if kind == "bn_mean":
        return torch.randn(shape, generator=g) * 0.1
    if kind == "bn_var":
        return 0.6 + 0.8 * torch.rand(shape, generator=g)
    if kind == "bn_count":
        return torch.tensor(1000, dtype=torch.int64)
    if kind == "gamma":
        return torch.full(shape, 0.5)
    if kind == "emb":
        return torch.randn(shape, generator=g)
    if kind == "index":
        n = shape[0]
        return torch.arange(n).view(1, -1) - torch.arange(n).view(-1, 1) + n - 1
    raise ValueError(kind)


def make_state_dict(sch, seed, gain=1.2, module_prefix=False):
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, (shape, kind) in sch.items():
        t = _draw(tuple(shape), kind, g, gain)
        if name.endswith("flow_head.conv2.weight") or name.endswith("flow_head.conv2.bias"):
            t = t * 0.08      # keep the random-weight flow in a sane regime (~1 px / iteration at 1/8)
        sd[("module." + name) if module_prefix else name] = t
    return sd


def gma_state_dict(seed=GMA_SEED, module_prefix=False):
    return make_state_dict(schema.gma_schema(), seed, gain=1.2, module_prefix=module_prefix)


SEQUENCE_POSE_GAIN = (1.0, 40.0)   # (rotation, translation) head gains of the golden SEQUENCE fixture


def atdnvo_state_dict(seed=ATDNVO_SEED, pose_gain=None):
    """``pose_gain`` = (rotation gain, translation gain) scales the last (bias-free) layer of the two regressor heads:
    the random-weight network then moves far enough per frame for the keyframe rule (neural_slam.py:288-302: 10 degrees or
    15 units since the last keyframe) to fire every few frames (tests/golden/make_golden_sequence.py)."""
    sd = make_state_dict(schema.atdnvo_schema(), seed, gain=1.0)
    if pose_gain is not None:
        sd["rotation_regressor.2.weight"] = sd["rotation_regressor.2.weight"] * pose_gain[0]
        sd["translation_regressor.2.weight"] = sd["translation_regressor.2.weight"] * pose_gain[1]
    return sd


def vae_state_dict(seed=VAE_SEED):
    return make_state_dict(schema.vae_encoder_schema(), seed, gain=1.0)


def state_dict_digest(sd):
    """sha256 over the fp32 bytes of every floating tensor in name order (fixture guard)."""
    h = hashlib.sha256()
    for k in sorted(sd):
        v = sd[k]
        if v.is_floating_point():
            h.update((k[7:] if k.startswith("module.") else k).encode())
            h.update(v.detach().cpu().float().contiguous().numpy().tobytes())
    return h.hexdigest()


# ----------------------------------------------------------------------------------------------
# frames
# ----------------------------------------------------------------------------------------------
def texture_canvas(height, width, seed=FRAME_SEED, margin=64):
    """Low-frequency texture: bicubic up-sampling of rand(3, H/8, W/8) to a canvas larger than the
    frame, rescaled to 0..255 (SURVEY.md §8(d) 'synthetic frames')."""
    g = torch.Generator().manual_seed(seed)
    ch, cw = height + 2 * margin, width + 2 * margin
    low = torch.rand(1, 3, max(4, ch // 8), max(4, cw // 8), generator=g)
    fine = torch.rand(1, 3, max(4, ch // 2), max(4, cw // 2), generator=g)
    canvas = torch.nn.functional.interpolate(low, size=(ch, cw), mode="bicubic", align_corners=False)
    canvas = canvas + 0.25 * torch.nn.functional.interpolate(fine, size=(ch, cw), mode="bicubic",
                                                             align_corners=False)
    canvas = (canvas - canvas.amin()) / (canvas.amax() - canvas.amin())
    return (canvas[0] * 255.0).contiguous()


def frame_sequence(num_frames, height=376, width=1232, seed=FRAME_SEED, max_shift=6.0, integer=True, start=0, dtype=torch.float32,
                   indices=None):
    """``num_frames`` crops of one canvas at a smoothly varying offset (<= max_shift px/frame) so that
    consecutive pairs have real sub-window displacements.  Returns float32 [T,3,H,W] holding 0..255
    (integer-valued when ``integer``: the reference consumes float tensors of uint8 range).
    ``start``: return frames start .. start + num_frames - 1 of the (unbounded) sequence -- a rank's shard of a long
    sequence without materialising the rest; ``indices`` (ascending list, may repeat) picks arbitrary frames instead
    (``num_frames`` / ``start`` are then ignored); ``dtype=torch.uint8`` stores the (integer) frames as bytes."""
    margin = 64
    canvas = texture_canvas(height, width, seed, margin)
    frames = []
    ox, oy = float(margin), float(margin)
    if indices is not None:
        indices = [int(i) for i in indices]
        assert all(a <= b for a, b in zip(indices, indices[1:])), "indices must be ascending"
        count = {}
        for i in indices:
            count[i] = count.get(i, 0) + 1
        start, num_frames = (indices[0], indices[-1] - indices[0] + 1) if indices else (0, 0)
    for t in range(start + num_frames):
        # deterministic smooth trajectory
        ox += max_shift * math.sin(0.37 * t + 0.5)
        oy += 0.5 * max_shift * math.cos(0.23 * t)
        ox = min(max(ox, 2.0), 2.0 * margin - 2.0)
        oy = min(max(oy, 2.0), 2.0 * margin - 2.0)
        if t < start or (indices is not None and t not in count):
            continue
        ix, iy = int(math.floor(ox)), int(math.floor(oy))
        fx, fy = ox - ix, oy - iy
        c = canvas[:, iy:iy + height + 1, ix:ix + width + 1]
        f = ((1 - fy) * (1 - fx)) * c[:, :-1, :-1] + ((1 - fy) * fx) * c[:, :-1, 1:] \
            + (fy * (1 - fx)) * c[:, 1:, :-1] + (fy * fx) * c[:, 1:, 1:]
        if integer:
            f = f.round().clamp_(0, 255)
        f = f.to(dtype)
        frames.extend([f] * (count[t] if indices is not None else 1))
    return torch.stack(frames, 0).contiguous()


def synthetic_flows(num, height=376, width=1232, seed=3):
    """Smooth random flow fields [num,2,H,W] with KITTI-like magnitudes (ATDNVO parity input)."""
    g = torch.Generator().manual_seed(seed)
    f = torch.randn(num, 2, height, width, generator=g) * torch.tensor([20.0, 6.0]).view(1, 2, 1, 1)
    return (torch.nn.functional.avg_pool2d(f, 9, 1, 4) * 6).contiguous()


def noise_frames(num_frames, height=376, width=1232, seed=FRAME_SEED):
    """Pure randint(0,256) frames: the no-structure worst case."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (num_frames, 3, height, width), generator=g).float()


def keyframe_db(num_keyframes, dim=15360, seed=7, duplicates=((3, 11),)):
    """Synthetic keyframe embeddings with planted exact duplicates (first-index tie-break test)."""
    g = torch.Generator().manual_seed(seed)
    db = torch.randn(num_keyframes, dim, generator=g)
    for a, b in duplicates:
        if a < num_keyframes and b < num_keyframes:
            db[b] = db[a]
    return db
