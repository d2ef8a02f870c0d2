"""Parameter schemas of the three reference networks on the hot path.

The drop-in classes consume the reference's state dicts *unchanged*, so the tensor
names and shapes below are the interface, not an implementation:

- GMA flow net  : 185 tensors (``GMA.whl!/GMA/core/network.py:40-43`` builds fnet, cnet,
  update_block, att); checkpoints carry a ``module.`` prefix from ``DataParallel``
  (``atdn_vslam/slam_framework/neural_slam.py:51-52``).
- ATDNVO (CLVO) : 127 tensors (``atdn_vslam/odometry/network.py:42-119``).
- MappingVAE    : encoder + ``mean_lin`` only (``atdn_vslam/localization/network.py:29-45``);
  decoder tensors are accepted and ignored.

Each schema is an ordered ``{name: (shape, kind)}`` with kind in
{"conv_w", "lin_w", "bias", "bn_w", "bn_b", "bn_mean", "bn_var", "bn_count", "gamma",
"emb", "index"} which the seeded weight factory (``synth.py``) uses to pick a distribution.
"""
from __future__ import annotations

from collections import OrderedDict

Schema = "OrderedDict[str, tuple[tuple[int, ...], str]]"


def _conv(s, name, cout, cin, kh, kw, bias=True):
    s[name + ".weight"] = ((cout, cin, kh, kw), "conv_w")
    if bias:
        s[name + ".bias"] = ((cout,), "bias")


def _bn(s, name, c):
    s[name + ".weight"] = ((c,), "bn_w")
    s[name + ".bias"] = ((c,), "bn_b")
    s[name + ".running_mean"] = ((c,), "bn_mean")
    s[name + ".running_var"] = ((c,), "bn_var")
    s[name + ".num_batches_tracked"] = ((), "bn_count")


def _lin(s, name, cout, cin, bias=True):
    s[name + ".weight"] = ((cout, cin), "lin_w")
    if bias:
        s[name + ".bias"] = ((cout,), "bias")


# ----------------------------------------------------------------------------------------------
# GMA (RAFTGMA) -- GMA.whl!/GMA/core/{network,extractor,update,gma}.py
# ----------------------------------------------------------------------------------------------
ENCODER_STAGES = ((64, 1), (96, 2), (128, 2))  # (planes, stride of first block) extractor.py:134-136


def _basic_encoder(s, p, out_dim, batch_norm):
    if batch_norm:
        _bn(s, p + "norm1", 64)
    _conv(s, p + "conv1", 64, 3, 7, 7)
    cin = 64
    for li, (planes, stride) in enumerate(ENCODER_STAGES, start=1):
        for bi in range(2):
            q = f"{p}layer{li}.{bi}."
            st = stride if bi == 0 else 1
            _conv(s, q + "conv1", planes, cin, 3, 3)
            _conv(s, q + "conv2", planes, planes, 3, 3)
            if batch_norm:
                _bn(s, q + "norm1", planes)
                _bn(s, q + "norm2", planes)
                if st != 1:
                    _bn(s, q + "norm3", planes)
            if st != 1:
                _conv(s, q + "downsample.0", planes, cin, 1, 1)
                if batch_norm:
                    _bn(s, q + "downsample.1", planes)
            cin = planes
    _conv(s, p + "conv2", out_dim, 128, 1, 1)


def gma_schema():
    s = OrderedDict()
    _basic_encoder(s, "fnet.", 256, batch_norm=False)   # instance norm: no tensors
    _basic_encoder(s, "cnet.", 256, batch_norm=True)
    u = "update_block."
    _conv(s, u + "encoder.convc1", 256, 324, 1, 1)
    _conv(s, u + "encoder.convc2", 192, 256, 3, 3)
    _conv(s, u + "encoder.convf1", 128, 2, 7, 7)
    _conv(s, u + "encoder.convf2", 64, 128, 3, 3)
    _conv(s, u + "encoder.conv", 126, 256, 3, 3)
    for n, (kh, kw) in (("1", (1, 5)), ("2", (5, 1))):
        for g in "zrq":
            _conv(s, f"{u}gru.conv{g}{n}", 128, 512, kh, kw)
    _conv(s, u + "flow_head.conv1", 256, 128, 3, 3)
    _conv(s, u + "flow_head.conv2", 2, 256, 3, 3)
    _conv(s, u + "mask.0", 256, 128, 3, 3)
    _conv(s, u + "mask.2", 576, 256, 1, 1)
    s[u + "aggregator.gamma"] = ((1,), "gamma")
    _conv(s, u + "aggregator.to_v", 128, 128, 1, 1, bias=False)
    _conv(s, "att.to_qk", 256, 128, 1, 1, bias=False)
    s["att.pos_emb.rel_ind"] = ((160, 160), "index")           # gma.py:16-18 (unused on the path)
    s["att.pos_emb.rel_height.weight"] = ((319, 128), "emb")
    s["att.pos_emb.rel_width.weight"] = ((319, 128), "emb")
    return s


# ----------------------------------------------------------------------------------------------
# ATDNVO -- atdn_vslam/odometry/network.py:42-119, layers/conv.py:8-90, layers/linear.py:5-42
# ----------------------------------------------------------------------------------------------
def _conv_block(s, p, cout, cin, k):
    _conv(s, p + "conv", cout, cin, k, k)
    _bn(s, p + "bn", cout)


def _residual_conv(s, p, cin, cout):
    _conv_block(s, p + "conv.0.", cin, cin, 3)
    _conv_block(s, p + "conv.1.", cout, cin, 3)
    _conv(s, p + "skip_layer", cout, cin, 1, 1)
    _bn(s, p + "out_block.1", cout)


def atdnvo_schema():
    s = OrderedDict()
    _bn(s, "polar_norm", 2)                                    # network.py:43 (unused in forward)
    s["encoder_CNN.0.weight"] = ((2, 1, 1, 1), "conv_w")       # depthwise 1x1, groups=2
    s["encoder_CNN.0.bias"] = ((2,), "bias")
    _conv_block(s, "encoder_CNN.1.", 16, 2, 7)
    for i in range(2, 6):
        _residual_conv(s, f"encoder_CNN.{i}.", 16, 16)
    _conv_block(s, "encoder_CNN.6.", 16, 16, 3)
    _lin(s, "encoder_CNN.8.linear", 512, 832)
    for n in ("lstm1", "lstm2"):
        s[n + ".weight_ih"] = ((2048, 512), "lin_w")
        s[n + ".weight_hh"] = ((2048, 512), "lin_w")
        s[n + ".bias_ih"] = ((2048,), "bias")
        s[n + ".bias_hh"] = ((2048,), "bias")
        if n == "lstm1":
            _lin(s, "lstm_linear.linear", 512, 512)
    for head in ("translation_regressor", "rotation_regressor"):
        _lin(s, head + ".0.linear", 128, 512)
        _lin(s, head + ".1.linear", 64, 128)
        _lin(s, head + ".2", 3, 64, bias=False)
    return s


# ----------------------------------------------------------------------------------------------
# MappingVAE encoder -- atdn_vslam/localization/network.py:29-45
# ----------------------------------------------------------------------------------------------
VAE_CHANNELS = (3, 16, 16, 32, 64, 128, 128)


def vae_encoder_schema():
    s = OrderedDict()
    _conv_block(s, "encoder.0.", 3, 3, 7)
    for i in range(1, 7):
        _residual_conv(s, f"encoder.{i}.", VAE_CHANNELS[i - 1], VAE_CHANNELS[i])
    _conv(s, "mean_lin", 128, 128, 1, 1)
    return s


def strip_module_prefix(sd):
    """Accept both plain and DataParallel-wrapped ('module.'-prefixed) state dicts."""
    if all(k.startswith("module.") for k in sd):
        return OrderedDict((k[len("module."):], v) for k, v in sd.items())
    return sd
