"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference CLVO pose network (ATDNVO),
the pose-assembly helpers, the MappingVAE keyframe encoder and the keyframe search.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  Pinned against the reference itself by ``tests/golden/make_golden.py``
(see ``gma_oracle.py`` header).  Citations are relative to ``/root/reference/atdn_vslam/``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

FLOW_STD = (58.1837, 17.7647)          # utils/normalizations.py:8-10
RGB_MEAN = (0.485, 0.456, 0.406)       # utils/normalizations.py:4-6
RGB_STD = (0.229, 0.224, 0.225)


def _bn(x, sd, name):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], training=False, eps=1e-5)


def conv_block(x, sd, p, stride=1, padding=0):
    """layers/conv.py:36-37: bn(mish(conv(x))) -- batch norm AFTER the activation."""
    y = F.conv2d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], stride=stride, padding=padding)
    return _bn(F.mish(y), sd, p + "bn")


def residual_conv(x, sd, p, stride):
    """layers/conv.py:83-90"""
    y = conv_block(x, sd, p + "conv.0.", 1, 1)
    y = conv_block(y, sd, p + "conv.1.", stride, 1)
    skip = F.conv2d(x, sd[p + "skip_layer.weight"], sd[p + "skip_layer.bias"], stride=stride)
    return _bn(F.mish(y + skip), sd, p + "out_block.1")


def linear_block(x, sd, p):
    """layers/linear.py:35-42 with activation=Mish, norm=False, dropout=False."""
    return F.mish(F.linear(x, sd[p + "linear.weight"], sd[p + "linear.bias"]))


def atdnvo_encode(sd, flows):
    """odometry/network.py:131-134, 63-73: flow / std -> CNN -> [B, 512]."""
    std = torch.tensor(FLOW_STD, dtype=flows.dtype, device=flows.device).view(1, 2, 1, 1)
    x = flows / std
    x = F.conv2d(x, sd["encoder_CNN.0.weight"], sd["encoder_CNN.0.bias"], groups=2)
    x = conv_block(x, sd, "encoder_CNN.1.", 2, 3)
    for i in range(2, 6):
        x = residual_conv(x, sd, f"encoder_CNN.{i}.", 2)
    x = conv_block(x, sd, "encoder_CNN.6.", 3, 0)
    x = x.flatten(1)
    return linear_block(x, sd, "encoder_CNN.8.")


def lstm_cell(x, h, c, sd, p):
    """torch.nn.LSTMCell semantics, gate order (i, f, g, o)."""
    gates = F.linear(x, sd[p + ".weight_ih"], sd[p + ".bias_ih"]) + F.linear(h, sd[p + ".weight_hh"], sd[p + ".bias_hh"])
    i, f, g, o = gates.chunk(4, dim=1)
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h = torch.sigmoid(o) * torch.tanh(c)
    return h, c


def zero_state(batch=1, device=None):
    z = lambda: torch.zeros(batch, 512, device=device)
    return [z(), z(), z(), z()]          # h1, c1, h2, c2 (odometry/network.py:95-104)


def atdnvo_recurrent(sd, feat, state):
    """odometry/network.py:137-144: one LSTM step + both heads.  ``state`` is updated in place."""
    h1, c1, h2, c2 = state
    h1, c1 = lstm_cell(feat, h1, c1, sd, "lstm1")
    x = linear_block(h1, sd, "lstm_linear.")
    h2, c2 = lstm_cell(x, h2, c2, sd, "lstm2")
    state[:] = [h1, c1, h2, c2]

    def head(p):
        y = linear_block(h2, sd, p + ".0.")
        y = linear_block(y, sd, p + ".1.")
        return F.linear(y, sd[p + ".2.weight"])

    return head("rotation_regressor"), head("translation_regressor")


def atdnvo_forward(sd, flows, state):
    """Reference ``ATDNVO.forward`` (stateful across calls through ``state``)."""
    with torch.no_grad():
        return atdnvo_recurrent(sd, atdnvo_encode(sd, flows.float()), state)


# ----------------------------------------------------------------------------------------------
# pose assembly -- utils/transforms.py:25-51, 54-94, 97-119; slam_framework/neural_slam.py:288-302
# ----------------------------------------------------------------------------------------------
def euler2matrix(r):
    """'yxz' convention, utils/transforms.py:68-81: fp32 torch trig, products formed in fp32."""
    r = r.detach().float()
    c1, c2, c3 = torch.cos(r[0]), torch.cos(r[1]), torch.cos(r[2])
    s1, s2, s3 = torch.sin(r[0]), torch.sin(r[1]), torch.sin(r[2])
    return torch.tensor([[c1 * c3 + s1 * s2 * s3, c3 * s1 * s2 - c1 * s3, c2 * s1],
                         [c2 * s3, c2 * c3, -s2],
                         [c1 * s2 * s3 - c3 * s1, c1 * c3 * s2 + s1 * s3, c1 * c2]], dtype=torch.float32)


def matrix2euler(R):
    """'yxz' convention, utils/transforms.py:41-44."""
    a = torch.atan2(R[0, 2], R[2, 2])
    b = torch.atan2(-R[1, 2], torch.sqrt(1 - R[1, 2] ** 2))
    g = torch.atan2(R[1, 0], R[1, 1])
    return torch.tensor([a, b, g])


def transform(rot, tr):
    """utils/transforms.py:97-119 -> [4,4] fp32."""
    m = torch.eye(4, dtype=torch.float32)
    m[:3, :3] = euler2matrix(rot)
    m[:3, 3] = tr.float()
    return m


def chain_and_keyframes(rots, trs, rot_thresh_deg=10.0, tr_thresh=15.0):
    """neural_slam.py:204-215, 288-302: pose chaining + keyframe decisions over a sequence.
    Returns (poses [T+1,4,4], keyframe frame indices; frame 0 is always a keyframe :224-225)."""
    pose = torch.eye(4, dtype=torch.float32)
    prop = torch.eye(4, dtype=torch.float32)
    poses, keys = [pose.clone()], [0]
    rt = rot_thresh_deg / 180 * math.pi
    for t in range(len(rots)):
        m = transform(rots[t], trs[t])
        pose = pose @ m
        prop = prop @ m
        if torch.norm(matrix2euler(prop[:3, :3])) > rt or torch.norm(prop[:3, 3]) > tr_thresh:
            keys.append(t + 1)
            prop = torch.eye(4, dtype=torch.float32)
        poses.append(pose.clone())
    return torch.stack(poses), keys


# ----------------------------------------------------------------------------------------------
# localization -- localization/network.py:29-45, 57-72; neural_slam.py:373-384
# ----------------------------------------------------------------------------------------------
def vae_embed(sd, image):
    """MappingVAE encoder + mean_lin -> mu [B,128,H/64,W/64] (non-variational path)."""
    with torch.no_grad():
        mean = torch.tensor(RGB_MEAN).view(1, 3, 1, 1)
        std = torch.tensor(RGB_STD).view(1, 3, 1, 1)
        x = (image.float() / 255.0 - mean) / std
        x = conv_block(x, sd, "encoder.0.", 1, 3)
        for i in range(1, 7):
            x = residual_conv(x, sd, f"encoder.{i}.", 2)
        return F.conv2d(x, sd["mean_lin.weight"], sd["mean_lin.bias"])


def keyframe_search(embeddings, code):
    """neural_slam.py:373-384: L2 distance of every keyframe embedding to ``code``; arg-min returns
    the FIRST minimum.  embeddings [K, D], code [D] -> (index, distances [K])."""
    d = torch.stack([torch.norm(embeddings[i] - code, p=2) for i in range(embeddings.shape[0])])
    return int(torch.argmin(d)), d
