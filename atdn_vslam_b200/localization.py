"""Keyframe embedding and similarity search of the localization network.

``MappingEncoder`` is the encoder + ``mean_lin`` half of the reference ``MappingVAE``
(``atdn_vslam/localization/network.py:29-45, 57-72``): image -> ``mu [B,128,H/64,W/64]``.  It loads
the reference's ``MappingVAE`` state dict unchanged (decoder tensors are accepted and ignored: the
decoder only feeds the training loss).  ``forward`` returns the
reference's 4-tuple ``(mu, logvar, latent, decoded)`` with ``decoded=None``.

``KeyframeIndex`` replaces the Python loop of ``NeuralSLAM.__get_closest_keyframe``
(``atdn_vslam/slam_framework/neural_slam.py:373-384``) with one streaming L2 + first-arg-min kernel
over a contiguous ``[K, D]`` fp32 database; ``search_sharded`` is the multi-GPU variant
(row-sharded database, all-gather of R (distance, global index) pairs, lowest-index tie-break).
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import ops, schema
from .gma import _build_module_tree
from .odometry import _ConvBlock, _ResidualBlock, run_conv_block, run_residual_block

RGB_MEAN = (0.485, 0.456, 0.406)   # atdn_vslam/utils/normalizations.py:4-6
RGB_STD = (0.229, 0.224, 0.225)


class MappingEncoder(nn.Module):
    def __init__(self, variational=False):
        super().__init__()
        if variational:
            raise NotImplementedError("the SLAM uses the non-variational MappingVAE (localization/network.py:12)")
        _build_module_tree(self, schema.vae_encoder_schema())
        self._packed = {}          # kernel-layout weights per device
        super().train(False)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._packed = {}
        own = set(self.state_dict().keys())
        sd = {k: v for k, v in state_dict.items() if k in own or not k.startswith("decoder.")}
        return super().load_state_dict(sd, strict=strict, **kw)

    def _apply(self, fn, *a, **kw):
        self._packed = {}
        return super()._apply(fn, *a, **kw)

    def train(self, mode=True):
        """Inference only: the map training of ``NeuralSLAM.__create_map`` (neural_slam.py:305-352) needs the decoder and
        autograd of the reference ``MappingVAE``; train that, then ``load_state_dict`` its weights here."""
        if mode:
            raise NotImplementedError("MappingEncoder is the inference-only keyframe embedder (no decoder, no autograd): keep "
                                      "the reference MappingVAE for __create_map and load its state dict into this class")
        return super().train(False)

    def _weights(self, dev):
        key = str(dev)
        if key not in self._packed:
            sd = {k: v.detach().to(dev) for k, v in self.state_dict().items()}
            mean = torch.tensor(RGB_MEAN, dtype=torch.float32, device=dev)
            std = torch.tensor(RGB_STD, dtype=torch.float32, device=dev)
            self._packed[key] = {
                # Normalize(0,255) then Normalize(mean,std): x * 1/(255 std) - mean/std
                "in_scale": (1.0 / (255.0 * std)).contiguous(), "in_shift": (-mean / std).contiguous(),
                "stem": _ConvBlock(sd, "encoder.0."),
                "res": [_ResidualBlock(sd, f"encoder.{i}.") for i in range(1, 7)],
                "mean_w": sd["mean_lin.weight"].float().contiguous(), "mean_b": sd["mean_lin.bias"].float().contiguous(),
            }
        return self._packed[key]

    @torch.no_grad()
    def embed(self, image):
        """image fp32 [B,3,H,W] in 0..255 -> mu [B,128,H/64,W/64]."""
        L.require_cuda(image)
        p = self._weights(image.device)
        x = run_conv_block(image.float().contiguous(), p["stem"], 1, 3, in_scale=p["in_scale"], in_shift=p["in_shift"])
        for blk in p["res"]:
            x = run_residual_block(x, blk, 2)
        mu = torch.empty_like(x)
        ops.conv32(x, p["mean_w"], p["mean_b"], mu)
        return mu

    def forward(self, image):
        mu = self.embed(image)
        return mu, None, mu, None


class KeyframeIndex:
    """Contiguous device-resident embedding database with append + nearest-keyframe search."""

    def __init__(self, dim=15360, capacity=1024, device="cuda"):
        self.dim, self.device = dim, torch.device(device)
        self._db = torch.empty(capacity, dim, dtype=torch.float32, device=self.device)
        self._n = 0
        self._dist = torch.empty(capacity, dtype=torch.float32, device=self.device)
        self._idx = torch.empty(1, dtype=torch.int32, device=self.device)

    def __len__(self):
        return self._n

    @property
    def embeddings(self):
        return self._db[: self._n]

    def add(self, embedding):
        e = embedding.detach().reshape(-1, self.dim).to(self.device, torch.float32)
        need = self._n + e.shape[0]
        if need > self._db.shape[0]:
            cap = max(need, 2 * self._db.shape[0])
            db = torch.empty(cap, self.dim, dtype=torch.float32, device=self.device)
            db[: self._n] = self._db[: self._n]
            self._db, self._dist = db, torch.empty(cap, dtype=torch.float32, device=self.device)
        self._db[self._n:need] = e
        self._n = need

    def search_device(self, code):
        """Enqueue the search; returns device tensors (index int32 [1], distances fp32 [K]) without a sync."""
        if self._n == 0:
            raise RuntimeError("keyframe search on an empty database")
        q = code.detach().reshape(-1).to(self.device, torch.float32).contiguous()
        if q.numel() != self.dim:
            raise RuntimeError(f"query has {q.numel()} elements, database rows have {self.dim}")
        ops.keyframe_search(self._db[: self._n], q, self._dist, self._idx)
        return self._idx, self._dist[: self._n]

    def search(self, code):
        """-> (index of the closest keyframe (first minimum), distances [K]) like neural_slam.py:373-384."""
        idx, dist = self.search_device(code)
        i = int(idx.item())
        if not 0 <= i < self._n:
            raise RuntimeError(f"keyframe search returned index {i} for {self._n} keyframes")
        return i, dist.clone()

    def search_sharded(self, code, group=None):
        """Every rank holds a row shard; returns the GLOBAL arg-min (lowest global index on ties).
        Global index = sum of lower ranks' sizes + local index (contiguous row sharding)."""
        if self._n > 0:
            idx, d = self.search_device(code)
            local = (d[idx.long()].reshape(()), idx.long().reshape(()))
        else:
            local = None
        return global_first_minimum(local, self._n, self.device, group)


def global_first_minimum(local, n_local, device, group=None):
    """Exchange step of the sharded keyframe search: ``local`` = (distance, local index) 0-d tensors of
    this rank's first minimum (None for an empty shard), ``n_local`` = rows in this rank's shard.  Two small all-gathers
    (shard sizes, then one (distance, global index) pair per rank); every rank returns the same
    (global index, distance) with the lowest global index on ties, like the reference's serial loop."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n_local], dtype=torch.int64, device=device), group=group)
    offset = int(sum(int(s.item()) for s in sizes[:rank]))
    if local is not None and n_local > 0:
        best = torch.stack([local[0].double(), (local[1] + offset).double()]).to(device)
    else:
        best = torch.tensor([float("inf"), -1.0], dtype=torch.float64, device=device)
    allb = [torch.zeros(2, dtype=torch.float64, device=device) for _ in range(world)]
    dist.all_gather(allb, best, group=group)
    return merge_shard_minima([(float(t[0]), int(t[1])) for t in allb])


class Relocalizer:
    """Relocalisation + refinement (``neural_slam.py:355-370, 387-399``) on the B200 path, as one stream of
    kernels: embed the query (encoder only -- the reference also runs the unused VAE decoder), search the keyframe
    database, run flow + pose between the closest keyframe's image and the query, compose
    ``refined = initial @ transform(rot, tr)``.  Keyframe images live in device memory here (the reference
    ``torch.load``s ``rgb/%06d.pth`` per query); poses stay on the host like ``Frame.pose``.

    ``flow_net`` / ``odometry_net`` / ``mapping_net``: ``RAFTGMA``, ``ATDNVO``, ``MappingEncoder`` on one device.
    The odometry net's LSTM state advances exactly as in the reference, which calls the same stateful network."""

    def __init__(self, flow_net, odometry_net, mapping_net, iters=12):
        self.flow_net, self.odometry_net, self.mapping_net, self.iters = flow_net, odometry_net, mapping_net, iters
        self.index = None
        self.images, self.poses = [], []

    def add_keyframe(self, image, pose):
        """image fp32/uint8 [3,H,W] (0..255, already at the SLAM size), pose [4,4]; embeds and registers it."""
        dev = next(self.mapping_net.parameters()).device
        img = image.to(dev).float().unsqueeze(0).contiguous()
        mu = self.mapping_net.embed(img)
        if self.index is None:
            self.index = KeyframeIndex(dim=mu[0].numel(), device=dev)
        self.index.add(mu)
        self.images.append(img)
        self.poses.append(pose.detach().to("cpu", torch.float32).clone())
        return len(self.images) - 1

    @torch.no_grad()
    def relocalize(self, image):
        """image [1,3,H,W] or [3,H,W] -> (initial_pose [4,4], refined_pose [4,4], distances [K], keyframe index)."""
        from .poses import transform
        if self.index is None or len(self.index) == 0:
            raise RuntimeError("relocalisation without keyframes")
        dev = self.index.device
        img = image.to(dev).float()
        img = img.unsqueeze(0) if img.dim() == 3 else img
        mu = self.mapping_net.embed(img.contiguous())
        k, distances = self.index.search(mu)
        initial = self.poses[k]
        _, flow = self.flow_net(self.images[k], img, iters=self.iters, test_mode=True)      # :395
        rot, tr = self.odometry_net(flow)                                                    # :396
        pose_diff = transform(rot.squeeze(), tr.squeeze())                                   # :397
        return initial, initial @ pose_diff, distances, k


def merge_shard_minima(pairs):
    """[(distance, global_index)] per shard -> (global_index, distance) of the first global minimum."""
    best = None
    for d, i in pairs:
        if i < 0:
            continue
        if best is None or d < best[0] or (d == best[0] and i < best[1]):
            best = (d, i)
    if best is None:
        raise RuntimeError("keyframe search on an empty database")
    return best[1], best[0]
