"""Sequence-level odometry: batches of consecutive frame pairs through the flow net and the CLVO
encoder, pair-range sharding across the GPUs of one box, one all-gather of the per-pair 512-d CLVO
features, then the serial LSTM scan and host pose chain.

Why features and not poses are gathered: ``ATDNVO`` is stateful (``odometry/network.py:95-104,137-140``),
so the relative pose of pair t depends on all earlier flows; only flow + CNN encoder are
pair-parallel.  Gathering [P,512] features (2 KB per pair) and scanning serially reproduces the
reference exactly; gathering poses would reset the LSTM at every shard boundary.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .poses import PoseChain

SLAM_SIZE = (376, 1232)   # neural_slam.py:54,198: every frame is resized to this before the flow net


def shard_ranges(num_pairs, world):
    """Contiguous pair ranges [start, end) per rank; the first ``num_pairs % world`` ranks get one extra."""
    base, extra = divmod(num_pairs, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < extra else 0)
        out.append((s, e))
        s = e
    return out


def batch_ranges(num_pairs, batch_pairs, short_first=False):
    """Pair ranges [start, end) of the batches of one rank.  ``short_first``: the first batch of a streamed (host
    frames) run cannot overlap its own host->device copy, so it is a quarter batch -- the start-up bubble is the copy of
    ~batch_pairs/4 frames instead of a whole batch."""
    first = max(1, batch_pairs // 4) if (short_first and num_pairs > batch_pairs) else batch_pairs
    out, s = [], 0
    while s < num_pairs:
        e = min(num_pairs, s + (first if s == 0 else batch_pairs))
        out.append((s, e))
        s = e
    return out


def gather_features(local, num_pairs, group=None):
    """All-gather variable-length [p_r, D] shards (contiguous ranges of ``shard_ranges``) into [P, D]."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    ranges = shard_ranges(num_pairs, world)
    width = max(e - s for s, e in ranges)
    d = local.shape[1]
    padded = torch.zeros(width, d, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty(world * width, d, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * width: r * width + (e - s)] for r, (s, e) in enumerate(ranges)], 0)


def preprocess(frames, size=SLAM_SIZE):
    """Caller-side resize of ``NeuralSLAM.__call__`` (``TF.resize`` = antialiased bilinear,
    neural_slam.py:197-199); the 376x1232 ``InputPadder`` is a no-op.  frames: [T,3,H,W] float 0..255."""
    if tuple(frames.shape[-2:]) == tuple(size):
        return frames
    return torch.nn.functional.interpolate(frames, size=size, mode="bilinear", antialias=True, align_corners=False)


class OdometryPipeline:
    """flow_net: ``RAFTGMA``; odometry_net: ``ATDNVO`` (both on the same CUDA device).

    ``use_graphs``: the ~330 kernel launches of one batch (encoders, corr pyramid, attention, 12 update
    iterations, upsampling, CLVO encoder) are captured once per batch size into a CUDA graph and
    replayed, so the 12-iteration loop is not launch-bound (SURVEY.md hard part H4)."""

    def __init__(self, flow_net, odometry_net, batch_pairs=6, iters=12, use_graphs=True):
        self.flow_net, self.odometry_net = flow_net, odometry_net
        self.batch_pairs, self.iters, self.use_graphs = batch_pairs, iters, use_graphs
        self._graphs, self._graph_gen = {}, None
        self._copy_stream, self._staging, self._staging_key = None, None, None

    def _batch_eager(self, frames):
        _, flow_up = self.flow_net.forward_frames(frames, iters=self.iters, test_mode=True)
        return self.odometry_net.encode(flow_up)

    def _batch(self, frames):
        """frames [b+1,3,H,W] -> CLVO features [b,512]"""
        if not self.use_graphs:
            return self._batch_eager(frames)
        # captured graphs hold raw pointers to the packed weights and plan buffers: the nets' generation counters change
        # whenever those are dropped (load_state_dict, .to(), ...), which retires every graph captured before
        gen = (getattr(self.flow_net, "generation", 0), getattr(self.odometry_net, "generation", 0))
        if gen != self._graph_gen:
            self._graphs, self._graph_gen = {}, gen
        key = tuple(frames.shape) + (str(frames.device), self.iters)
        g = self._graphs.get(key)
        if g is None:
            static_in = frames.clone()
            self._batch_eager(static_in)                 # warm-up: packs weights, sizes buffers, sets attributes
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._batch_eager(static_in)
            g = self._graphs[key] = (graph, static_in, static_out)
        graph, static_in, static_out = g
        static_in.copy_(frames)
        graph.replay()
        return static_out.clone()

    @torch.no_grad()
    def pair_features(self, frames):
        """frames [T,3,H,W] -> CLVO features [T-1,512] (pair-parallel part).  Device frames are consumed in
        place; HOST frames (pinned memory) are streamed: the host->device copy of batch k+1 runs on a copy
        stream while batch k computes, through two staging buffers."""
        if not frames.is_cuda:
            return self._pair_features_streamed(frames)
        frames = preprocess(frames)
        t = frames.shape[0]
        feats = [self._batch(frames[s:e + 1]) for s, e in batch_ranges(t - 1, self.batch_pairs)]
        return torch.cat(feats, 0) if feats else torch.empty(0, 512, device=frames.device)

    def _pair_features_streamed(self, host_frames):
        dev = next(self.flow_net.parameters()).device
        t = host_frames.shape[0]
        if t < 2:
            return torch.empty(0, 512, device=dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        nb = self.batch_pairs + 1
        key = (nb,) + tuple(host_frames.shape[1:]) + (host_frames.dtype,)
        if self._staging_key != key:
            self._staging = [torch.empty((nb,) + tuple(host_frames.shape[1:]), dtype=host_frames.dtype, device=dev) for _ in range(2)]
            self._staging_key = key
        main = torch.cuda.current_stream(dev)
        ranges = batch_ranges(t - 1, self.batch_pairs, short_first=True)
        ready = [torch.cuda.Event() for _ in ranges]
        consumed = [torch.cuda.Event() for _ in ranges]

        def enqueue_copy(k):
            s, e = ranges[k]
            with torch.cuda.stream(self._copy_stream):
                if k >= 2:
                    self._copy_stream.wait_event(consumed[k - 2])       # the staging buffer is free again
                else:
                    self._copy_stream.wait_stream(main)                  # ... or was last used by an earlier call
                self._staging[k & 1][: e - s + 1].copy_(host_frames[s:e + 1], non_blocking=True)
                ready[k].record(self._copy_stream)

        feats = []
        enqueue_copy(0)
        for k, (s, e) in enumerate(ranges):
            if k + 1 < len(ranges):
                enqueue_copy(k + 1)
            main.wait_event(ready[k])
            feats.append(self._batch(preprocess(self._staging[k & 1][: e - s + 1].float())))
            consumed[k].record(main)
        return torch.cat(feats, 0)

    @torch.no_grad()
    def run(self, local_frames, num_pairs=None, group=None, chain=True):
        """``local_frames``: this rank's frames (its pair range plus the one-frame halo), on the device or in
        (pinned) host memory -- host frames are streamed batch by batch under the compute.  Returns
        (rot [P,3], tr [P,3], poses [P+1,4,4] or None, keyframe indices or None); the LSTM scan and
        pose chain run redundantly on every rank (they are microseconds per pair)."""
        feats = self.pair_features(local_frames)
        total = feats.shape[0] if num_pairs is None else num_pairs
        feats = gather_features(feats, total, group)
        rot, tr = self.odometry_net.recurrent_scan(feats)
        if not chain:
            return rot, tr, None, None
        poses, keys = PoseChain().extend(rot, tr)    # ONE device->host copy for the whole sequence
        return rot, tr, poses, keys
