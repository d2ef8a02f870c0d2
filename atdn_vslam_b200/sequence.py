"""Sequence-level odometry: batches of consecutive frame pairs through the flow net and the CLVO
encoder, sharding of ONE sequence across the GPUs of one box, exchange of the per-pair 512-d CLVO
features, then the serial LSTM scan and host pose chain.

Why features and not poses are gathered: ``ATDNVO`` is stateful (``odometry/network.py:95-104,137-140``),
so the relative pose of pair t depends on all earlier flows; only flow + CNN encoder are
pair-parallel.  Gathering [P,512] features (2 KB per pair) and scanning serially reproduces the
reference exactly; gathering poses would reset the LSTM at every shard boundary.

Two sharding schemes:
  * contiguous (``run``): rank r owns one contiguous pair range; one all-gather at the end, every rank scans.  Simple,
    but the serial scan (~7 us per pair) trails the step and is repeated on every rank.
  * batch-interleaved (``run_interleaved``): the sequence is cut into rounds of R consecutive batches, rank r computes
    batch r of every round, the features of a round are all-gathered asynchronously (NCCL stream), and RANK 0 ALONE runs
    the scan, one round behind, between its own batches; it is given a slightly smaller batch per round (``lead_pairs``)
    so that flow + scan on rank 0 take as long as flow alone on the others.  The scan is neither redundant nor exposed;
    one broadcast of the [P,6] relative poses ends the step.
"""
from __future__ import annotations

import math
import os

import torch
import torch.distributed as dist

from .poses import PoseChain

SLAM_SIZE = (376, 1232)   # neural_slam.py:54,198: every frame is resized to this before the flow net

# measured on B200 (profiles/r02o_scan_bench.txt, r02n_bench_8gpu.json): one scan call costs ~0.1 ms + ~14-15 us per pair
# (two grid barriers per LSTM step); in the sharded run rank 0 spent 82.6 ms per 4540 pairs in 11 calls.  The pair-parallel
# part costs ~1.23 ms per pair (1.32 before the mixed-precision attention probabilities; with the old constant rank 0 was the
# critical rank by 14 ms of a 720 ms step, profiles/r02al_bench_8gpu.json).  Rounds are scanned in groups of at least SCAN_MIN_PAIRS pairs.
SCAN_US_PER_CALL, SCAN_US_PER_PAIR, FLOW_US_PER_PAIR, SCAN_MIN_PAIRS = 500.0, 17.0, 1230.0, 400


def shard_ranges(num_pairs, world):
    """Contiguous pair ranges [start, end) per rank; the first ``num_pairs % world`` ranks get one extra."""
    base, extra = divmod(num_pairs, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < extra else 0)
        out.append((s, e))
        s = e
    return out


_FIRST_BATCH_DIV = int(os.environ.get("ATDN_FIRST_BATCH_DIV", "4"))     # A/B: size of the short first batch of a streamed run


def batch_ranges(num_pairs, batch_pairs, short_first=False):
    """Pair ranges [start, end) of the batches of one rank.  ``short_first``: the first batch of a streamed (host
    frames) run cannot overlap its own host->device copy, so it is a quarter batch -- the start-up bubble is the copy of
    ~batch_pairs/4 frames instead of a whole batch."""
    first = max(1, batch_pairs // _FIRST_BATCH_DIV) if (short_first and num_pairs > batch_pairs) else batch_pairs
    out, s = [], 0
    while s < num_pairs:
        e = min(num_pairs, s + (first if s == 0 else batch_pairs))
        out.append((s, e))
        s = e
    return out


def default_lead_pairs(world, batch_pairs):
    """Batch size of rank 0 (the sequencer) in a full round: ``batch_pairs`` minus the pairs whose flow time equals the
    scan of one round (world * batch_pairs pairs)."""
    if world == 1:
        return batch_pairs
    per_round = world * batch_pairs
    calls_per_round = min(1.0, per_round / SCAN_MIN_PAIRS)
    scan_us = SCAN_US_PER_CALL * calls_per_round + SCAN_US_PER_PAIR * per_round
    return max(1, batch_pairs - math.ceil(scan_us / FLOW_US_PER_PAIR))


def interleaved_rounds(num_pairs, world, batch_pairs, lead_pairs=None):
    """Batch-interleaved sharding of ONE sequence of ``num_pairs`` pairs: ``rounds[j][r] = (start, end)`` is the pair
    range of rank r in round j (empty when start == end); the ranges of a round are consecutive in rank order and the
    rounds are consecutive, so concatenating them in (round, rank) order restores the sequence.  Full rounds give
    ``lead_pairs`` pairs to rank 0 and ``batch_pairs`` to the others; the remainder is one last round split evenly
    (rank 0 in proportion), so no rank idles for a whole batch."""
    lead = default_lead_pairs(world, batch_pairs) if lead_pairs is None else max(1, min(batch_pairs, lead_pairs))
    sizes = [lead] + [batch_pairs] * (world - 1)
    per_round = sum(sizes)
    rounds, s = [], 0
    while num_pairs - s >= per_round:
        row = []
        for n in sizes:
            row.append((s, s + n))
            s += n
        rounds.append(row)
    rem = num_pairs - s
    if rem > 0:
        # even split in proportion to the full-round sizes (largest-remainder rounding)
        want = [rem * n / per_round for n in sizes]
        got = [int(math.floor(w)) for w in want]
        order = sorted(range(world), key=lambda r: (want[r] - got[r]), reverse=True)
        for r in order[: rem - sum(got)]:
            got[r] += 1
        row = []
        for n in got:
            row.append((s, s + n))
            s += n
        rounds.append(row)
    return rounds


def local_frame_ranges(rounds, rank):
    """Frame ranges (first, last) INCLUSIVE that ``rank`` needs, one per non-empty round (pair range + one halo frame)."""
    return [(row[rank][0], row[rank][1]) for row in rounds if row[rank][1] > row[rank][0]]


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
    neural_slam.py:197-199); the 376x1232 ``InputPadder`` is a no-op.  frames: [T,3,H,W] 0..255, float or uint8 -> float32.
    CUDA frames go through ``atdn_resize_aa`` (one kernel instead of ATen's generic one: 0.2 vs 1.4 ms per 55 frames);
    host tensors (fixtures, the CPU test suite) through ``torch.nn.functional.interpolate``, the reference's own op."""
    if tuple(frames.shape[-2:]) == tuple(size):
        return frames.float()
    if frames.is_cuda:
        from . import ops
        return ops.resize_aa(frames, size)
    return torch.nn.functional.interpolate(frames.float(), size=size, mode="bilinear", antialias=True, align_corners=False)


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
        self.timing = None        # set to {} to collect CUDA-event timings of the exchange / scan steps (bench.py)

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

    # -- pair-parallel part ---------------------------------------------------------------------------------
    def _feature_batches(self, frames, ranges):
        """Generator over ``ranges`` = [(first frame, last frame)] (inclusive: last - first pairs each) of ``frames``
        [T,3,H,W] -> CLVO features [pairs,512] per range.  Device frames are consumed in place; HOST frames (pinned
        memory, float or uint8) are streamed: the host->device copy of range k+1 runs on a copy stream while range k
        computes, through two staging buffers."""
        dev = next(self.flow_net.parameters()).device
        if frames.device == dev:          # resident frames
            for s, e in ranges:
                yield self._batch(preprocess(frames[s:e + 1]))
            return
        if not ranges:
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        nb = max(e - s + 1 for s, e in ranges)
        key = (nb,) + tuple(frames.shape[1:]) + (frames.dtype,)
        if self._staging_key is None or self._staging_key[1:] != key[1:] or self._staging_key[0] < nb:
            self._staging = [torch.empty((nb,) + tuple(frames.shape[1:]), dtype=frames.dtype, device=dev) for _ in range(2)]
            self._staging_key = key
        main = torch.cuda.current_stream(dev)
        ready = [torch.cuda.Event() for _ in ranges]
        consumed = [torch.cuda.Event() for _ in ranges]

        def enqueue_copy(k):
            s, e = ranges[k]
            with torch.cuda.stream(self._copy_stream):
                if k >= 2:
                    self._copy_stream.wait_event(consumed[k - 2])       # the staging buffer is free again
                else:
                    self._copy_stream.wait_stream(main)                  # ... or was last used by an earlier call
                self._staging[k & 1][: e - s + 1].copy_(frames[s:e + 1], non_blocking=True)
                ready[k].record(self._copy_stream)

        enqueue_copy(0)
        for k, (s, e) in enumerate(ranges):
            if k + 1 < len(ranges):
                enqueue_copy(k + 1)
            main.wait_event(ready[k])
            feats = self._batch(preprocess(self._staging[k & 1][: e - s + 1]))
            consumed[k].record(main)
            yield feats

    @torch.no_grad()
    def pair_features(self, frames):
        """frames [T,3,H,W] (device, or pinned host memory: streamed) -> CLVO features [T-1,512]."""
        t = frames.shape[0]
        dev = next(self.flow_net.parameters()).device
        if t < 2:
            return torch.empty(0, 512, device=dev)
        ranges = batch_ranges(t - 1, self.batch_pairs, short_first=frames.device != dev)
        return torch.cat(list(self._feature_batches(frames, ranges)), 0)

    # -- whole sequence ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def run(self, local_frames, num_pairs=None, group=None, chain=True):
        """Contiguous sharding.  ``local_frames``: this rank's frames (its pair range plus the one-frame halo), on the
        device or in (pinned) host memory -- host frames are streamed batch by batch under the compute.  Returns
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

    @torch.no_grad()
    def run_interleaved(self, local_frames, rounds, group=None, chain=True):
        """Batch-interleaved sharding of one sequence (module docstring).  ``rounds`` = ``interleaved_rounds(...)``
        (identical on every rank); ``local_frames`` = the concatenation, over this rank's non-empty rounds, of frames
        first..last of ``local_frame_ranges(rounds, rank)`` (each range carries its own halo frame), on the device or in
        pinned host memory -- or simply the whole sequence ([P+1] frames).  Returns the same tuple as ``run`` on every rank."""
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank(group) if world > 1 else 0
        dev = next(self.flow_net.parameters()).device
        num_pairs = rounds[-1][-1][1] if rounds else 0
        mine = [row[rank] for row in rounds]
        local, off = [], 0
        for s, e in mine:
            if e > s:
                local.append((off, off + (e - s)))
                off += e - s + 1
        if local_frames.shape[0] == num_pairs + 1 and off != num_pairs + 1:
            local = [(s, e) for s, e in mine if e > s]       # the WHOLE sequence was passed: index it globally
        elif off != local_frames.shape[0]:
            raise RuntimeError(f"rank {rank}: {local_frames.shape[0]} local frames, the rounds need {off}")
        batches = self._feature_batches(local_frames, local)
        width = max((e - s for row in rounds for s, e in row), default=1)
        timing = self.timing
        works, rots, trs = [], [], []

        def ev():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e

        scanned = [0]                                         # rounds [0, scanned) have gone through the LSTM

        def scan_upto(j_end, force=False):
            """Scan rounds [scanned, j_end) in ONE call once they hold >= SCAN_MIN_PAIRS pairs (or at the end)."""
            j0 = scanned[0]
            if j_end <= j0 or (not force and rounds[j_end - 1][-1][1] - rounds[j0][0][0] < SCAN_MIN_PAIRS):
                return
            t0 = ev() if timing is not None else None
            rows = []
            for j in range(j0, j_end):
                work, out = works[j]
                if work is not None:
                    work.wait()                               # the compute stream waits for the NCCL stream
                rows += [out[r, : e - s] for r, (s, e) in enumerate(rounds[j]) if e > s]
            t1 = ev() if timing is not None else None
            r_, t_ = self.odometry_net.recurrent_scan(torch.cat(rows, 0))
            rots.append(r_)
            trs.append(t_)
            scanned[0] = j_end
            if timing is not None:
                timing.setdefault("events", []).append((t0, t1, ev()))

        for j, row in enumerate(rounds):
            s, e = row[rank]
            send = torch.zeros(width, 512, dtype=torch.float32, device=dev)
            if e > s:
                send[: e - s] = next(batches)
            if world > 1:
                out = torch.empty(world, width, 512, dtype=torch.float32, device=dev)
                work = dist.all_gather_into_tensor(out.view(world * width, 512), send, group=group, async_op=True)
            else:
                out, work = send.view(1, width, 512), None
            works.append((work, out))
            if rank == 0:
                scan_upto(j)                                  # one round behind: never waits on a slower rank's batch
        if rank == 0:
            scan_upto(len(rounds), force=True)
            rt = torch.cat([torch.cat(rots, 0), torch.cat(trs, 0)], 1) if rots else torch.empty(0, 6, device=dev)
        else:
            for work, _ in works:
                work.wait()
            rt = torch.empty(num_pairs, 6, dtype=torch.float32, device=dev)
        if world > 1:
            dist.broadcast(rt, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        rot, tr = rt[:, :3].contiguous(), rt[:, 3:].contiguous()
        if not chain:
            return rot, tr, None, None
        poses, keys = PoseChain().extend(rot, tr)
        return rot, tr, poses, keys
