#!/usr/bin/env python
"""Benchmark of the hot path: KITTI-shaped frame pairs/s for GMA flow (iters=12) + CLVO pose.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                      (the reference's CPU path, see below)

One step = one pass of the hot path over ONE synthetic sequence: at N=1 the 271-frame sequence (270 pairs;
BASELINE.json configs[1]); at N>1 the 4541-frame sequence (4540 pairs; configs[3]) sharded batch-interleaved
across the ranks (sequence.run_interleaved): every rank computes its batches of consecutive pairs, the [pairs,512]
CLVO features of a round are all-gathered asynchronously (NCCL), rank 0 runs the serial LSTM scan one round behind
between its own (slightly smaller) batches and broadcasts the relative poses.  `value` times the step
with frames resident in HBM; `e2e` times the public API with frames in pinned HOST memory (H2D of
the frames and D2H of the relative poses inside the timed region, plus the host pose chain).
`online_b1` (extra key, rank 0) is the wall-clock latency of the reference's per-frame loop shape
(one pair per call through the drop-in forward()s, host read after every frame, neural_slam.py:202-204).

`--impl reference`: /root/reference does not exist on the GPU box and the reference is pure Python,
so this arm times the oracle port of the reference's device=cpu path (oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frame_pairs_per_s"
UNIT = "pairs/s"
FRAMES = 271          # BASELINE.json configs[1] (N=1)
SEQ_FRAMES = 4541     # BASELINE.json configs[3] (N>1): one sequence sharded across the ranks
H_RAW, W_RAW = 376, 1241


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's device=cpu path
# ----------------------------------------------------------------------------------------------
def cpu_pairs_per_s(num_pairs, warm=1):
    from atdn_vslam_b200 import synth
    from atdn_vslam_b200.sequence import preprocess
    from oracle import gma_oracle, clvo_oracle
    torch.set_num_threads(os.cpu_count() or 1)
    gsd, vsd = synth.gma_state_dict(), synth.atdnvo_state_dict()
    frames = preprocess(synth.frame_sequence(num_pairs + warm + 1, H_RAW, W_RAW))
    state = clvo_oracle.zero_state()

    def pair(t):
        _, up = gma_oracle.raftgma_forward(gsd, frames[t:t + 1], frames[t + 1:t + 2], iters=12, aten_ops=True)
        return clvo_oracle.atdnvo_forward(vsd, up, state)

    for t in range(warm):
        pair(t)
    t0 = time.perf_counter()
    for t in range(warm, warm + num_pairs):
        pair(t)
    dt = time.perf_counter() - t0
    return num_pairs / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 2
    times = []
    for _ in range(args.warmup):
        cpu_pairs_per_s(1, warm=0)
    for _ in range(args.steps):
        v, dt = cpu_pairs_per_s(sample, warm=0)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms / 1e3)
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "seq271_376x1241_gma12_clvo", "sample": f"{sample} consecutive pairs per step",
                       "note": "oracle port of the reference device=cpu path (reference is pure Python; /root/reference is absent on the GPU box)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{sample} pairs/step x {args.steps} steps"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# ours
# ----------------------------------------------------------------------------------------------
class KernelProfile:
    """Per-label CUDA-event timing of C-ABI launches on the launching stream (installed as L.PROFILER)."""

    def __init__(self):
        self.events = []

    @contextlib.contextmanager
    def __call__(self, label, flops, nbytes):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        e.record()
        self.events.append((label, flops, nbytes, s, e))

    def table(self):
        torch.cuda.synchronize()
        agg = {}
        for label, flops, nbytes, s, e in self.events:
            a = agg.setdefault(label, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            a["launches"] += 1
            a["ms"] += s.elapsed_time(e)
            a["flops"] += flops
            a["bytes"] += nbytes
        return agg


def online_latency(flow, vo, dev_frames, reps=10, warm=5):
    """Wall-clock time per frame of the reference's online loop shape: flow_net(im1, im2, iters=12, test_mode=True) ->
    odometry_net(flow) -> host read of (rot, tr), batch 1, frames already on the device at the SLAM size.  The warm-up covers the
    one-off work of the drop-ins' forward(): eager first sightings of the (shape, fmap-reuse) variants and their graph captures."""
    reps = max(1, min(reps, dev_frames.shape[0] - 1 - warm))
    vo.reset_lstm()

    def frame(t):
        _, up = flow(dev_frames[t:t + 1], dev_frames[t + 1:t + 2], iters=12, test_mode=True)
        rot, tr = vo(up)
        return rot.cpu(), tr.cpu()

    for t in range(warm):
        frame(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(warm, warm + reps):
        frame(t)
    dt = (time.perf_counter() - t0) / reps
    vo.reset_lstm()
    return {"ms_per_pair": 1e3 * dt, "pairs_per_s": 1.0 / dt, "pairs": reps,
            "note": "batch 1, forward() graph replays, host sync per frame (wall clock); the metric above batches 54 pairs"}


def localization_extras(dev, frame, pk, keyframes=8192, reps=10):
    """Rows A16/A17 of SURVEY.md section 8 (not part of the odometry step): keyframe embedding latency of one
    376x1232 frame and the similarity search over a synthetic database that does not fit the L2 (K x 15360 fp32),
    as achieved HBM GB/s.  CUDA events on the launching stream."""
    from atdn_vslam_b200 import synth
    from atdn_vslam_b200.localization import KeyframeIndex, MappingEncoder
    enc = MappingEncoder()
    enc.load_state_dict(synth.vae_state_dict())
    enc = enc.to(dev).eval()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    mu = enc.embed(frame)
    embed_ms = timed(lambda: enc.embed(frame))
    dim = mu[0].numel()
    index = KeyframeIndex(dim=dim, capacity=keyframes + 1, device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    index.add(torch.randn(keyframes, dim, device=dev, generator=g))
    index.add(mu)                                     # the query's own embedding: the search must return the last row
    search_ms = timed(lambda: index.search_device(mu))
    found = int(index.search_device(mu)[0].item())
    nbytes = float(len(index) * dim * 4)
    gbs = nbytes / (search_ms * 1e-3) / 1e9
    return {"embed_ms_per_frame": embed_ms, "embedding_dim": dim,
            "search": {"keyframes": len(index), "ms": search_ms, "gbs": gbs, "frac_of_hbm_peak": gbs / pk["hbm_gbs"],
                       "found_planted_row": found == len(index) - 1}}


def peaked_attention_extra(flow, dev, dev_frames, batch_pairs, pk, temps=(1.0, 3.0, 8.0), reps=3):
    """The seeded weights give nearly flat attention rows, the regime in which the mixed fp16 / e4m3 storage keeps EVERY block
    of the probabilities in e4m3 -- the best case for P.V.  Trained attention is peaked: here q . k is scaled by ``temp``
    (tools/fp8_attention_sensitivity.py) on a copy of the flow net, the attention of one batch is rebuilt from real context
    features and one to_v + P.V iteration is timed next to the share of blocks that stayed fp16."""
    from atdn_vslam_b200 import gma
    out = []
    frames = dev_frames[: batch_pairs + 1]
    for temp in temps:
        m = type(flow)(flow.args)
        sd = {k: v.clone() for k, v in flow.state_dict().items()}
        key = next(k for k in sd if k.endswith("att.to_qk.weight"))
        sd[key] = sd[key] * (temp ** 0.5)
        m.load_state_dict(sd)
        m = m.to(dev).eval()
        m.capture_forward = False
        m.forward_frames(frames, iters=1, test_mode=True)          # fills the plan: context features, attention, motion features
        plan = m._plan(batch_pairs, frames.shape[-2], frames.shape[-1], dev)
        wts = m._weights(dev)
        m._aggregate(plan, wts)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            m._aggregate(plan, wts)
        s1.record()
        torch.cuda.synchronize()
        ent = {"logit_scale": temp, "to_v_plus_pv_ms": round(s0.elapsed_time(s1) / reps, 4)}
        if getattr(plan, "mixed", False):
            ent["fp16_blocks"] = round(float(plan.p_hot2.float().mean()), 4)
            ent["hot_sub_blocks"] = round(float(plan.p_hot.float().mean()), 4)
        out.append(ent)
        del m, plan, wts
        torch.cuda.empty_cache()
    return {"batch_pairs": batch_pairs, "regimes": out,
            "note": "one to_v + P.V iteration per batch; logit_scale 1 = the benchmarked weights, 3 / 8 = peaked rows as in trained GMA"}


def training_shape_extra(flow, vo, dev, reps=2):
    """BASELINE.json configs[2] (train_odometry.py:32-48): 24 sequences x 7 frames = 144 frame pairs at 376x1241 (resized to
    the SLAM size), flow for every pair, then ATDNVO on batch 24 x 6 steps (one stateful scan with a reset before it).
    CUDA events on the launching stream; frames resident."""
    from atdn_vslam_b200 import synth
    from atdn_vslam_b200.sequence import preprocess
    seqs, steps, per = 24, 6, 8                      # 8 sequences (48 pairs) per flow batch
    frames = torch.stack([preprocess(synth.frame_sequence(steps + 1, H_RAW, W_RAW, seed=500 + i).to(dev)) for i in range(seqs)])   # [24,7,3,376,1232]
    vo24 = type(vo)(batch_size=seqs)
    vo24.load_state_dict(vo.state_dict())
    vo24 = vo24.to(dev).eval()

    def step():
        feats = []
        for s in range(0, seqs, per):
            im1 = frames[s:s + per, :-1].reshape(-1, 3, *frames.shape[-2:])
            im2 = frames[s:s + per, 1:].reshape(-1, 3, *frames.shape[-2:])
            _, up = flow(im1, im2, iters=12, test_mode=True)
            feats.append(vo24.encode(up).view(per, steps, 512))
        f = torch.cat(feats, 0).transpose(0, 1).contiguous()      # [6, 24, 512]
        vo24.reset_lstm()
        return vo24.recurrent_scan(f)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(reps):
        rot, tr = step()
    s1.record()
    torch.cuda.synchronize()
    ms = s0.elapsed_time(s1) / reps
    return {"workload": "24 sequences x 6 pairs (144 pairs), flow batch 48 through forward(), ATDNVO batch 24 x 6-step scan",
            "ms_per_step": ms, "pairs_per_s": seqs * steps / (ms / 1e3), "finite": bool(torch.isfinite(rot).all() and torch.isfinite(tr).all())}


def sharded_search_extra(dev, rank, world, pk, rows_per_gpu=131072, dim=15360, reps=5):
    """BASELINE.json configs[4] (neural_slam.py:373-384 over a synthetic database): K = 1 048 576 / 8 keyframe embeddings PER GPU
    (8 GB of fp32 rows each), row-sharded across the ranks, planted duplicates of the query on two different ranks: the global
    first minimum must be the LOWER global index.  Reports the local scan as HBM GB/s per GPU (max time over ranks)."""
    import torch.distributed as dist
    from atdn_vslam_b200.localization import KeyframeIndex
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    index = KeyframeIndex(dim=dim, capacity=rows_per_gpu, device=dev)
    for s in range(0, rows_per_gpu, 8192):
        index.add(torch.randn(min(8192, rows_per_gpu - s), dim, device=dev, generator=g))
    q = torch.randn(dim, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    # plant the query itself at local row 1234 of the LAST rank and at local row 77 of the middle rank
    plant = {world - 1: 1234, world // 2: 77} if world > 1 else {0: 1234}
    if rank in plant:
        index._db[plant[rank]] = q
    expect = min(r * rows_per_gpu + i for r, i in plant.items())
    if world > 1:
        gi, gd = index.search_sharded(q)
    else:
        gi, d = index.search(q)
        gd = float(d[gi])
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(reps):
        index.search_device(q)
    s1.record()
    torch.cuda.synchronize()
    t = torch.tensor([s0.elapsed_time(s1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    gbs = rows_per_gpu * dim * 4 / (ms * 1e-3) / 1e9
    return {"keyframes_total": rows_per_gpu * world, "rows_per_gpu": rows_per_gpu, "embedding_dim": dim, "local_scan_ms_max": ms,
            "gbs_per_gpu": gbs, "frac_of_hbm_peak": gbs / pk["hbm_gbs"], "global_index": int(gi), "expected_index": int(expect),
            "index_exact": int(gi) == int(expect), "distance": float(gd)}


def torch_eager_extra(dev, dev_frames, pairs=6):
    """The 'bar to beat' of SURVEY.md 8(d): the reference's own CUDA path = PyTorch eager (cuDNN / cuBLAS) under its fp16 autocast
    regions, one pair per call with a host read per frame -- here the oracle port (same ATen calls, same cast regions) on THIS
    GPU.  Wall clock over consecutive pairs after 2 warm-up pairs."""
    from atdn_vslam_b200 import synth
    from oracle import clvo_oracle, gma_oracle
    gsd = {k: v.to(dev) for k, v in synth.gma_state_dict().items()}
    vsd = {k: v.to(dev) for k, v in synth.atdnvo_state_dict().items()}
    st = clvo_oracle.zero_state(device=dev)

    def pair(t):
        _, up = gma_oracle.raftgma_forward(gsd, dev_frames[t:t + 1], dev_frames[t + 1:t + 2], iters=12, aten_ops=True, mixed_precision=True)
        rot, tr = clvo_oracle.atdnvo_forward(vsd, up.float(), st)
        return rot.cpu(), tr.cpu()

    for t in range(2):
        pair(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(2, 2 + pairs):
        pair(t)
    dt = (time.perf_counter() - t0) / pairs
    return {"ms_per_pair": 1e3 * dt, "pairs_per_s": 1.0 / dt, "pairs": pairs,
            "note": "oracle port on cuda: torch eager, cuDNN/cuBLAS, fp16 autocast regions of network.py:85,93,112, batch 1, host read per pair"}


def run_ours(args):
    import torch.distributed as dist
    from atdn_vslam_b200 import _lib as L, synth
    from atdn_vslam_b200.gma import RAFTGMA
    from atdn_vslam_b200.odometry import ATDNVO
    from atdn_vslam_b200.sequence import OdometryPipeline, preprocess

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")      # its banner goes to stdout, which carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    L.check(L.load().atdn_check_device(local), "atdn_check_device")

    class Args:
        mixed_precision, num_heads, position_only, position_and_content = True, 1, False, False

        def __contains__(self, k):
            return hasattr(self, k)

    flow = RAFTGMA(Args())
    flow.load_state_dict(synth.gma_state_dict(module_prefix=True))
    flow = flow.to(dev).eval()
    vo = ATDNVO()
    vo.load_state_dict(synth.atdnvo_state_dict())
    vo = vo.to(dev).eval()
    pipe = OdometryPipeline(flow, vo, batch_pairs=args.batch_pairs, iters=12, use_graphs=not args.no_graphs)

    from atdn_vslam_b200.sequence import interleaved_rounds, local_frame_ranges
    sharded = world > 1 or args.seq_frames is not None
    if sharded:
        # ONE sequence, batch-interleaved across the ranks; every rank materialises only the frames of its own batches
        seq_frames = args.seq_frames or SEQ_FRAMES
        total_pairs = seq_frames - 1
        rounds = interleaved_rounds(total_pairs, world, args.batch_pairs, args.lead_pairs)
        idx = [t for s, e in local_frame_ranges(rounds, rank) for t in range(s, e + 1)]
        host_frames = synth.frame_sequence(0, H_RAW, W_RAW, seed=synth.FRAME_SEED, indices=idx, dtype=torch.uint8).pin_memory()
        dev_frames = torch.cat([preprocess(host_frames[i:i + 64].to(dev).float()) for i in range(0, host_frames.shape[0], 64)], 0)
        pairs = sum(row[rank][1] - row[rank][0] for row in rounds)
        workload = f"seq{seq_frames}_376x1241_gma12_clvo"
    else:
        pairs = args.frames - 1
        host_frames = synth.frame_sequence(args.frames, H_RAW, W_RAW, seed=synth.FRAME_SEED).pin_memory()
        dev_frames = preprocess(host_frames.to(dev))
        total_pairs = pairs
        rounds = None
        workload = f"seq{args.frames}_376x1241_gma12_clvo"
    group = None if world == 1 else dist.group.WORLD

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        vo.reset_lstm()
        if sharded:
            return pipe.run_interleaved(dev_frames, rounds, group=group, chain=False)[:2]
        return vo.recurrent_scan(pipe.pair_features(dev_frames))

    def step_e2e():
        vo.reset_lstm()
        if sharded:
            rot, tr, poses, keys = pipe.run_interleaved(host_frames, rounds, group=group)
        else:
            rot, tr, poses, keys = pipe.run(host_frames)
        return poses, keys

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = L.LAUNCHES
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e) / steps
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), L.LAUNCHES - launches0

    for _ in range(max(args.warmup, 3)):
        step_resident()
    with ClockSampler(local) as clocks:
        ms, launches = timed(step_resident, args.steps)
    step_e2e()
    ms_e2e, _ = timed(step_e2e, max(1, min(args.steps, 3)))

    value = total_pairs / (ms / 1e3)
    h2d = torch.tensor([host_frames.numel() * host_frames.element_size()], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(h2d)
    # exchange / scan timings of one extra sharded step (CUDA events on rank 0's compute stream)
    shard_timing = None
    if sharded:
        pipe.timing = {}
        step_resident()
        torch.cuda.synchronize()
        evs = pipe.timing.get("events", [])
        pipe.timing = None
        if rank == 0:
            shard_timing = {"rounds": len(rounds), "lead_pairs": rounds[0][0][1] - rounds[0][0][0], "pairs_rank0": pairs,
                            "scan_ms": round(sum(t1.elapsed_time(t2) for _, t1, t2 in evs), 3),
                            "allgather_wait_ms": round(sum(t0.elapsed_time(t1) for t0, t1, _ in evs), 3),
                            "note": "per step on rank 0: serial LSTM scan chunks (run between its batches, one round behind) and the "
                                    "time its compute stream waited for the asynchronous feature all-gathers"}
    # graph replays launch the captured kernels without passing through the C ABI again: count them
    if pipe.use_graphs:
        launches = None

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": workload, "pairs_per_step": total_pairs, "pairs_this_rank": pairs,
                       "image": "376x1241 -> SLAM resize 376x1232", "iters": 12, "batch_pairs": args.batch_pairs,
                       "cuda_graphs": pipe.use_graphs,
                       "parallelism": (f"one sequence, batch-interleaved over {world} rank(s); async all-gather of [pairs,512] features per round, "
                                       "serial scan on rank 0 one round behind, broadcast of [P,6] poses") if sharded else "single GPU",
                       "host_frames": str(host_frames.dtype).replace("torch.", ""),
                       "l2": "per-batch working set (corr pyramid + attention, ~0.6 GB/pair) exceeds the 126 MB L2"},
            "e2e": {"value": total_pairs / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": int(total_pairs * 6 * 4) * world},
            "clocks": clocks.summary()}
    if shard_timing is not None:
        line["sharding"] = shard_timing

    pk = peaks()
    if not args.no_extras:
        try:     # every rank takes part (row-sharded database, NCCL exchange of the per-rank minima)
            line["sharded_keyframe_search"] = sharded_search_extra(dev, rank, world, pk)
        except Exception as exc:
            line["sharded_keyframe_search"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        torch.cuda.empty_cache()
    if rank == 0:
        # ---- per-kernel timing (eager, events on the launching stream) + roofline of the dominant kernel
        prof = KernelProfile()
        eager = OdometryPipeline(flow, vo, batch_pairs=args.batch_pairs, iters=12, use_graphs=False)
        eager.pair_features(dev_frames[: args.batch_pairs + 1])
        torch.cuda.synchronize()
        l0 = L.LAUNCHES
        L.PROFILER = prof
        t0 = time.perf_counter()
        eager.pair_features(dev_frames[: args.batch_pairs + 1])
        torch.cuda.synchronize()
        L.PROFILER = None
        per_batch_launches = L.LAUNCHES - l0
        table = prof.table()
        batches = -(-pairs // args.batch_pairs)
        l1 = L.LAUNCHES
        vo.recurrent_scan(torch.zeros(1, 512, device=dev))
        scan_launches = L.LAUNCHES - l1
        # kernels launched inside the timed region (graph replays re-launch the captured kernels)
        scans = len(rounds) if sharded else 1
        line["gpu_launches"] = int(args.steps * (per_batch_launches * batches + scan_launches * scans))
        kernels = {}
        for label, a in sorted(table.items(), key=lambda kv: -kv[1]["ms"]):
            ent = {"launches": a["launches"], "ms_per_batch": round(a["ms"], 4)}
            if a["flops"]:
                ent["tflops"] = round(a["flops"] / (a["ms"] * 1e-3) / 1e12, 2)
                ent["frac_of_tensor_peak"] = round(ent["tflops"] / pk["tflops_sustained"], 4)
            if a["bytes"]:
                ent["gbs"] = round(a["bytes"] / (a["ms"] * 1e-3) / 1e9, 1)
                ent["frac_of_hbm_peak"] = round(ent["gbs"] / pk["hbm_gbs"], 4)
            kernels[label] = ent
        line["kernels"] = kernels
        # the two kernels BASELINE.json's metric names, against its 60% targets (corr GEMM vs tensor peak, lookup vs HBM)
        line["north_star_kernels"] = {
            "corr_pyramid_frac_of_tensor_peak": kernels.get("corr_pyramid", {}).get("frac_of_tensor_peak"),
            "corr_pyramid_frac_of_hbm_peak": kernels.get("corr_pyramid", {}).get("frac_of_hbm_peak"),
            "corr_lookup_frac_of_hbm_peak": kernels.get("corr_lookup", {}).get("frac_of_hbm_peak"),
            "target": 0.6}
        # roofline of the dominant kernel: the binding roofline is the one with the larger time floor for the
        # kernel's algorithmic work (flops / tensor peak vs bytes / HBM peak); kernels inside the long step are
        # compared with the SUSTAINED bf16 peak, HBM with the measured copy bandwidth
        lab, a = max(table.items(), key=lambda kv: kv[1]["ms"])
        t_tensor = a["flops"] / (pk["tflops_sustained"] * 1e12)
        t_hbm = a["bytes"] / (pk["hbm_gbs"] * 1e9)
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # dram bytes per launch from the committed ncu --set full captures

        def ncu_traffic(label):
            """(dram bytes per launch at this batch size, note) from the committed captures, or (None, None)."""
            if not os.path.exists(tp):
                return None, None
            caps = json.load(open(tp)).get(label, [])
            exact = [c for c in caps if c.get("batch_pairs") == args.batch_pairs]
            if exact:
                return exact[-1]["dram_bytes_per_launch"], None
            if caps:   # the launch processes independent pairs: bytes scale with the pairs per launch
                c = caps[-1]
                return c["dram_bytes_per_launch"] * args.batch_pairs / c["batch_pairs"], f"scaled from the batch_pairs={c['batch_pairs']} capture"
            return None, None

        traffic, traffic_note = ncu_traffic(lab)
        for key in ("corr_pyramid", "corr_lookup"):     # the two north-star kernels: measured DRAM traffic next to the algorithmic bytes
            t, _ = ncu_traffic(key)
            if t is not None and key in table:
                line["north_star_kernels"][key + "_traffic_bytes_per_launch"] = t
                line["north_star_kernels"][key + "_algorithmic_bytes_per_launch"] = table[key]["bytes"] / table[key]["launches"]
        if t_tensor >= t_hbm:
            ach = a["flops"] / (a["ms"] * 1e-3) / 1e12
            line["roofline"] = {"kernel": lab, "bound": "tensor", "achieved": ach, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                                "frac": ach / pk["tflops_sustained"], "traffic": traffic, "peak_source": pk["source"] + " sustained bf16"}
        else:
            ach = a["bytes"] / (a["ms"] * 1e-3) / 1e9
            line["roofline"] = {"kernel": lab, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                "frac": ach / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk["source"] + " copy bandwidth",
                                "algorithmic_bytes_per_launch": a["bytes"] / a["launches"], "launches": a["launches"]}
        if traffic_note:
            line["roofline"]["traffic_note"] = traffic_note
        # ---- online call pattern of NeuralSLAM.__call__ (neural_slam.py:202-204): ONE pair per call through the drop-in
        # forward()s, eager launches, relative pose read back on the host after every frame.  Latency, not the metric.
        try:
            line["online_b1"] = online_latency(flow, vo, dev_frames)
        except Exception as exc:   # an extra: never lose the bench line over it
            line["online_b1"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        try:
            line["localization"] = localization_extras(dev, dev_frames[0:1], pk)
        except Exception as exc:
            line["localization"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        if world == 1 and not args.no_extras:
            try:
                line["training_shape_b24x6"] = training_shape_extra(flow, vo, dev)
            except Exception as exc:
                line["training_shape_b24x6"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
            try:
                line["attention_regimes"] = peaked_attention_extra(flow, dev, dev_frames, args.batch_pairs, pk)
            except Exception as exc:
                line["attention_regimes"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
            try:
                line["torch_eager_b200"] = torch_eager_extra(dev, dev_frames)
            except Exception as exc:
                line["torch_eager_b200"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        # ---- CPU baseline (oracle port of the reference device=cpu path) on a bounded sample
        if world == 1 and not args.no_cpu_baseline:
            v, dt = cpu_pairs_per_s(args.cpu_pairs, warm=1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{args.cpu_pairs} consecutive 376x1232 pairs, iters=12 + CLVO, after 1 warm-up pair ({dt:.1f} s)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--seq-frames", type=int, default=None, help="length of the ONE sharded sequence (default 4541 at N>1)")
    ap.add_argument("--lead-pairs", type=int, default=None, help="pairs per round of rank 0, the sequencer (default: sequence.default_lead_pairs)")
    ap.add_argument("--batch-pairs", type=int, default=54)
    ap.add_argument("--cpu-pairs", type=int, default=12)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[2] / configs[4] / torch-eager extras")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
