"""CPU-side checks: the C-ABI library loads and exports every symbol include/atdn_b200.h declares,
descriptor structs match the header, the drop-in modules accept the reference state dicts, packing
and the host pose / sharding logic behave (no compute kernels are called here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

from atdn_vslam_b200 import _lib as L
from atdn_vslam_b200 import ops, schema, synth
from atdn_vslam_b200.poses import PoseChain, transform
from atdn_vslam_b200.sequence import shard_ranges
from atdn_vslam_b200.localization import merge_shard_minima
from oracle import clvo_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Args:
    mixed_precision, num_heads, position_only, position_and_content = True, 1, False, False

    def __contains__(self, k):
        return hasattr(self, k)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "atdn_b200.h")).read()
    declared = set(re.findall(r"\b(atdn_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.atdn_version() == 100


def test_struct_layouts_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "atdn_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(atdn_tc_desc), offsetof(atdn_tc_desc, corr_w), sizeof(atdn_conv32_desc), offsetof(atdn_conv32_desc, mish));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c, d = map(int, subprocess.check_output([str(exe)]).split())
    assert (a, b) == (C.sizeof(L.TcDesc), L.TcDesc.corr_w.offset)
    assert (c, d) == (C.sizeof(L.Conv32Desc), L.Conv32Desc.mish.offset)


def test_no_cpu_fallback():
    from atdn_vslam_b200.gma import RAFTGMA
    m = RAFTGMA(Args())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 128, 160), torch.zeros(1, 3, 128, 160), iters=1, test_mode=True)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "atdn_vslam_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            txt = open(os.path.join(pkg, f)).read()
            assert "oracle" not in re.sub(r"#.*", "", txt).replace("oracle port", ""), f


def test_reference_state_dicts_load_unchanged():
    from atdn_vslam_b200.gma import RAFTGMA
    from atdn_vslam_b200.odometry import ATDNVO
    from atdn_vslam_b200.localization import MappingEncoder
    g = RAFTGMA(Args())
    assert len(g.state_dict()) == 185
    assert g.load_state_dict(synth.gma_state_dict(module_prefix=True)).missing_keys == []
    assert g.load_state_dict(synth.gma_state_dict()).unexpected_keys == []
    assert g.args.corr_levels == 4 and g.args.corr_radius == 4 and g.args.dropout == 0
    vo = ATDNVO()
    assert len(vo.state_dict()) == 127 and sum(p.numel() for p in vo.parameters()) > 5_000_000
    vo.load_state_dict(synth.atdnvo_state_dict())
    assert vo.suffix == "_c" and vo.batch_size == 1 and tuple(vo.lstm1_h.shape) == (1, 512)
    enc = MappingEncoder()
    sd = dict(synth.vae_state_dict())
    sd["decoder.0.skip_layer.weight"] = torch.zeros(128, 128, 1, 1)     # decoder tensors are ignored
    enc.load_state_dict(sd)


def test_weight_packing_layout():
    w = torch.arange(2 * 3 * 1 * 5, dtype=torch.float32).reshape(2, 3, 1, 5)
    wp = ops.pack_conv_weight(w)
    assert tuple(wp.shape) == (2, 5 * 64)
    # K index = tap * 64 + channel
    assert wp[1, 2 * 64 + 1] == w[1, 1, 0, 2] and wp[0, 3] == 0
    assert tuple(ops.pack_rows_weight(torch.ones(4, 147)).shape) == (4, 192)
    assert ops.pad_bias(torch.ones(126)).numel() == 128
    assert ops.pyramid_shapes(47, 154) == [(47, 154, 160), (23, 77, 80), (11, 38, 40), (5, 19, 24)]


def test_pose_chain_matches_oracle():
    g = torch.Generator().manual_seed(0)
    rots = torch.randn(40, 3, generator=g) * 0.03
    trs = torch.randn(40, 3, generator=g) * 2.0
    poses, keys = PoseChain().extend(rots, trs)
    o_poses, o_keys = clvo_oracle.chain_and_keyframes(rots, trs)
    assert keys == o_keys and len(keys) > 2
    assert torch.equal(poses, o_poses)
    assert torch.equal(transform(rots[0], trs[0]), clvo_oracle.transform(rots[0], trs[0]))
    # the native whole-sequence loop (atdn_pose_chain, a HOST function of the C ABI) == the per-step push() loop,
    # and a chain built by extend() can be continued with push()
    inc = PoseChain()
    for t in range(40):
        inc.push(rots[t], trs[t])
    assert inc.keyframes == keys and torch.equal(torch.stack(inc.poses), poses)
    mixed = PoseChain()
    mixed.extend(rots[:25], trs[:25])
    for t in range(25, 40):
        mixed.push(rots[t], trs[t])
    assert mixed.keyframes == keys
    assert (torch.stack(mixed.poses) - poses).abs().max() <= 1e-5 * poses.abs().max()
    # long sequence (KITTI seq 00 length): same keyframes, poses within 1e-6 relative of the oracle chain
    rots = torch.randn(4540, 3, generator=g) * 0.02
    trs = torch.randn(4540, 3, generator=g) * 1.2
    poses, keys = PoseChain().extend(rots, trs)
    o_poses, o_keys = clvo_oracle.chain_and_keyframes(rots, trs)
    assert keys == o_keys and (poses - o_poses).norm() <= 1e-6 * o_poses.norm()


def test_shard_ranges_and_minima_merge():
    assert shard_ranges(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_ranges(4540, 8)[-1][1] == 4540
    assert merge_shard_minima([(2.0, 7), (1.0, 20), (1.0, 12), (float("inf"), -1)]) == (12, 1.0)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from atdn_vslam_b200.sequence import gather_features, shard_ranges
    total = 7
    s, e = shard_ranges(total, world)[rank]
    local = torch.arange(s, e, dtype=torch.float32).view(-1, 1).repeat(1, 4)
    full = gather_features(local, total)
    q.put((rank, full[:, 0].tolist()))
    dist.destroy_process_group()


def _gloo_search_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from atdn_vslam_b200.localization import global_first_minimum
    from oracle import clvo_oracle
    g = torch.Generator().manual_seed(7)
    db = torch.randn(11, 32, generator=g)
    db[9] = db[2]                                   # planted duplicate: rows 2 (rank 0) and 9 (rank 1) tie
    code = db[2] + 0.0
    bounds = [(0, 6), (6, 11)]
    out = []
    for case in range(3):
        if case == 1:                               # the minimum lives on the last rank only
            code = db[10] + 0.0
        if case == 2:                               # rank 0 holds nothing
            bounds = [(0, 0), (0, 11)]
        s, e = bounds[rank]
        if e > s:
            k, d = clvo_oracle.keyframe_search(db[s:e], code)      # stand-in for the local CUDA search
            local = (d[k].reshape(()), torch.tensor(k))
        else:
            local = None
        gi, gd = global_first_minimum(local, e - s, torch.device("cpu"))
        ok, od = clvo_oracle.keyframe_search(db, code)
        out.append((gi, ok, abs(gd - float(od[ok])) < 1e-6))
    q.put((rank, out))
    dist.destroy_process_group()


def test_sharded_keyframe_minimum_world2_gloo():
    """Exchange step of the sharded search (localization.global_first_minimum) == the serial first minimum."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_search_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, out in res:
        assert [o[0] for o in out] == [2, 10, 10]
        assert all(o[0] == o[1] and o[2] for o in out)


def test_feature_gather_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, vals in res:
        assert vals == [float(i) for i in range(7)]


def test_tiled_state_layout_matches_the_header_formula():
    """ops.state_from_nhwc / state_to_nhwc implement the index formula documented for atdn_tc_desc.h32 / z32."""
    import torch
    from atdn_vslam_b200 import ops
    g = torch.Generator().manual_seed(0)
    b, h, w = 2, 37, 45
    x = torch.randn(b, h, w, 128, generator=g)
    t = ops.state_from_nhwc(x)
    th_n, tw_n = (h + 15) // 16, (w + 7) // 8
    assert tuple(t.shape) == (b, th_n, tw_n, 4, 32, 32, 4) and t.numel() == b * th_n * 16 * tw_n * 8 * 128
    flat = t.flatten()
    for (bb, hh, ww, c) in [(0, 0, 0, 0), (1, 20, 13, 77), (0, 36, 44, 127), (1, 15, 8, 3), (0, 16, 7, 64)]:
        r = ((hh & 15) << 3) | (ww & 7)
        idx = ((((bb * th_n + (hh >> 4)) * tw_n + (ww >> 3)) * 4 + (r >> 5)) * 32 + c // 4) * 128 + (r & 31) * 4 + c % 4
        assert flat[idx] == x[bb, hh, ww, c]
    assert torch.equal(ops.state_to_nhwc(t, h, w), x)


def test_strip_pyramid_layout_matches_the_header_formula():
    """ops.pyramid_untile inverts the strip-layout formulas documented for atdn_corr_pyramid (half_levels = 4)."""
    import torch
    from atdn_vslam_b200 import ops
    h8, w8 = 23, 39
    th, tw = (h8 + 7) // 8, (w8 + 31) // 32
    tw3 = (tw + 1) // 2 * 2
    levels, refs = [], []
    for l in range(4):
        H, W = h8 >> l, w8 >> l
        ref = (torch.arange(2 * H * W) % 2048).float().view(2, H, W)
        chunk = (256, 64, 16, 4)[l]
        t = torch.full((2, th * tw if l < 3 else th * tw3, chunk), -1.0)
        for y in range(H):
            for x in range(W):
                if l < 3:
                    tile = (y >> (3 - l)) * tw + (x >> (5 - l))
                    off = ((x >> 3) & ((4 >> l) - 1)) * (64 >> l) * (1 if l < 2 else 0) + (y & ((8 >> l) - 1)) * 8 + (x & 7)
                    # the general form of the header: rowpart(y) + (x >> 3) * (64 >> l) + (x & 7)
                    flat = (y >> (3 - l)) * tw * chunk + (y & ((8 >> l) - 1)) * 8 + (x >> 3) * (64 >> l) + (x & 7)
                    assert flat == tile * chunk + off
                else:
                    tile, off = y * tw3 + (x >> 2), x & 3
                    assert tile * 4 + off == y * tw3 * 4 + (x >> 3) * 8 + (x & 7)
                t[:, tile, off] = ref[:, y, x]
        levels.append(t.half())
        refs.append(ref)
    for got, ref in zip(ops.pyramid_untile(levels, h8, w8), refs):
        assert torch.equal(got, ref)
    assert ops.stats_parts(188, 616, True) == 12 * 10 * 2 * 4 and ops.stats_parts(188, 616, False) == 12 * 20 * 4


def test_flag_and_epilogue_constants_match_header(tmp_path):
    """_lib.py mirrors the enums of include/atdn_b200.h by value (ctypes has no access to C enums)."""
    names = ["ATDN_EPI_STORE16", "ATDN_EPI_STORE32", "ATDN_EPI_CORR", "ATDN_EPI_GRU_ZR", "ATDN_EPI_GRU_Q", "ATDN_EPI_PV", "ATDN_EPI_FLOW",
             "ATDN_F_RELU", "ATDN_F_RESID", "ATDN_F_FLOWTAIL", "ATDN_F_TANH_LO", "ATDN_F_B_BATCHED", "ATDN_F_A_SHARED", "ATDN_F_PAIR",
             "ATDN_F_STATS", "ATDN_F_TILED32", "ATDN_F_PRE16", "ATDN_F_Z16", "ATDN_F_H16", "ATDN_MODE_ROWS", "ATDN_MODE_PATCH"]
    src = tmp_path / "en.c"
    src.write_text('#include <stdio.h>\n#include "atdn_b200.h"\nint main(){printf("' + " ".join(["%d"] * len(names)) + '\\n", ' +
                   ", ".join(f"(int){n}" for n in names) + ");return 0;}\n")
    exe = tmp_path / "en"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    vals = dict(zip(names, map(int, subprocess.check_output([str(exe)]).split())))
    for n, v in vals.items():
        py = n.replace("ATDN_", "")
        assert getattr(L, py) == v, (n, v, getattr(L, py))


def test_batch_ranges_cover_every_pair_once():
    from atdn_vslam_b200.sequence import batch_ranges
    assert batch_ranges(270, 54) == [(0, 54), (54, 108), (108, 162), (162, 216), (216, 270)]
    assert batch_ranges(270, 54, short_first=True) == [(0, 13), (13, 67), (67, 121), (121, 175), (175, 229), (229, 270)]
    assert batch_ranges(5, 54, short_first=True) == [(0, 5)]            # a single batch is never split
    assert batch_ranges(0, 8) == [] and batch_ranges(3, 2, short_first=True) == [(0, 1), (1, 3)]
    for n, b, sf in [(4540, 54, True), (271, 27, False), (7, 3, True), (1, 1, True)]:
        r = batch_ranges(n, b, sf)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == c[0] for a, c in zip(r, r[1:]))
        assert all(0 < e - s <= b for s, e in r)


def test_flow_head_tap_decomposition_identity():
    """flow_head.conv2 as '1x1 conv to 18 per-tap products + shifted sum' (gma._Packed.fh2, atdn_flow_head_gather) is the
    3x3 convolution with zero padding: checked in plain PyTorch on the weight layout the product packs."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    b, c, h, w = 2, 256, 9, 11
    x = torch.randn(b, c, h, w, generator=g, dtype=torch.float64)
    w2 = torch.randn(2, c, 3, 3, generator=g, dtype=torch.float64)
    bias = torch.randn(2, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w2, bias, padding=1)
    w18 = w2.permute(2, 3, 0, 1).reshape(18, c)                       # row = (dy*3 + dx)*2 + co, as in gma._Packed
    d = torch.einsum("rc,bchw->bhwr", w18, x)                          # the 1x1 tensor-core conv: d[b, y, x, tap*2 + co]
    out = bias.view(1, 2, 1, 1).repeat(b, 1, h, w).clone()
    for tap in range(9):
        dy, dx = tap // 3 - 1, tap % 3 - 1
        for co in range(2):
            src = d[..., tap * 2 + co]
            ys, xs = slice(max(0, -dy), h - max(0, dy)), slice(max(0, -dx), w - max(0, dx))       # destination pixels p
            yq, xq = slice(max(0, dy), h + min(0, dy)), slice(max(0, dx), w + min(0, dx))         # neighbours p + off(tap)
            out[:, co, ys, xs] += src[:, yq, xq]
    assert (out - ref).abs().max() < 1e-10


def test_gru_context_precompute_identity():
    """conv(W, [h | inp | mf | mfg]) = conv(W[:, rest], [h | mf | mfg]) + conv(W[:, 128:256], inp): the split the product
    uses to take the context features out of the per-iteration SepConvGRU convolutions (gma._Packed.gru / gru_pre)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(4)
    hx = torch.randn(1, 512, 6, 12, generator=g, dtype=torch.float64)
    for cout, k, pad in ((256, (1, 5), (0, 2)), (128, (5, 1), (2, 0))):
        wt = torch.randn(cout, 512, *k, generator=g, dtype=torch.float64)
        bias = torch.randn(cout, generator=g, dtype=torch.float64)
        full = F.conv2d(hx, wt, bias, padding=pad)
        rest_w = torch.cat([wt[:, :128], wt[:, 256:]], 1)
        rest_x = torch.cat([hx[:, :128], hx[:, 256:]], 1)
        pre = F.conv2d(hx[:, 128:256], wt[:, 128:256], bias, padding=pad)
        assert (F.conv2d(rest_x, rest_w, None, padding=pad) + pre - full).abs().max() < 1e-9


def test_nchw_pitch_of_padded_and_degenerate_maps():
    """ops.nchw_pitch: the row pitch of CLVO maps that live in row-padded buffers (154 -> 156 floats); size-1
    dimensions have arbitrary strides in torch and must not decide (or fail) the check."""
    from atdn_vslam_b200 import ops
    buf = torch.zeros(2, 3, 5, 156)
    assert ops.nchw_pitch(buf) == 156
    assert ops.nchw_pitch(buf[..., :154]) == 156
    assert ops.nchw_pitch(torch.zeros(2, 3, 1, 156)[..., :154]) == 156          # H == 1: pitch from the channel stride
    assert ops.nchw_pitch(torch.zeros(2, 1, 1, 156)[..., :154]) == 156          # H == C == 1: from the batch stride
    assert ops.nchw_pitch(torch.zeros(1, 1, 1, 156)[..., :154]) == 154          # nothing to contradict a dense row
    assert ops.nchw_pitch(torch.zeros(1, 16, 5, 8).as_strided((1, 16, 5, 8), (12345, 40, 8, 1))) == 8   # B == 1
    assert ops.nchw_pitch(torch.zeros(4, 512, 1, 1)) == 1
    with pytest.raises(AssertionError):
        ops.nchw_pitch(torch.zeros(2, 3, 5, 8).permute(0, 1, 3, 2))            # transposed rows
    with pytest.raises(AssertionError):
        ops.nchw_pitch(torch.zeros(2, 6, 5, 8)[:, ::2])                        # channel gaps


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs first): one JSON line with the contract's keys, timed on
    the ATen-op form of the oracle port; no GPU needed."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "frame_pairs_per_s" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["workload"] == "seq271_376x1241_gma12_clvo"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


# ----------------------------------------------------------------------------------------------
# batch-interleaved sharding of one sequence (sequence.run_interleaved): host logic under gloo with stub networks
# ----------------------------------------------------------------------------------------------
def test_interleaved_rounds_cover_the_sequence_in_order():
    from atdn_vslam_b200.sequence import default_lead_pairs, interleaved_rounds, local_frame_ranges
    for pairs, world, bp, lead in [(4540, 8, 54, None), (4540, 2, 54, None), (270, 4, 54, 50), (7, 2, 3, 2), (5, 4, 3, None), (3, 1, 54, None)]:
        rounds = interleaved_rounds(pairs, world, bp, lead)
        flat = [rng for row in rounds for rng in row]
        assert flat[0][0] == 0 and flat[-1][1] == pairs
        assert all(a[1] == b[0] for a, b in zip(flat, flat[1:])) and all(0 <= e - s <= bp for s, e in flat)
        assert all(len(row) == world for row in rounds)
        ld = default_lead_pairs(world, bp) if lead is None else lead
        for row in rounds[:-1]:                      # full rounds: rank 0 leads with the smaller batch
            assert row[0][1] - row[0][0] == ld and all(e - s == bp for s, e in row[1:])
        sizes = [e - s for s, e in rounds[-1]]
        assert max(sizes[1:] or [0]) - min(sizes[1:] or [0]) <= 1        # the remainder is spread evenly
        for r in range(world):
            assert sum(e - s for s, e in local_frame_ranges(rounds, r)) == sum(row[r][1] - row[r][0] for row in rounds)
    assert default_lead_pairs(1, 54) == 54 and default_lead_pairs(8, 54) == 47 and default_lead_pairs(2, 54) == 52


class _StubFlow(torch.nn.Module):
    """Stands in for RAFTGMA on CPU: 'flow' of pair t = mean(frame t+1) - mean(frame t) per pair."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.generation = 0

    def forward_frames(self, frames, iters=12, test_mode=True):
        m = frames.float().mean(dim=(1, 2, 3))
        return None, (m[1:] * 3.0 - m[:-1]).view(-1, 1)


class _StubVO:
    """Stands in for ATDNVO: encode = broadcast to 512 features; the 'scan' is an order-sensitive recurrence."""
    generation = 0

    def __init__(self):
        self.state = torch.zeros(())

    def encode(self, flow):
        return flow.repeat(1, 512)

    def recurrent_scan(self, feats):
        out = []
        for f in feats[:, 0]:
            self.state = self.state * 0.9 + f
            out.append(self.state.clone())
        o = torch.stack(out) if out else torch.zeros(0)
        return torch.stack([o, o * 2, o * 3], 1), torch.stack([o + 1, o + 2, o + 3], 1)


def _interleaved_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from atdn_vslam_b200 import sequence
    from atdn_vslam_b200.sequence import OdometryPipeline, interleaved_rounds, local_frame_ranges
    sequence.SCAN_MIN_PAIRS = 9                                      # several scan calls, each covering >= 2 rounds
    pairs = 23
    g = torch.Generator().manual_seed(1)
    frames = torch.rand(pairs + 1, 3, 376, 1232, generator=g)       # SLAM-sized: preprocess is the identity
    rounds = interleaved_rounds(pairs, world, 4, lead_pairs=3)
    local = torch.cat([frames[s:e + 1] for s, e in local_frame_ranges(rounds, rank)], 0)
    pipe = OdometryPipeline(_StubFlow(), _StubVO(), batch_pairs=4, iters=1, use_graphs=False)
    rot, tr, _, _ = pipe.run_interleaved(local, rounds, chain=False)
    q.put((rank, rot.tolist(), tr.tolist()))
    dist.destroy_process_group()


def test_interleaved_run_world2_gloo_equals_single_rank():
    """R-rank batch-interleaved run == the 1-rank run, bit for bit, on every rank (exchange + lagged scan on rank 0 +
    final broadcast), with an order-sensitive recurrence standing in for the LSTM."""
    import torch.multiprocessing as mp
    from atdn_vslam_b200.sequence import OdometryPipeline, interleaved_rounds
    pairs = 23
    g = torch.Generator().manual_seed(1)
    frames = torch.rand(pairs + 1, 3, 376, 1232, generator=g)
    ref_pipe = OdometryPipeline(_StubFlow(), _StubVO(), batch_pairs=4, iters=1, use_graphs=False)
    rot1, tr1, _, _ = ref_pipe.run_interleaved(frames, interleaved_rounds(pairs, 1, 4), chain=False)
    rot0, tr0, _, _ = OdometryPipeline(_StubFlow(), _StubVO(), batch_pairs=5, iters=1, use_graphs=False).run(frames, chain=False)
    assert torch.equal(rot1, rot0) and torch.equal(tr1, tr0)        # interleaved (world 1) == contiguous
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_interleaved_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, rot, tr in res:
        assert rot == rot1.tolist() and tr == tr1.tolist()


def test_mapping_encoder_is_inference_only():
    """the localization drop-in embeds keyframes; the reference MappingVAE keeps __create_map's training
    (neural_slam.py:305-352).  The class must refuse training mode loudly instead of failing inside the caller."""
    from atdn_vslam_b200.localization import MappingEncoder
    m = MappingEncoder()
    assert not m.training and m.eval() is m
    with pytest.raises(NotImplementedError):
        m.train()
    assert all(not p.requires_grad for p in m.parameters())
