"""End-to-end GPU bring-up against the golden fixtures / oracle, one process per check.
Usage on a GPU box:  python tests/gpu_e2e.py [check ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def _stat(name, got, ref, tol, rel=True):
    import torch
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    scale = float(ref.abs().max()) + 1e-12 if rel else 1.0
    e = float(err.max()) / scale
    print(f"{'PASS' if e <= tol else 'FAIL'} {name}: max_err={float(err.max()):.3e} mean_err={float(err.mean()):.3e} "
          f"ref_max={float(ref.abs().max()):.3e} nan={int(torch.isnan(got).sum())}", flush=True)
    return e <= tol


def _epe(name, got, ref, tol_mean):
    import torch
    got, ref = got.float().cpu(), ref.float().cpu()
    epe = (got - ref).pow(2).sum(1).sqrt()
    ok = float(epe.mean()) <= tol_mean
    print(f"{'PASS' if ok else 'FAIL'} {name}: EPE mean={float(epe.mean()):.3e} p99={float(epe.flatten().quantile(0.99)):.3e} "
          f"max={float(epe.max()):.3e} |flow| mean={float(ref.abs().mean()):.2f} nan={int(torch.isnan(got).sum())}", flush=True)
    return ok


def _gma(dev="cuda"):
    from atdn_vslam_b200 import synth
    from atdn_vslam_b200.gma import RAFTGMA

    class Args:
        mixed_precision, num_heads, position_only, position_and_content = True, 1, False, False

        def __contains__(self, k):
            return hasattr(self, k)

    m = RAFTGMA(Args())
    sd = synth.gma_state_dict(module_prefix=True)
    m.load_state_dict(sd)
    return m.to(dev).eval(), sd


def check_gma_small():
    import numpy as np
    import torch
    from atdn_vslam_b200 import synth
    g = np.load(os.path.join(GOLD, "gma_small.npz"))
    m, sd = _gma()
    assert str(g["digest"]) == synth.state_dict_digest(sd), "weight RNG drift"
    im1 = torch.from_numpy(g["im1"]).float().cuda()
    im2 = torch.from_numpy(g["im2"]).float().cuda()
    lo, up = m(im1, im2, iters=4, test_mode=True)
    torch.cuda.synchronize()
    plan = next(iter(m._plans.values()))
    ok = True
    fm = plan.buffer("fmap", (2, 16, 20, 256), torch.float16).float().permute(0, 3, 1, 2)
    ok &= _stat("fnet fmap1", fm[:1], torch.from_numpy(g["fmap1"]), 5e-3)
    ok &= _stat("fnet fmap2", fm[1:], torch.from_numpy(g["fmap2"]), 5e-3)
    ok &= _stat("cnet inp", plan.hx[..., 128:256].float().permute(0, 3, 1, 2), torch.from_numpy(g["inp"]), 5e-3)
    from atdn_vslam_b200 import ops
    ok &= _stat("corr level 3", ops.pyramid_untile(plan.pyr, 16, 20)[3].reshape(-1), torch.from_numpy(g["pyr3"]).reshape(-1), 5e-3)
    ok &= _epe("small flow_lo vs reference", lo, torch.from_numpy(g["flow_lo"]), 2e-3)
    ok &= _epe("small flow_up vs reference", up, torch.from_numpy(g["flow_up"]), 1e-2)
    preds = m(im1, im2, iters=2, test_mode=False)
    ok &= isinstance(preds, list) and len(preds) == 2 and tuple(preds[0].shape) == (1, 2, 128, 160)
    return ok


def check_gma_stages():
    """Stage-by-stage comparison against the oracle's intermediates (localises a failing kernel)."""
    import numpy as np
    import torch
    from atdn_vslam_b200 import ops, synth
    from oracle import gma_oracle
    g = np.load(os.path.join(GOLD, "gma_small.npz"))
    m, sd = _gma()
    im1 = torch.from_numpy(g["im1"]).float()
    im2 = torch.from_numpy(g["im2"]).float()
    ok = True
    for iters in (1, 2):
        lo_o, up_o, it = gma_oracle.raftgma_forward(sd, im1, im2, iters=iters, return_intermediates=True)
        lo, up = m(im1.cuda(), im2.cuda(), iters=iters, test_mode=True)
        torch.cuda.synchronize()
        plan = next(iter(m._plans.values()))
        ok &= _stat(f"iters={iters} lookup corr (last iter)", plan.corrfeat[..., :324].float().permute(0, 3, 1, 2), it["corr"][-1], 5e-3)
        ok &= _stat(f"iters={iters} net", ops.state_to_nhwc(plan.h32, 16, 20).permute(0, 3, 1, 2), it["net"], 3e-2)
        ok &= _stat(f"iters={iters} mask", plan.mask32.view(1, 16, 20, 576).permute(0, 3, 1, 2), it["mask"], 5e-3)
        ok &= _epe(f"iters={iters} flow_lo", lo, lo_o, 1e-3)
        ok &= _epe(f"iters={iters} flow_up", up, up_o, 5e-3)
    return ok


def check_gma_full():
    import numpy as np
    import torch
    from atdn_vslam_b200 import synth
    g = np.load(os.path.join(GOLD, "gma_full.npz"))
    m, sd = _gma()
    frames = synth.frame_sequence(2, 376, 1232, seed=synth.FRAME_SEED)
    assert abs(float(frames[0].double().sum()) - float(g["frame_sum"][0])) < 0.5, "frame RNG drift"
    im1, im2 = frames[0:1].cuda(), frames[1:2].cuda()
    lo, up = m(im1, im2, iters=12, test_mode=True)
    torch.cuda.synchronize()
    ok = _epe("full flow_lo vs reference", lo, torch.from_numpy(g["flow_lo"]), 2.5e-3)
    ok &= _epe("full flow_up (stride-4 samples) vs reference", up[:, :, ::4, ::4], torch.from_numpy(g["flow_up_s4"]), 1e-2)
    t0 = time.time()
    for _ in range(3):
        m(im1, im2, iters=12, test_mode=True)
    torch.cuda.synchronize()
    print(f"INFO eager (ungraphed) batch-1 latency: {(time.time() - t0) / 3 * 1e3:.1f} ms/pair", flush=True)
    # batch of 2 identical pairs must reproduce the batch-1 result
    lo2, up2 = m(torch.cat([im1, im1]), torch.cat([im2, im2]), iters=12, test_mode=True)
    ok &= _epe("batch-2 vs batch-1", up2[1:], up, 1e-3)
    return ok


def check_atdnvo():
    import numpy as np
    import torch
    from atdn_vslam_b200 import synth
    from atdn_vslam_b200.odometry import ATDNVO
    g = np.load(os.path.join(GOLD, "atdnvo.npz"))
    sd = synth.atdnvo_state_dict()
    assert str(g["digest"]) == synth.state_dict_digest(sd), "weight RNG drift"
    vo = ATDNVO()
    vo.load_state_dict(sd)
    vo = vo.to("cuda").eval()
    flows = synth.synthetic_flows(3, seed=int(g["flow_seed"])).cuda()
    ok = True
    feats = vo.encode(flows)
    ok &= _stat("clvo features (batch 3)", feats, torch.from_numpy(g["feat"]), 1e-4)
    for t in range(3):
        r, tr = vo(flows[t:t + 1])
        ok &= _stat(f"step {t} rot", r, torch.from_numpy(g["rot"][t:t + 1]), 1e-4)
        ok &= _stat(f"step {t} tr", tr, torch.from_numpy(g["tr"][t:t + 1]), 1e-4)
    vo.reset_lstm()
    r, tr = vo.recurrent_scan(feats)
    ok &= _stat("scan rot", r, torch.from_numpy(g["rot"]), 1e-4)
    ok &= _stat("scan tr", tr, torch.from_numpy(g["tr"]), 1e-4)
    return ok


def check_scan_long(t_steps=270, batch=1):
    """Persistent scan kernel vs the per-step kernels (lstm_cell / linear32) and vs the oracle LSTM on a long
    random feature sequence; scanning in two pieces must continue the state exactly."""
    import time
    import torch
    from atdn_vslam_b200 import synth
    from atdn_vslam_b200.odometry import ATDNVO
    from oracle import clvo_oracle
    sd = synth.atdnvo_state_dict()
    vo = ATDNVO(batch_size=batch)
    vo.load_state_dict(sd)
    vo = vo.to("cuda").eval()
    g = torch.Generator().manual_seed(12)
    feats = torch.randn(t_steps, batch, 512, generator=g) * 0.7
    fd = feats.cuda()
    # reference 1: the per-step kernels
    p = vo._weights(fd.device)
    state = [torch.zeros(batch, 512, device="cuda") for _ in range(4)]
    gates, tmp = vo._tmp(batch, fd.device)
    rs, ts = [], []
    for t in range(t_steps):
        r, x = vo._step(p, fd[t], state, gates, tmp)
        rs.append(r)
        ts.append(x)
    r_ref, t_ref = torch.stack(rs), torch.stack(ts)
    vo.reset_lstm()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r1, t1 = vo.recurrent_scan(fd if batch > 1 else fd[:, 0])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if batch == 1:
        r1, t1 = r1.unsqueeze(1), t1.unsqueeze(1)
    ok = _stat(f"scan T={t_steps} B={batch} rot vs per-step kernels ({dt * 1e3:.2f} ms wall)", r1, r_ref, 1e-5)
    ok &= _stat("scan tr vs per-step kernels", t1, t_ref, 1e-5)
    ok &= _stat("scan final h2", vo.lstm2_h, state[2], 1e-5) and _stat("scan final c2", vo.lstm2_c, state[3], 1e-5)
    # reference 2: the oracle (torch CPU fp32 restatement of the reference module) on a prefix
    n = min(t_steps, 16)
    st = clvo_oracle.zero_state(batch)
    o_r = []
    for t in range(n):
        rr, _ = clvo_oracle.atdnvo_recurrent(sd, feats[t], st)
        o_r.append(rr)
    ok &= _stat("scan rot vs oracle (first 16 steps)", r1[:n], torch.stack(o_r), 1e-4)
    # two-piece scan continues the state
    vo.reset_lstm()
    k = t_steps // 3
    f2 = fd if batch > 1 else fd[:, 0]
    ra, _ = vo.recurrent_scan(f2[:k])
    rb, _ = vo.recurrent_scan(f2[k:])
    rc = torch.cat([ra, rb], 0)
    if batch == 1:
        rc = rc.unsqueeze(1)
    ok &= bool(torch.equal(rc, r1))
    print(("PASS" if torch.equal(rc, r1) else "FAIL") + " two-piece scan is bit-identical to one scan", flush=True)
    return ok


def check_localization():
    import numpy as np
    import torch
    from atdn_vslam_b200 import synth
    from atdn_vslam_b200.localization import MappingEncoder, KeyframeIndex
    from oracle import clvo_oracle
    g = np.load(os.path.join(GOLD, "vae.npz"))
    sd = synth.vae_state_dict()
    assert str(g["digest"]) == synth.state_dict_digest(sd), "weight RNG drift"
    enc = MappingEncoder()
    enc.load_state_dict(sd)
    enc = enc.to("cuda").eval()
    img = synth.frame_sequence(1, 376, 1232, seed=21).cuda()
    mu = enc(img)[0]
    ok = _stat("vae mu", mu, torch.from_numpy(g["mu"]), 1e-4)
    db = synth.keyframe_db(512)
    q = db[11] + 0.01 * torch.randn(15360, generator=torch.Generator().manual_seed(9))
    ref_i, ref_d = clvo_oracle.keyframe_search(db, q)
    idx = KeyframeIndex(capacity=16)
    idx.add(db[:100].cuda())
    idx.add(db[100:].cuda())
    i, d = idx.search(q.cuda())
    ok &= _stat("keyframe distances", d, ref_d, 1e-5)
    ok &= (i == ref_i == 3)
    print(f"{'PASS' if i == ref_i == 3 else 'FAIL'} keyframe index {i} (oracle {ref_i}, planted duplicate of 3 at 11)", flush=True)
    return ok


CHECKS = {"gma_stages": check_gma_stages, "gma_small": check_gma_small, "gma_full": check_gma_full,
          "atdnvo": check_atdnvo, "localization": check_localization}


def main():
    names = sys.argv[1:]
    if len(names) == 1 and names[0] in CHECKS and os.environ.get("ATDN_DIAG_CHILD"):
        sys.exit(0 if CHECKS[names[0]]() else 1)
    results = {}
    for n in names or list(CHECKS):
        env = dict(os.environ, ATDN_DIAG_CHILD="1")
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), n], env=env, timeout=600,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            tail = "\n".join(p.stdout.strip().splitlines()[-30:])
            results[n] = p.returncode
            print(f"=== {n} (exit {p.returncode})\n{tail}", flush=True)
        except subprocess.TimeoutExpired:
            results[n] = "timeout"
            print(f"=== {n} TIMEOUT", flush=True)
    print("SUMMARY", results)


if __name__ == "__main__":
    main()
