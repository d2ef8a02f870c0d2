"""GPU bring-up diagnostics: each check runs in its own process (a trapped kernel poisons the CUDA
context) and prints one line.  Usage on a GPU box:  python tests/gpu_diag.py [check ...]"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _cmp(name, got, ref, tol):
    import torch
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    rel = float(err.max() / (ref.abs().max() + 1e-12))
    bad = int((err > tol * (ref.abs().max() + 1e-12)).sum())
    print(f"{'PASS' if rel <= tol else 'FAIL'} {name}: max_abs_err={float(err.max()):.3e} rel={rel:.3e} "
          f"ref_max={float(ref.abs().max()):.3e} bad={bad}/{err.numel()} nan={int(torch.isnan(got).sum())}", flush=True)
    if rel > tol:
        idx = torch.nonzero(err > tol * (ref.abs().max() + 1e-12))[:8]
        for i in idx:
            t = tuple(int(v) for v in i)
            print("    at", t, "got", float(got[t]), "ref", float(ref[t]), flush=True)
    return rel <= tol


def check_rows(m=300, n=200, k=256, bn=128, batch=1, pair=False):
    import torch
    from atdn_vslam_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(1)
    a = (torch.randn(batch, m, k, generator=g)).half().cuda()
    kp = (k + 7) // 8 * 8
    ap = torch.zeros(batch, m, kp, dtype=torch.half, device="cuda")
    ap[:, :, :k] = a
    b = torch.randn(n, k, generator=g).half().cuda()
    bp = ops.pack_rows_weight(b.float())
    npitch = (n + 7) // 8 * 8
    out = torch.full((batch, m, npitch), float("nan"), dtype=torch.float32, device="cuda")
    ops.gemm_rows(L.ptr(ap), k, m, kp, batch, L.ptr(bp), n, bp.shape[1], L.ptr(out), npitch, n_valid=n, bn=bn,
                  epi=L.EPI_STORE32, alpha=0.5, flags=L.F_PAIR if pair else 0)
    torch.cuda.synchronize()
    ref = 0.5 * torch.matmul(a.float().cpu(), b.float().cpu().t())
    return _cmp(f"rows m={m} n={n} k={k} bn={bn} batch={batch} pair={pair}", out[:, :, :n], ref, 2e-5)


def check_conv(cin=64, cout=64, kh=3, kw=3, stride=1, h=20, w=37, batch=2, bn=64, relu=False, resid=False, split=0, pair=False, mt=0):
    import torch
    import torch.nn.functional as F
    from atdn_vslam_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(2)
    x = torch.randn(batch, cin, h, w, generator=g).half()
    wt = (torch.randn(cout, cin, kh, kw, generator=g) / (cin * kh * kw) ** 0.5).half()
    bias = torch.randn(cout, generator=g)
    ph, pw = kh // 2, kw // 2
    ref = F.conv2d(x.float(), wt.float(), bias, stride=stride, padding=(ph, pw))
    oh, ow = ref.shape[2:]
    cp = (cin + 7) // 8 * 8
    xn = torch.zeros(batch, h, w, cp, dtype=torch.half)
    xn[..., :cin] = x.permute(0, 2, 3, 1)
    xn = xn.cuda()
    wp = ops.pack_conv_weight(wt.float().cuda())
    bp = ops.pad_bias(bias.cuda())
    op = (cout + 7) // 8 * 8
    out = torch.full((batch, oh, ow, op), float("nan"), dtype=torch.half, device="cuda")
    flags = (L.F_RELU if relu else 0) | (L.F_RESID if resid else 0) | (L.F_PAIR if pair else 0)
    rs = None
    if resid:
        r = torch.randn(batch, oh, ow, op, generator=g).half()
        rs = ops.View(r.cuda())
        y = F.relu(ref) if relu else ref
        ref = F.relu(r[..., :cout].permute(0, 3, 1, 2).float() + y)
    elif relu:
        ref = F.relu(ref)
    if split:
        a1 = ops.View(xn[..., :split].contiguous())
        a2 = ops.View(xn[..., split:].contiguous())
        ops.conv_tc(a1, wp, bp, ops.View(out), cout=cout, taps=(kh, kw), pad=(ph, pw), stride=stride, bn=bn, flags=flags,
                    a2=a2, resid=rs, mt=mt)
    else:
        ops.conv_tc(ops.View(xn, 0, cin), wp, bp, ops.View(out), cout=cout, taps=(kh, kw), pad=(ph, pw), stride=stride,
                    bn=bn, flags=flags, resid=rs, mt=mt)
    torch.cuda.synchronize()
    got = out[..., :cout].permute(0, 3, 1, 2)
    return _cmp(f"conv cin={cin} cout={cout} k={kh}x{kw} s={stride} {h}x{w} b={batch} bn={bn} relu={relu} resid={resid} split={split} pair={pair} mt={mt}",
                got, ref, 2e-3)



def check_gru(mt=2, kind="zr", h=37, w=45, batch=2, taps=(1, 5), pair=False, bn=128, qbn=None):
    """SepConvGRU epilogues of the halo kernel against a PyTorch fp32 reference of the same op (update.py:48-63)."""
    import torch
    import torch.nn.functional as F
    from atdn_vslam_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(5)
    pad = (taps[0] // 2, taps[1] // 2)
    hx = torch.randn(batch, h, w, 512, generator=g).half()
    h32 = torch.randn(batch * h * w, 128, generator=g)
    hx[..., :128] = h32.view(batch, h, w, 128).half()
    ok = True
    if kind == "zr":
        wt = (torch.randn(256, 512, *taps, generator=g) / (512 * 5) ** 0.5).half()
        bias = torch.randn(256, generator=g) * 0.1
        ref = F.conv2d(hx.permute(0, 3, 1, 2).float(), wt.float(), bias, padding=pad).permute(0, 2, 3, 1).reshape(-1, 256)
        z_ref = torch.sigmoid(ref[:, :128])
        rh_ref = torch.sigmoid(ref[:, 128:]) * h32
        z32 = ops.state_alloc(batch, h, w, "cuda").fill_(float("nan"))     # tiled fp32 state layout (tc_epilogue.cuh)
        rh = torch.full((batch, h, w, 128), float("nan"), dtype=torch.half, device="cuda")
        ops.conv_tc(ops.View(hx.cuda()), ops.pack_conv_weight(wt.float().cuda()), ops.pad_bias(bias.cuda()), None, cout=256,
                    taps=taps, pad=pad, bn=bn, epi=L.EPI_GRU_ZR, h32=ops.state_from_nhwc(h32.view(batch, h, w, 128).cuda()), z32=z32,
                    rh16=rh, mt=mt, flags=L.F_PAIR if pair else 0)
        torch.cuda.synchronize()
        ok &= _cmp(f"gru_zr z mt={mt} taps={taps} pair={pair} bn={bn}", ops.state_to_nhwc(z32, h, w).reshape(-1, 128), z_ref, 2e-3)
        ok &= _cmp(f"gru_zr r*h mt={mt} taps={taps} pair={pair} bn={bn}", rh.reshape(-1, 128), rh_ref, 2e-3)
    else:
        wt = (torch.randn(128, 512, *taps, generator=g) / (512 * 5) ** 0.5).half()
        bias = torch.randn(128, generator=g) * 0.1
        rhx = torch.randn(batch, h, w, 128, generator=g).half()
        z = torch.rand(batch * h * w, 128, generator=g)
        xin = torch.cat([rhx, hx[..., 128:]], -1)
        ref = F.conv2d(xin.permute(0, 3, 1, 2).float(), wt.float(), bias, padding=pad).permute(0, 2, 3, 1).reshape(-1, 128)
        h_ref = (1 - z) * h32 + z * torch.tanh(ref)
        h32d = ops.state_from_nhwc(h32.view(batch, h, w, 128).cuda())
        hxd = hx.cuda()
        out = torch.full((batch, h, w, 128), float("nan"), dtype=torch.half, device="cuda")
        ops.conv_tc(ops.View(rhx.cuda()), ops.pack_conv_weight(wt.float().cuda()), ops.pad_bias(bias.cuda()), ops.View(out), cout=128,
                    taps=taps, pad=pad, bn=qbn or (128 if mt <= 2 else 64), epi=L.EPI_GRU_Q, a2=ops.View(hxd, 128, 384), h32=h32d,
                    z32=ops.state_from_nhwc(z.view(batch, h, w, 128).cuda()), mt=mt, flags=L.F_PAIR if pair else 0)
        torch.cuda.synchronize()
        ok &= _cmp(f"gru_q h32 mt={mt} taps={taps} pair={pair}", ops.state_to_nhwc(h32d, h, w).reshape(-1, 128), h_ref, 2e-3)
        ok &= _cmp(f"gru_q h16 mt={mt} taps={taps} pair={pair}", out.reshape(-1, 128), h_ref, 2e-3)
    return ok


def bench_conv(name, cin, cout, taps, bn, mt, batch=6, h=47, w=154, reps=20, epi=None, stamps=False, pair=False):
    """Timing of one conv layer shape: legacy kernel (mt=0) or halo kernel (mt>0)."""
    import torch
    from atdn_vslam_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(0)
    pad = (taps[0] // 2, taps[1] // 2)
    cp = (cin + 7) // 8 * 8
    x = torch.randn(batch, h, w, cp, generator=g).half().cuda()
    wp = ops.pack_conv_weight((torch.randn(cout, cin, *taps, generator=g) / 50).cuda())
    bias = ops.pad_bias(torch.zeros(cout).cuda())
    out = torch.empty(batch, h, w, cout, dtype=torch.half, device="cuda")
    h32 = ops.state_from_nhwc(torch.randn(batch, h, w, 128, generator=g).cuda())
    z32 = ops.state_from_nhwc(torch.rand(batch, h, w, 128, generator=g).cuda())
    rh = torch.empty(batch, h, w, 128, dtype=torch.half, device="cuda")
    st = torch.zeros(64, dtype=torch.int64, device="cuda") if stamps else None
    fp = L.F_PAIR if pair else 0

    def run():
        if epi == "zr":
            ops.conv_tc(ops.View(x), wp, bias, None, cout=256, taps=taps, pad=pad, bn=bn, epi=L.EPI_GRU_ZR, h32=h32, z32=z32,
                        rh16=rh, mt=mt, stamps=st, flags=fp)
        elif epi == "q":
            ops.conv_tc(ops.View(rh), wp, bias, ops.View(out), cout=128, taps=taps, pad=pad, bn=bn, epi=L.EPI_GRU_Q,
                        a2=ops.View(x, 128, 384), h32=h32, z32=z32, mt=mt, stamps=st, flags=fp)
        else:
            ops.conv_tc(ops.View(x, 0, cin), wp, bias, ops.View(out), cout=cout, taps=taps, pad=pad, bn=bn, flags=L.F_RELU | fp, mt=mt,
                        stamps=st)
    for _ in range(3):
        run()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        run()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / reps * 1e3
    flops = 2.0 * batch * h * w * cout * taps[0] * taps[1] * cin
    print(f"PASS timing {name} cin={cin} cout={cout} taps={taps} batch={batch} {h}x{w} mt={mt} bn={bn} pair={pair}: {us:.1f} us  {flops / us / 1e6:.0f} TFLOP/s", flush=True)
    if stamps:
        t = st.cpu().tolist()
        print("    kernel cycles", t[1] - t[0], "mma_done", [v - t[0] for v in t[8:16] if v], "epi_start", [v - t[0] for v in t[16:24] if v],
              "epi_end", [v - t[0] for v in t[24:32] if v], flush=True)
    return True


def time_halo():
    for b in (6, 24):
        bench_conv("gru_zr", 512, 256, (1, 5), 128, 0, batch=b, epi="zr")
        bench_conv("gru_zr", 512, 256, (1, 5), 256, 1, batch=b, epi="zr", stamps=True)
        bench_conv("gru_zr", 512, 256, (1, 5), 256, 1, batch=b, epi="zr", stamps=True, pair=True)
        bench_conv("gru_zr", 512, 256, (5, 1), 256, 1, batch=b, epi="zr", pair=True)
        bench_conv("gru_zr", 512, 256, (1, 5), 128, 2, batch=b, epi="zr", pair=True)
        bench_conv("gru_q", 512, 128, (1, 5), 128, 0, batch=b, epi="q")
        bench_conv("gru_q", 512, 128, (1, 5), 128, 2, batch=b, epi="q")
        bench_conv("gru_q", 512, 128, (1, 5), 128, 2, batch=b, epi="q", pair=True, stamps=True)
        bench_conv("gru_q", 512, 128, (1, 5), 128, 1, batch=b, epi="q", pair=True)
        bench_conv("c3x3 256->192", 256, 192, (3, 3), 192, 0, batch=b)
        bench_conv("c3x3 256->192", 256, 192, (3, 3), 192, 1, batch=b)
        bench_conv("c3x3 256->192", 256, 192, (3, 3), 192, 1, batch=b, pair=True)
        bench_conv("c3x3 128->256", 128, 256, (3, 3), 128, 0, batch=b)
        bench_conv("c3x3 128->256", 128, 256, (3, 3), 256, 1, batch=b)
        bench_conv("c3x3 128->256", 128, 256, (3, 3), 256, 1, batch=b, pair=True)
        bench_conv("c3x3 256->128", 256, 128, (3, 3), 128, 0, batch=b)
        bench_conv("c3x3 256->128", 256, 128, (3, 3), 128, 2, batch=b)
        bench_conv("c3x3 256->128", 256, 128, (3, 3), 128, 2, batch=b, pair=True)
        bench_conv("c3x3 256->128", 256, 128, (3, 3), 128, 1, batch=b, pair=True)
        bench_conv("c3x3 128->64", 128, 64, (3, 3), 64, 0, batch=b)
        bench_conv("c3x3 128->64", 128, 64, (3, 3), 64, 4, batch=b)
        bench_conv("c3x3 128->64", 128, 64, (3, 3), 64, 4, batch=b, pair=True)
        bench_conv("c3x3 128->64", 128, 64, (3, 3), 64, 2, batch=b, pair=True)
        bench_conv("c1x1 324->256", 324, 256, (1, 1), 128, 0, batch=b)
        bench_conv("c1x1 324->256", 324, 256, (1, 1), 256, 1, batch=b)
        bench_conv("c1x1 324->256", 324, 256, (1, 1), 256, 1, batch=b, pair=True)
    bench_conv("enc 64->64", 64, 64, (3, 3), 64, 0, batch=7, h=188, w=616)
    bench_conv("enc 64->64", 64, 64, (3, 3), 64, 4, batch=7, h=188, w=616)
    bench_conv("enc 64->64", 64, 64, (3, 3), 64, 4, batch=7, h=188, w=616, pair=True)
    bench_conv("enc 64->64", 64, 64, (3, 3), 64, 2, batch=7, h=188, w=616, pair=True)
    bench_conv("enc 96->96", 96, 96, (3, 3), 96, 0, batch=7, h=94, w=308)
    bench_conv("enc 96->96", 96, 96, (3, 3), 96, 2, batch=7, h=94, w=308)
    bench_conv("enc 96->96", 96, 96, (3, 3), 96, 2, batch=7, h=94, w=308, pair=True)
    bench_conv("enc 128->128", 128, 128, (3, 3), 128, 0, batch=7, h=47, w=154)
    bench_conv("enc 128->128", 128, 128, (3, 3), 128, 2, batch=7, h=47, w=154, pair=True)
    return True


def time_resident():
    """Resident-weights mode of the halo kernel (all (chunk, tap) weight tiles stay in shared memory) against the
    default streamed-weights ring (resident mode is opt-in: ATDN_B_RESIDENT=1) on the layers that qualify."""
    for on in ("0", "1"):
        os.environ["ATDN_B_RESIDENT"] = on
        print(f"--- ATDN_B_RESIDENT={on}", flush=True)
        bench_conv("c1x1 324->256", 324, 256, (1, 1), 256, 1, batch=27)
        bench_conv("c1x1 324->256", 324, 256, (1, 1), 256, 1, batch=27, pair=True)
        bench_conv("c1x1 128->256", 128, 256, (1, 1), 256, 1, batch=27)
        bench_conv("c1x1 128->256", 128, 256, (1, 1), 256, 1, batch=27, pair=True)
        bench_conv("enc 64->64", 64, 64, (3, 3), 64, 4, batch=14, h=188, w=616)
        bench_conv("enc 64->64", 64, 64, (3, 3), 64, 4, batch=14, h=188, w=616, pair=True)
        bench_conv("c3x3 128->64", 128, 64, (3, 3), 64, 2, batch=27, pair=True)
    os.environ.pop("ATDN_B_RESIDENT", None)
    return True


def check_corr_full(h8=47, w8=154, batch=2, reps=0):
    """streaming corr-pyramid kernel at full size: bit-identical to the one-tile-per-CTA kernel (same MMA shape,
    same K order, same pooling arithmetic), and timing."""
    import torch
    from atdn_vslam_b200 import ops
    g = torch.Generator().manual_seed(5)
    v1 = ops.View(torch.randn(batch, h8, w8, 256, generator=g).half().cuda())
    v2 = ops.View(torch.randn(batch, h8, w8, 256, generator=g).half().cuda())
    lv = ops.alloc_pyramid(batch, h8, w8, "cuda")
    ok = True
    if batch <= 4:
        ref = ops.alloc_pyramid(batch, h8, w8, "cuda")
        ops.corr_pyramid_build(v1, v2, ref, legacy=True)
        ops.corr_pyramid_build(v1, v2, lv)
        torch.cuda.synchronize()
        for l, (t, r) in enumerate(zip(lv, ref)):
            wl = w8 >> l
            same = bool(torch.equal(t[:, :, :wl], r[:, :, :wl]))
            print(f"{'PASS' if same else 'FAIL'} corr level {l} bit-identical to the legacy kernel: {same}", flush=True)
            ok &= same
    if reps:
        for legacy in (True, False):
            ops.corr_pyramid_build(v1, v2, lv, legacy=legacy)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(reps):
                ops.corr_pyramid_build(v1, v2, lv, legacy=legacy)
            s1.record()
            torch.cuda.synchronize()
            ms = s0.elapsed_time(s1) / reps
            nbytes = sum(t.shape[0] * t.shape[1] * (w8 >> l) * 4.0 for l, t in enumerate(lv))
            print(f"TIME corr_pyramid legacy={legacy} batch={batch}: {ms:.3f} ms, {nbytes / ms / 1e6:.0f} GB/s written", flush=True)
    return ok


def check_corr(h8=16, w8=20, batch=2, pair=False, legacy=False):
    import torch
    from atdn_vslam_b200 import ops
    from oracle import gma_oracle
    g = torch.Generator().manual_seed(3)
    f1 = torch.randn(batch, 256, h8, w8, generator=g).half()
    f2 = torch.randn(batch, 256, h8, w8, generator=g).half()
    pyr = gma_oracle.corr_pyramid(f1.float(), f2.float())
    v1 = ops.View(f1.permute(0, 2, 3, 1).contiguous().cuda())
    v2 = ops.View(f2.permute(0, 2, 3, 1).contiguous().cuda())
    lv = ops.alloc_pyramid(batch, h8, w8, "cuda")
    for t in lv:
        t.fill_(float("nan"))
    ops.corr_pyramid_build(v1, v2, lv, pair=pair, legacy=legacy or pair)
    torch.cuda.synchronize()
    ok = True
    for l, (t, r) in enumerate(zip(lv, pyr)):
        hl, wl = r.shape[-2:]
        ok &= _cmp(f"corr level {l} grid {h8}x{w8} b={batch} pair={pair}", t[:, :, :wl].reshape(batch, h8 * w8, hl, wl), r, 1e-5)
    # lookup against the oracle on the oracle's pyramid layout
    coords = gma_oracle.coords_grid(batch, h8, w8) + 2.5 * torch.randn(batch, 2, h8, w8, generator=g)
    coords[0, :, 0, 0] = torch.tensor([-7.3, 2.2])
    coords[0, :, h8 - 1, w8 - 1] = torch.tensor([float(w8 - 1), float(h8 - 1)])
    ref = gma_oracle.corr_lookup(pyr, coords)
    out32 = torch.full((batch * h8 * w8, 324), float("nan"), device="cuda")
    ops.corr_lookup(lv, coords.permute(0, 2, 3, 1).contiguous().cuda(), out32=out32)
    torch.cuda.synchronize()
    got = out32.reshape(batch, h8, w8, 324).permute(0, 3, 1, 2)
    ok &= _cmp(f"lookup grid {h8}x{w8} b={batch}", got, ref, 1e-5)
    return ok


def bench_gru(pair, bn, kind="zr", batch=6, h=47, w=154, reps=20):
    """Timing of the SepConvGRU z|r (Cout 256) or q (Cout 128) 1x5 conv at the benchmark shape."""
    import torch
    from atdn_vslam_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(0)
    cout = 256 if kind == "zr" else 128
    hx = torch.randn(batch, h, w, 512, generator=g).half().cuda()
    wp = ops.pack_conv_weight((torch.randn(cout, 512, 1, 5, generator=g) / 50).cuda())
    bias = ops.pad_bias(torch.zeros(cout).cuda())
    h32 = ops.state_from_nhwc(torch.randn(batch, h, w, 128, generator=g).cuda())
    z32 = ops.state_from_nhwc(torch.rand(batch, h, w, 128, generator=g).cuda())
    rh = torch.empty(batch, h, w, 128, dtype=torch.half, device="cuda")
    out = torch.empty(batch, h, w, 128, dtype=torch.half, device="cuda")
    fl = L.F_PAIR if pair else 0

    def run():
        if kind == "zr":
            ops.conv_tc(ops.View(hx), wp, bias, None, cout=256, taps=(1, 5), pad=(0, 2), bn=bn, epi=L.EPI_GRU_ZR, flags=fl,
                        h32=h32, z32=z32, rh16=rh)
        else:
            ops.conv_tc(ops.View(rh), wp, bias, ops.View(out), cout=128, taps=(1, 5), pad=(0, 2), bn=bn, epi=L.EPI_GRU_Q,
                        flags=fl, a2=ops.View(hx, 128, 384), h32=h32, z32=z32)
    for _ in range(3):
        run()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        run()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / reps * 1e3
    flops = 2.0 * batch * h * w * cout * 5 * 512
    print(f"PASS timing gru_{kind} pair={pair} bn={bn}: {us:.1f} us/launch  {flops / us / 1e6:.0f} TFLOP/s", flush=True)
    return True


def check_attn(n=1000, batch=2, reps=0):
    """atdn_attn_probs vs fp32 torch on the same fp16 q/k: P within one fp16 rounding, 1/rowsum 1e-4."""
    import torch
    from atdn_vslam_b200 import ops
    g = torch.Generator().manual_seed(7)
    h8 = 1
    qk = (torch.randn(batch, h8, n, 256, generator=g) * 1.5).half().cuda()
    np_ = (n + 63) // 64 * 64
    p16 = torch.full((batch, n, np_), float("nan"), dtype=torch.half, device="cuda")
    inv = torch.full((batch * n,), float("nan"), dtype=torch.float32, device="cuda")
    scale = 128 ** -0.5
    ops.attn_probs(qk, p16, inv, scale)
    torch.cuda.synchronize()
    q, k = qk[:, 0, :, :128].float(), qk[:, 0, :, 128:].float()
    s = torch.matmul(q, k.transpose(1, 2)) * scale
    pref = torch.exp(s - s.max(dim=2, keepdim=True).values)
    ok = _cmp(f"attn P n={n} batch={batch}", p16[:, :, :n], pref, 1.5e-3)
    n8 = (n + 7) // 8 * 8          # TMA store clipping is 16-byte granular: columns [n, n8) are written as zeros
    ok &= bool((p16[:, :, n:n8] == 0).all()) and bool(torch.isnan(p16[:, :, n8:]).all())
    isum = 1.0 / p16[:, :, :n].float().sum(2).reshape(-1)
    ok &= _cmp("attn inv_sum", inv, isum, 1e-4)
    ok &= _cmp("attn softmax", p16[:, :, :n].float() * inv.view(batch, n, 1), torch.softmax(s, 2), 2e-3)
    # tiled layout (blocks of 32 rows x 64 columns): the same values, zeros in the pad columns of the last column block
    rb = (n + 31) // 32
    pt = torch.full((batch, rb, np_ // 64, 32, 64), float("nan"), dtype=torch.half, device="cuda")
    inv_t = torch.full_like(inv, float("nan"))
    ops.attn_probs(qk, pt, inv_t, scale, tiled=True)
    torch.cuda.synchronize()
    flat = pt.permute(0, 1, 3, 2, 4).reshape(batch, rb * 32, np_)
    same = bool(torch.equal(flat[:, :n, :n], p16[:, :, :n])) and bool((flat[:, :n, n:] == 0).all()) and bool(torch.equal(inv_t, inv))
    print(f"{'PASS' if same else 'FAIL'} attn P tiled layout == row-major, pad columns zero", flush=True)
    ok &= same
    if reps:
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            ops.attn_probs(qk, p16, inv, scale)
        s1.record()
        torch.cuda.synchronize()
        ms = s0.elapsed_time(s1) / reps
        print(f"TIME attn_probs n={n} batch={batch}: {ms:.3f} ms  ({2.0 * batch * n * n / ms / 1e6:.0f} GB/s of P written, "
              f"{2 * 2.0 * batch * n * n * 128 / ms / 1e9:.0f} TFLOP/s issued)", flush=True)
    return ok


def check_aggregate(h8=47, w8=154, batch=2, sharp=1.0, same_qk=False):
    """Attention + Aggregate in isolation (gma.py:54-76, 102-115): the product's to_qk -> attn_probs -> to_v -> P.V with the
    EPI_PV epilogue (mf + gamma * acc / rowsum) against oracle.attention / oracle.aggregate on the same fp16 inputs.
    ``sharp`` scales the context features: > 1 gives peaked (trained-like) attention rows (effective support ~5100 / 140 /
    6 of 7238 positions at 1 / 2 / 3).  ``same_qk``: the reference soft-max is taken over the product's own fp16 q / k
    (isolates softmax + P.V from the fp16 rounding of q and k, whose effect on the logits grows with their magnitude)."""
    import torch
    import gpu_e2e
    from atdn_vslam_b200 import synth
    from oracle import gma_oracle
    m, sd = gpu_e2e._gma()
    sd = {k[7:]: v for k, v in sd.items()}
    dev = torch.device("cuda")
    wts = m._weights(dev)
    plan = m._plan(batch, h8 * 8, w8 * 8, dev)
    g = torch.Generator().manual_seed(31)
    inp = (torch.relu(torch.randn(batch, 128, h8, w8, generator=g)) * sharp).half()
    mf = torch.relu(torch.randn(batch, 128, h8, w8, generator=g)).half()
    plan.hx.zero_()
    plan.hx[..., 128:256] = inp.permute(0, 2, 3, 1).cuda()
    plan.hx[..., 256:384] = mf.permute(0, 2, 3, 1).cuda()
    m._attention(plan, wts)
    m._aggregate(plan, wts)
    torch.cuda.synchronize()
    if same_qk:
        q, k = plan.qk.float().cpu().reshape(batch, h8 * w8, 256).split(128, dim=2)
        attn = torch.softmax(torch.matmul(q * 128 ** -0.5, k.transpose(1, 2)), dim=-1).unsqueeze(1)
    else:
        attn = gma_oracle.attention(inp.float(), sd)
    ref = gma_oracle.aggregate(attn, mf.float(), sd)
    got = plan.hx[..., 384:512].float().permute(0, 3, 1, 2)
    neff = float((1.0 / attn.pow(2).sum(-1)).mean())
    ok = _cmp(f"aggregate {h8}x{w8} batch={batch} (effective attention support {neff:.0f} of {h8 * w8})", got, ref, 2e-3)
    # the attention term alone (gamma * attn.v), so that a wrong P.V cannot hide behind the residual
    ok &= _cmp("aggregate minus residual", got - mf.float().cuda(), ref - mf.float(), 4e-3)
    return ok


def decode_mixed_p(p16, fmt_rb, n):
    """Tiled mixed-precision P buffer [B, ceil(N/32), Np/64, 32, 64] fp16 slots + per-sub-block format [B, ceil(N/32), Np/64]
    (1 = fp16) -> fp32 [B, N, Np] of the stored values (fp16 sub-blocks as they are; e4m3 sub-blocks = the first 2 KiB of the
    slot as [32][64] bytes) and the per-element format mask."""
    import torch
    b, rb, cb = p16.shape[:3]
    f16 = p16.float()
    raw = p16.contiguous().view(torch.uint8).view(b, rb, cb, 4096)[..., :2048].contiguous()
    f8 = raw.view(torch.float8_e4m3fn).float().view(b, rb, cb, 32, 64)
    is16 = fmt_rb.bool()[..., None, None]
    vals = torch.where(is16, f16, f8)
    flat = vals.permute(0, 1, 3, 2, 4).reshape(b, rb * 32, cb * 64)[:, :n]
    el = is16.expand(b, rb, cb, 32, 64).permute(0, 1, 3, 2, 4).reshape(b, rb * 32, cb * 64)[:, :n]
    return flat, el


def check_aggregate_mixed(h8=47, w8=154, batch=2, sharp=1.0, energy=1e-2, reps=0, small_tiles=0):
    """Mixed fp16 / e4m3 attention probabilities (atdn_attn_probs block_hot -> atdn_attn_harmonize -> ATDN_F_A_MIXED P.V on CTA
    pairs, out8 planes of v^T), each stage against torch on the SAME stored operands -- the kernels must be exact up to
    accumulation order -- and the whole aggregate against the fp32 oracle (what the rounding costs).  ``small_tiles``: the
    small-problem threshold of gma.py (0: the 128-column P.V tiles of the batched path; None: the default, 64-column tiles here)."""
    import torch
    import gpu_e2e
    from atdn_vslam_b200 import gma
    from oracle import gma_oracle
    m, sd = gpu_e2e._gma()
    sd = {k[7:]: v for k, v in sd.items()}
    dev = torch.device("cuda")
    wts = m._weights(dev)
    old = gma._P_MIXED, gma._P_HOT_ENERGY, gma._SMALL_TILES
    gma._P_MIXED, gma._P_HOT_ENERGY, gma._SMALL_TILES = True, energy, (old[2] if small_tiles is None else small_tiles)
    try:
        plan = gma._Plan(batch, h8 * 8, w8 * 8, dev)
        n, np_ = plan.n, plan.np_
        rb, cb = plan.p_hot.shape[1:]
        g = torch.Generator().manual_seed(31)
        inp = (torch.relu(torch.randn(batch, 128, h8, w8, generator=g)) * sharp).half()
        mf = torch.relu(torch.randn(batch, 128, h8, w8, generator=g)).half()
        plan.hx.zero_()
        plan.hx[..., 128:256] = inp.permute(0, 2, 3, 1).cuda()
        plan.hx[..., 256:384] = mf.permute(0, 2, 3, 1).cuda()
        plan.p16.fill_(float("nan"))
        plan.p_hot.fill_(7)
        plan.p_hot2.fill_(7)
        plan.v8.fill_(0x7f)                      # e4m3 NaN: an unwritten byte inside the K extent would poison the output
        m._attention(plan, wts)
        m._aggregate(plan, wts)
        torch.cuda.synchronize()
        ok = bool(((plan.p_hot == 0) | (plan.p_hot == 1)).all())
        if reps:
            print(f"fp16 blocks {100 * float(plan.p_hot2.float().mean()):.1f}%", flush=True)
            for name, fn in (("attn_probs mixed + harmonize", lambda: m._attention(plan, wts)), ("to_v + P.V mixed", lambda: m._aggregate(plan, wts))):
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                fn()
                s0.record()
                for _ in range(reps):
                    fn()
                s1.record()
                torch.cuda.synchronize()
                print(f"TIME {name} batch={batch}: {s0.elapsed_time(s1) / reps:.3f} ms", flush=True)
            return ok
        # 0. the pair bitmap is the OR of the eight sub-block flags of a 256-row tile; it is the storage format after harmonisation
        want2 = torch.nn.functional.pad(plan.p_hot, (0, 0, 0, (-rb) % 8)).view(batch, -1, 8, cb).amax(2)
        same = bool(torch.equal(plan.p_hot2, want2))
        fmt = plan.p_hot2.repeat_interleave(8, dim=1)[:, :rb]
        print(f"{'PASS' if same else 'FAIL'} pair bitmap = OR of its sub-blocks; hot sub-blocks {100 * float(plan.p_hot.float().mean()):.1f}%, "
              f"fp16 blocks after harmonisation {100 * float(fmt.float().mean()):.1f}%", flush=True)
        ok &= same
        stored, fmt_el = decode_mixed_p(plan.p16, fmt, n)
        hot_el = plan.p_hot.bool()[..., None, None].expand(batch, rb, cb, 32, 64).permute(0, 1, 3, 2, 4).reshape(batch, rb * 32, np_)[:, :n]
        # 1. stored values = 256 * exp(s - rowmax) of the product's own fp16 q, k: fp16 rounding in the sub-blocks attn_probs kept
        #    hot, e4m3 rounding in the others (also where harmonisation rewrote them as fp16 afterwards)
        q, k = plan.qk.float().reshape(batch, n, 256).split(128, dim=2)
        sl = torch.matmul(q, k.transpose(1, 2)) * 128 ** -0.5
        pref = torch.exp(sl - sl.max(dim=2, keepdim=True).values) * 256.0
        err = (stored[:, :, :n] - pref).abs()
        tol = torch.where(hot_el[:, :, :n], pref * 1.5e-3 + 1e-5, pref * 0.0635 + 2.0 ** -10)
        bad = int((err > tol).sum())
        print(f"{'PASS' if bad == 0 else 'FAIL'} mixed P values: {bad} of {err.numel()} outside their format's rounding", flush=True)
        ok &= bad == 0
        ok &= bool((stored[:, :, n:] == 0).all())
        # 2. the sub-block flags follow the energy criterion; the kernel's row sums are estimates from every 4th column with the row
        #    maximum as a floor (exact for flat rows, up to ~4x off for rows carried by a few keys): only sub-blocks a factor 4.6
        #    away from the threshold are held to it
        rows = pref.sum(2, keepdim=True)
        pad = rb * 32 - n
        e2 = torch.nn.functional.pad(pref, (0, np_ - n, 0, pad)).pow(2).reshape(batch, rb, 32, cb, 64).sum(4).sqrt()
        ratio = (e2 / torch.nn.functional.pad(rows, (0, 0, 0, pad), value=float("inf")).reshape(batch, rb, 32, 1)).amax(2)
        want_hot, want_cold = ratio > energy * 4.6, ratio < energy / 4.6
        wrong = int((want_hot & (plan.p_hot == 0)).sum() + (want_cold & (plan.p_hot == 1)).sum())
        print(f"{'PASS' if wrong == 0 else 'FAIL'} mixed P bitmap: {wrong} sub-blocks on the wrong side of the criterion "
              f"({int((~want_hot & ~want_cold).sum())} of {ratio.numel()} within 4.6x of it)", flush=True)
        ok &= wrong == 0
        inv = plan.inv_sum.view(batch, n)
        # (e4m3 blocks are summed pairwise in fp16 first: exact for neighbours within 2^6 of each other, else ~2^-12 per level)
        ok &= _cmp("mixed inv_sum = 1 / sum of the stored values", inv, 1.0 / stored.sum(2), 3e-4)
        # 3. e4m3 planes of v^T: hi + lo reproduces the fp32 product to ~2^-7 relative; the fp16 copy is untouched
        vt32 = torch.matmul(wts.to_v.float().view(128, 128), plan.hx[..., 256:384].float().reshape(batch, n, 128).transpose(1, 2))
        v8 = plan.v8.view(torch.float8_e4m3fn).float()
        hi, lo = v8[:, 0, :, :n], v8[:, 1, :, :n]
        ok &= _cmp("v8 hi plane = e4m3(v)", hi, vt32, 0.07)
        e = ((hi + lo) - vt32).abs()
        badv = int((e > vt32.abs() * 2.0 ** -7 + 2.0 ** -9).sum())
        print(f"{'PASS' if badv == 0 else 'FAIL'} v^T e4m3 planes: hi + lo within 2^-7 relative of the fp32 product ({badv} outliers, max abs err {float(e.max()):.2e})", flush=True)
        ok &= badv == 0
        # 4. P.V on exactly these operands: fp16 blocks x fp16 v, e4m3 blocks x (hi + lo)
        v16 = plan.vt[:, :, :n].float()
        st_n, f_n = stored[:, :, :n], fmt_el[:, :, :n]
        acc = torch.matmul(torch.where(f_n, st_n, torch.zeros_like(st_n)), v16.transpose(1, 2)) + \
            torch.matmul(torch.where(f_n, torch.zeros_like(st_n), st_n), (hi + lo).transpose(1, 2))
        gamma = float(wts.gamma)
        resid = plan.hx[..., 256:384].float().reshape(batch, n, 128)
        got = plan.hx[..., 384:512].float().reshape(batch, n, 128)
        ok &= _cmp(f"mixed P.V {h8}x{w8} batch={batch} on the stored operands", got, resid + gamma * acc * inv.unsqueeze(2), 1e-3)
        ok &= _cmp("mixed P.V minus residual", got - resid, gamma * acc * inv.unsqueeze(2), 4e-3)
        # 5. against the fp32 oracle: what the rounding costs
        attn = gma_oracle.attention(inp.float(), sd)
        ref = gma_oracle.aggregate(attn, mf.float(), sd)
        neff = float((1.0 / attn.pow(2).sum(-1)).mean())
        ok &= _cmp(f"mixed aggregate vs fp32 oracle (effective support {neff:.0f} of {n})", got.reshape(batch, h8, w8, 128).permute(0, 3, 1, 2), ref, 2e-3)
    finally:
        gma._P_MIXED, gma._P_HOT_ENERGY, gma._SMALL_TILES = old
    return ok


def check_mixed_determinism(h8=47, w8=154, batch=3, sharp=1.5, energy=2e-2, runs=4):
    """The mixed-precision attention path twice over on the same inputs: bitmaps, stored probabilities, row sums, e4m3
    planes and the aggregate must repeat bit for bit (no launch-order dependence, no uninitialised reads)."""
    import torch
    import gpu_e2e
    from atdn_vslam_b200 import gma
    m, _ = gpu_e2e._gma()
    dev = torch.device("cuda")
    wts = m._weights(dev)
    old = gma._P_MIXED, gma._P_HOT_ENERGY, gma._SMALL_TILES
    gma._P_MIXED, gma._P_HOT_ENERGY, gma._SMALL_TILES = True, energy, 0
    ok = True
    try:
        plan = gma._Plan(batch, h8 * 8, w8 * 8, dev)
        n = plan.n
        rb = plan.p_hot.shape[1]
        g = torch.Generator().manual_seed(5)
        plan.hx.zero_()
        plan.hx[..., 128:256] = (torch.relu(torch.randn(batch, h8, w8, 128, generator=g)) * sharp).half().cuda()
        plan.hx[..., 256:384] = torch.relu(torch.randn(batch, h8, w8, 128, generator=g)).half().cuda()
        snaps = []
        for r in range(runs):
            plan.p16.fill_(float(r))             # different stale bytes under every run
            plan.v8.fill_(r)
            m._attention(plan, wts)
            m._aggregate(plan, wts)
            torch.cuda.synchronize()
            fmt = plan.p_hot2.repeat_interleave(8, dim=1)[:, :rb]
            stored, _ = decode_mixed_p(plan.p16, fmt, n)
            snaps.append({"p_hot": plan.p_hot.clone(), "p_hot2": plan.p_hot2.clone(), "stored": stored, "inv_sum": plan.inv_sum.clone(),
                          "v8": plan.v8[..., :n].clone(), "vt": plan.vt[..., :n].clone(), "out": plan.hx[..., 384:512].clone()})
        for key in snaps[0]:
            same = all(torch.equal(snaps[0][key], sn[key]) for sn in snaps[1:])
            if not same:
                d = [(snaps[0][key].float() - sn[key].float()).abs() for sn in snaps[1:]]
                print(f"FAIL {key} differs between runs: max abs diff {max(float(x.max()) for x in d):.3e}, elements {[int((x > 0).sum()) for x in d]}", flush=True)
            else:
                print(f"PASS {key} repeats bit for bit over {runs} runs", flush=True)
            ok &= same
        print(f"hot sub-blocks {100 * float(plan.p_hot.float().mean()):.1f}%, fp16 blocks {100 * float(plan.p_hot2.float().mean()):.1f}%", flush=True)
    finally:
        gma._P_MIXED, gma._P_HOT_ENERGY, gma._SMALL_TILES = old
    return ok


CHECKS = {
    "mixed_determinism": lambda: check_mixed_determinism(),
    "mixed_determinism_flat": lambda: check_mixed_determinism(sharp=1.0, energy=1e-2, batch=2),
    "aggregate_mixed_full": lambda: check_aggregate_mixed(),
    "aggregate_mixed_peaked": lambda: check_aggregate_mixed(batch=1, sharp=2.0),
    "aggregate_mixed_small_odd": lambda: check_aggregate_mixed(h8=23, w8=39, batch=3, energy=2e-2),
    "aggregate_mixed_partly_hot": lambda: check_aggregate_mixed(batch=1, sharp=1.5, energy=2e-2),
    "aggregate_mixed_partly_hot_bn64": lambda: check_aggregate_mixed(batch=1, sharp=1.5, energy=2e-2, small_tiles=None),
    "time_aggregate_mixed": lambda: check_aggregate_mixed(batch=27, reps=5),
    "aggregate_full": lambda: check_aggregate(),
    "aggregate_peaked": lambda: check_aggregate(batch=1, sharp=2.0),
    "aggregate_very_peaked_same_qk": lambda: check_aggregate(batch=1, sharp=3.0, same_qk=True),
    "aggregate_small_odd": lambda: check_aggregate(h8=23, w8=39, batch=3),
    "corr_full_vs_legacy": lambda: check_corr_full(),
    "time_corr": lambda: check_corr_full(batch=27, reps=5),
    "corr_legacy_small": lambda: check_corr(legacy=True),
    "corr_legacy_odd": lambda: check_corr(h8=23, w8=39, batch=1, legacy=True),
    "corr_wide": lambda: check_corr(h8=17, w8=70, batch=3),
    "attn_small": lambda: check_attn(),
    "attn_tiny": lambda: check_attn(n=256, batch=1),
    "attn_one_tile_plus": lambda: check_attn(n=300, batch=3),
    "attn_full": lambda: check_attn(n=7238, batch=2),
    "time_attn": lambda: check_attn(n=7238, batch=27, reps=5),
    "halo_conv3x3_mt1": lambda: check_conv(cin=64, cout=128, bn=128, mt=1),
    "halo_conv3x3_mt2": lambda: check_conv(cin=128, cout=256, bn=128, mt=2, relu=True),
    "halo_conv3x3_mt4": lambda: check_conv(cin=64, cout=64, bn=64, mt=4, relu=True, resid=True, h=50, w=70),
    "halo_conv3x3_c96": lambda: check_conv(cin=96, cout=96, bn=96, mt=2),
    "halo_conv3x3_bn192": lambda: check_conv(cin=256, cout=192, bn=192, mt=1, relu=True),
    "halo_conv3x3_bn256": lambda: check_conv(cin=128, cout=256, bn=256, mt=1, relu=True),
    "halo_conv1x1_c324": lambda: check_conv(cin=324, cout=256, kh=1, kw=1, bn=128, mt=2),
    "halo_conv1x5_split": lambda: check_conv(cin=512, cout=128, kh=1, kw=5, bn=128, split=128, mt=2),
    "halo_conv5x1": lambda: check_conv(cin=128, cout=128, kh=5, kw=1, bn=64, relu=True, mt=4),
    "halo_conv_big": lambda: check_conv(cin=64, cout=64, bn=64, mt=4, h=94, w=154, batch=3),
    "halo_gru_zr": lambda: check_gru(2, "zr"),
    "small_gru_zr_mt1_bn128": lambda: check_gru(1, "zr", bn=128),
    "small_gru_zr_mt1_bn128_5x1": lambda: check_gru(1, "zr", taps=(5, 1), bn=128),
    "small_gru_q_mt1_bn64": lambda: check_gru(1, "q", qbn=64),
    "small_gru_q_mt1_bn64_5x1": lambda: check_gru(1, "q", taps=(5, 1), qbn=64),
    "small_conv3x3_mt1_bn64": lambda: check_conv(cin=256, cout=192, bn=64, mt=1, relu=True),
    "small_conv1x1_mt1_bn128": lambda: check_conv(cin=324, cout=256, kh=1, kw=1, bn=128, mt=1),
    "small_conv3x3_c64_mt1_bn64": lambda: check_conv(cin=128, cout=64, bn=64, mt=1, relu=True),
    "halo_gru_zr_5x1": lambda: check_gru(2, "zr", taps=(5, 1)),
    "halo_gru_q": lambda: check_gru(2, "q"),
    "halo_gru_q_mt4": lambda: check_gru(4, "q", taps=(5, 1)),
    "hpair_conv3x3_mt1": lambda: check_conv(cin=64, cout=128, bn=128, mt=1, pair=True),
    "hpair_conv3x3_bn256": lambda: check_conv(cin=128, cout=256, bn=256, mt=1, relu=True, pair=True),
    "hpair_conv3x3_mt4": lambda: check_conv(cin=64, cout=64, bn=64, mt=4, relu=True, resid=True, h=50, w=70, pair=True),
    "hpair_conv3x3_c96": lambda: check_conv(cin=96, cout=96, bn=96, mt=2, pair=True),
    "hpair_conv3x3_bn192": lambda: check_conv(cin=256, cout=192, bn=192, mt=1, relu=True, pair=True),
    "hpair_conv1x1_c324": lambda: check_conv(cin=324, cout=256, kh=1, kw=1, bn=256, mt=1, pair=True),
    "hpair_conv1x5_split": lambda: check_conv(cin=512, cout=128, kh=1, kw=5, bn=128, split=128, mt=2, pair=True),
    "hpair_conv_big": lambda: check_conv(cin=64, cout=64, bn=64, mt=4, h=94, w=154, batch=3, pair=True),
    "hpair_gru_zr": lambda: check_gru(1, "zr", pair=True, bn=256),
    "hpair_gru_zr_5x1": lambda: check_gru(2, "zr", taps=(5, 1), pair=True),
    "hpair_gru_q": lambda: check_gru(2, "q", pair=True),
    "halo_gru_zr_bn256": lambda: check_gru(1, "zr", bn=256),
    "time_halo": time_halo,
    "time_resident": time_resident,
    "pair_rows_basic": lambda: check_rows(pair=True),
    "pair_rows_bn256": lambda: check_rows(m=700, n=512, k=256, bn=256, pair=True),
    "pair_rows_bn64_batch": lambda: check_rows(m=130, n=70, k=128, bn=64, batch=3, pair=True),
    "pair_rows_longk": lambda: check_rows(m=200, n=128, k=7238, bn=64, pair=True),
    "pair_conv3x3": lambda: check_conv(pair=True),
    "pair_conv3x3_c96": lambda: check_conv(cin=96, cout=96, bn=96, pair=True),
    "pair_conv1x5_split_bn128": lambda: check_conv(cin=512, cout=128, kh=1, kw=5, bn=128, split=128, pair=True),
    "pair_conv5x1_bn256": lambda: check_conv(cin=512, cout=256, kh=5, kw=1, bn=256, relu=True, pair=True),
    "pair_conv3x3_bn192": lambda: check_conv(cin=256, cout=192, bn=192, relu=True, pair=True),
    "pair_conv_s2_3x3": lambda: check_conv(cin=64, cout=96, stride=2, h=22, w=40, bn=96, pair=True),
    "pair_corr_small": lambda: check_corr(pair=True),
    "pair_corr_odd": lambda: check_corr(h8=23, w8=39, batch=1, pair=True),
    "time_gru": lambda: all([bench_gru(False, 128, "zr"), bench_gru(True, 128, "zr"), bench_gru(True, 256, "zr"),
                             bench_gru(False, 128, "q"), bench_gru(True, 128, "q"), bench_gru(False, 64, "q"), bench_gru(True, 64, "q")]),
    "rows_basic": lambda: check_rows(),
    "rows_k147": lambda: check_rows(m=1000, n=64, k=147, bn=64),
    "rows_bn64_batch": lambda: check_rows(m=130, n=70, k=128, bn=64, batch=3),
    "rows_bn96": lambda: check_rows(m=256, n=96, k=192, bn=96),
    "rows_bn192": lambda: check_rows(m=256, n=576, k=256, bn=192),
    "rows_longk": lambda: check_rows(m=200, n=128, k=7238, bn=64),
    "conv3x3": lambda: check_conv(),
    "conv3x3_c96": lambda: check_conv(cin=96, cout=96, bn=96),
    "conv3x3_c324_1x1": lambda: check_conv(cin=324, cout=256, kh=1, kw=1, bn=128),
    "conv1x5_split": lambda: check_conv(cin=512, cout=128, kh=1, kw=5, bn=128, split=128),
    "conv5x1": lambda: check_conv(cin=128, cout=128, kh=5, kw=1, bn=128, relu=True),
    "conv_relu_resid": lambda: check_conv(cin=64, cout=64, relu=True, resid=True),
    "conv_s2_3x3": lambda: check_conv(cin=64, cout=96, stride=2, h=22, w=40, bn=96),
    "conv_s2_1x1": lambda: check_conv(cin=64, cout=96, kh=1, kw=1, stride=2, h=22, w=40, bn=96),
    "corr_small": lambda: check_corr(),
    "corr_odd": lambda: check_corr(h8=23, w8=39, batch=1),
}


def main():
    names = sys.argv[1:]
    if names and names[0] == "--inproc":       # all checks in ONE process (cheap; a trapped kernel fails the rest)
        import traceback
        results = {}
        for n in names[1:] or list(CHECKS):
            print(f"=== {n}", flush=True)
            try:
                results[n] = 0 if CHECKS[n]() else 1
            except Exception:
                traceback.print_exc()
                results[n] = "exception"
        print("SUMMARY", results, flush=True)
        sys.exit(0 if all(v == 0 for v in results.values()) else 1)
    if len(names) == 1 and names[0] in CHECKS and os.environ.get("ATDN_DIAG_CHILD"):
        ok = CHECKS[names[0]]()
        sys.exit(0 if ok else 1)
    names = names or list(CHECKS)
    results = {}
    for n in names:
        env = dict(os.environ, ATDN_DIAG_CHILD="1")
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), n], env=env, timeout=180,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            tail = "\n".join(p.stdout.strip().splitlines()[-14:])
            results[n] = p.returncode
            print(f"=== {n} (exit {p.returncode})\n{tail}", flush=True)
        except subprocess.TimeoutExpired:
            results[n] = "timeout"
            print(f"=== {n} TIMEOUT", flush=True)
    print("SUMMARY", results)




if __name__ == "__main__":
    main()
