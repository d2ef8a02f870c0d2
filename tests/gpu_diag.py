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


def check_conv(cin=64, cout=64, kh=3, kw=3, stride=1, h=20, w=37, batch=2, bn=64, relu=False, resid=False, split=0, pair=False):
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
                    a2=a2, resid=rs)
    else:
        ops.conv_tc(ops.View(xn, 0, cin), wp, bp, ops.View(out), cout=cout, taps=(kh, kw), pad=(ph, pw), stride=stride,
                    bn=bn, flags=flags, resid=rs)
    torch.cuda.synchronize()
    got = out[..., :cout].permute(0, 3, 1, 2)
    return _cmp(f"conv cin={cin} cout={cout} k={kh}x{kw} s={stride} {h}x{w} b={batch} bn={bn} relu={relu} resid={resid} split={split} pair={pair}",
                got, ref, 2e-3)


def check_corr(h8=16, w8=20, batch=2, pair=False):
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
    ops.corr_pyramid_build(v1, v2, lv, pair=pair)
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
    h32 = torch.randn(batch * h * w, 128, generator=g).cuda()
    z32 = torch.rand(batch * h * w, 128, generator=g).cuda()
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


CHECKS = {
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




def experiment_pipeline():
    """Where does a K step go?  zr-conv shape, single-CTA kernel, with loads and/or MMAs disabled."""
    import torch
    from atdn_vslam_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(0)
    for batch in (1, 2, 3, 6, 12):
        h, w = 47, 154
        hx = torch.randn(batch, h, w, 512, generator=g).half().cuda()
        hxc = torch.randn(batch, h, w, 64, generator=g).half().cuda()
        wp = ops.pack_conv_weight((torch.randn(256, 512, 1, 5, generator=g) / 50).cuda())
        wpc = ops.pack_conv_weight((torch.randn(256, 64, 1, 5, generator=g) / 50).cuda())
        bias = ops.pad_bias(torch.zeros(256).cuda())
        out = torch.empty(batch, h, w, 256, dtype=torch.half, device="cuda")
        for name, fl in (("full", 0), ("no_mma", 128), ("no_tma", 256), ("neither", 384)):
            def run():
                ops.conv_tc(ops.View(hx), wp, bias, ops.View(out), cout=256, taps=(1, 5), pad=(0, 2), bn=128, flags=fl)
            for _ in range(3):
                run()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            for _ in range(10):
                run()
            e.record()
            torch.cuda.synchronize()
            print(f"PASS exp batch={batch} ctas={batch * 120} {name}: {s.elapsed_time(e) / 10 * 1e3:.1f} us", flush=True)
    return True


def experiment_stamps():
    import torch
    from atdn_vslam_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(0)
    batch, h, w = 6, 47, 154
    hx = torch.randn(batch, h, w, 512, generator=g).half().cuda()
    wp = ops.pack_conv_weight((torch.randn(256, 512, 1, 5, generator=g) / 50).cuda())
    bias = ops.pad_bias(torch.zeros(256).cuda())
    out = torch.empty(batch, h, w, 256, dtype=torch.half, device="cuda")
    dbuf = torch.zeros(128, dtype=torch.int64, device="cuda")
    for name, fl in (("full", 0), ("no_mma", 128), ("no_tma", 256), ("neither", 384)):
        for rep in range(3):
            d = L.TcDesc()
            a = ops.View(hx)
            d.bn, d.epi, d.flags, d.a_mode, d.b_mode = 128, L.EPI_STORE16, fl | 512, L.MODE_PATCH, L.MODE_ROWS
            d.out_h, d.out_w, d.taps_h, d.taps_w, d.pad_h, d.pad_w, d.stride = h, w, 1, 5, 0, 2, 1
            d.a = a.ptr(); L._set(d.a_dims, (512, w, h, batch)); L._set(d.a_strides, (512, w * 512, h * w * 512))
            ops._fill_weight(d, wp); ops._fill_out(d, ops.View(out), 256, 1.0, bias)
            d.lvl[2] = dbuf.data_ptr()
            L.tc_gemm(d)
            torch.cuda.synchronize()
        t = dbuf.cpu().tolist()
        print("    prod:", [t[8 + i] - t[0] for i in range(0, 40, 3)], flush=True)
        print("    cons:", [t[64 + i] - t[0] for i in range(0, 40, 3)], flush=True)
        print(f"PASS stamps {name}: setup->prod_end {t[1]-t[0]} mma_end {t[2]-t[0]} epi_start {t[3]-t[0]} epi_end {t[4]-t[0]} exit {t[5]-t[0]} cycles", flush=True)
    return True


CHECKS["exp_stamps"] = experiment_stamps
CHECKS["exp_pipeline"] = experiment_pipeline
if __name__ == "__main__":
    main()
