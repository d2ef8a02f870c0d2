"""Experiment (not a test): corr-pyramid kernel timing.  The switches are read once per process, so each variant runs in
its own process:  python tools/experiments/exp_corr.py            (runs all variants as sub-processes)
ATDN_CORR_DBG: 1 = no level 1..3 stores, 2 = no stores at all (MMA + operand feed + epilogue arithmetic only), 8 / 16 = every
level-0 strip by TMA / by 256-bit register stores (default: odd strips direct, even strips TMA); ATDN_CORR_NO_PAIR=1: cta_group::1 kernel."""
import os, subprocess, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(reps): fn()
    s1.record(); torch.cuda.synchronize()
    return s0.elapsed_time(s1) / reps


if len(sys.argv) > 1 and sys.argv[1] == "child":
    from atdn_vslam_b200 import ops
    b, h8, w8 = int(os.environ.get("CORR_BATCH", "54")), 47, 154
    v1 = ops.View(torch.randn(b, h8, w8, 256).half().cuda()); v2 = ops.View(torch.randn(b, h8, w8, 256).half().cuda())
    for half in (4, 0):
        lv = ops.alloc_pyramid(b, h8, w8, "cuda", half_levels=half)
        for alpha in ((None, 1.0) if half else (None,)):
            ms = timeit(lambda: ops.corr_pyramid_build(v1, v2, lv, alpha=alpha))
            n = h8 * w8
            print(f"pair={os.environ.get('ATDN_CORR_NO_PAIR') != '1'} dbg={os.environ.get('ATDN_CORR_DBG', '0')} half_levels={half} alpha={alpha} batch={b}: {ms:.3f} ms  "
                  f"{2.0 * b * n * n * 256 / ms / 1e9:.0f} TFLOP/s", flush=True)
else:
    for nopair, dbg in (("0", "0"), ("0", "8"), ("0", "16"), ("0", "1"), ("0", "9"), ("0", "17"), ("0", "2"), ("1", "0")):
        env = dict(os.environ, ATDN_CORR_NO_PAIR=nopair, ATDN_CORR_DBG=dbg)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
