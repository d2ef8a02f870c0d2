import sys, torch
sys.path.insert(0, '/root/repo')
from atdn_vslam_b200 import ops
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(reps): fn()
    s1.record(); torch.cuda.synchronize()
    return s0.elapsed_time(s1) / reps
n = 28
for (h, w, c) in [(188, 616, 64), (94, 308, 96), (47, 154, 128)]:
    x = torch.randn(n, h, w, c, device="cuda").half()
    r = torch.randn(n, h, w, c, device="cuda").half()
    y = torch.empty_like(x)
    mb = x.numel() * 2 / 1e6
    for parts in (64, 16, 8):
        scratch = torch.empty(n * parts * c * 2, device="cuda")
        stats = torch.empty(n, c, 2, device="cuda")
        ms = timeit(lambda: ops.inorm_stats(ops.View(x), scratch, parts, stats))
        print(f"{h}x{w}x{c} stats parts={parts}: {ms*1e3:.1f} us  {mb/ms/1e3:.0f} GB/s")
    ms = timeit(lambda: ops.inorm_apply(ops.View(x), stats, ops.View(x)))
    print(f"{h}x{w}x{c} apply in-place: {ms*1e3:.1f} us  {2*mb/ms/1e3:.0f} GB/s")
    ms = timeit(lambda: ops.inorm_apply(ops.View(x), stats, ops.View(y)))
    print(f"{h}x{w}x{c} apply out-of-place: {ms*1e3:.1f} us  {2*mb/ms/1e3:.0f} GB/s")
    ms = timeit(lambda: ops.inorm_apply(ops.View(x), stats, ops.View(x), resid=ops.View(r)))
    print(f"{h}x{w}x{c} apply +resid in-place: {ms*1e3:.1f} us  {3*mb/ms/1e3:.0f} GB/s")
    ms = timeit(lambda: ops.inorm_apply(ops.View(x), stats, ops.View(y), resid=ops.View(r)))
    print(f"{h}x{w}x{c} apply +resid out-of-place: {ms*1e3:.1f} us  {3*mb/ms/1e3:.0f} GB/s")
