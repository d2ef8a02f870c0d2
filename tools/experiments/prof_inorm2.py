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
n, h, w, c = 28, 188, 616, 64
x = torch.randn(n, h, w, c, device="cuda").half(); y = torch.empty_like(x)
stats = torch.rand(n, c, 2, device="cuda")
mb = x.numel() * 2 / 1e6
print("copy_", timeit(lambda: y.copy_(x)) * 1e3, "us", mb, "MB")
print("clamp out", timeit(lambda: torch.clamp(x, min=0, out=y)) * 1e3, "us")
print("relu_ in-place", timeit(lambda: x.relu_()) * 1e3, "us")
print("apply", timeit(lambda: ops.inorm_apply(ops.View(x), stats, ops.View(y))) * 1e3, "us")
x2 = torch.randn(4 * n, h, w, c, device="cuda").half(); y2 = torch.empty_like(x2)
print("copy_ 4x size", timeit(lambda: y2.copy_(x2)) * 1e3 / 4, "us per 415MB")
