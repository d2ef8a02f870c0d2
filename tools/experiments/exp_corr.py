"""Experiment (not a test): where does the streaming corr-pyramid kernel spend its time?  ATDN_CORR_DBG switches:
1 = no level 1..3 stores, 2 = no stores at all (MMA + operand feed + epilogue arithmetic only)."""
import os, sys, torch
sys.path.insert(0, '/root/repo')
from atdn_vslam_b200 import ops
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(reps): fn()
    s1.record(); torch.cuda.synchronize()
    return s0.elapsed_time(s1) / reps
x = torch.empty(2 * 1024**3, dtype=torch.float32, device="cuda")   # 8 GB
ms = timeit(lambda: x.fill_(1.0)); print(f"fill_ 8GB: {ms:.3f} ms {x.numel()*4/ms/1e6:.0f} GB/s write-only")
y = torch.empty_like(x)
ms = timeit(lambda: y.copy_(x)); print(f"copy 8GB: {ms:.3f} ms {2*x.numel()*4/ms/1e6:.0f} GB/s r+w")
del x, y
b, h8, w8 = 27, 47, 154
v1 = ops.View(torch.randn(b, h8, w8, 256).half().cuda()); v2 = ops.View(torch.randn(b, h8, w8, 256).half().cuda())
for half in (4, 0):
    lv = ops.alloc_pyramid(b, h8, w8, "cuda", half_levels=half)
    for dbg in (0, 1, 2):
        os.environ["ATDN_CORR_DBG"] = str(dbg)
        ms = timeit(lambda: ops.corr_pyramid_build(v1, v2, lv))
        print(f"corr half_levels={half} dbg={dbg}: {ms:.3f} ms")
os.environ.pop("ATDN_CORR_DBG")
