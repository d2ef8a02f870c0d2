import sys, torch
sys.path.insert(0, '/root/repo')
from atdn_vslam_b200 import ops
b, h8, w8 = 6, 47, 154
v1 = ops.View(torch.randn(b, h8, w8, 256).half().cuda()); v2 = ops.View(torch.randn(b, h8, w8, 256).half().cuda())
lv = ops.alloc_pyramid(b, h8, w8, "cuda")
for _ in range(3):
    ops.corr_pyramid_build(v1, v2, lv)
torch.cuda.synchronize()
