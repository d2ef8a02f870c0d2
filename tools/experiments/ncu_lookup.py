import sys, torch
sys.path.insert(0, '/root/repo')
from atdn_vslam_b200 import ops
from oracle import gma_oracle
b, h8, w8 = 8, 47, 154
lv = ops.alloc_pyramid(b, h8, w8, "cuda")
for t in lv: t.normal_()
g = torch.Generator().manual_seed(0)
coords = (gma_oracle.coords_grid(b, h8, w8) + 6.0 * torch.randn(b, 2, h8, w8, generator=g)).permute(0, 2, 3, 1).contiguous().cuda()
out = torch.empty(b, h8, w8, 328, dtype=torch.half, device="cuda")
for _ in range(3):
    ops.corr_lookup(lv, coords, out16=ops.View(out))
torch.cuda.synchronize()
