"""One mixed-precision attention + aggregate at 27 pairs for ncu (kernel regex tc_gemm2_kernel / attn_probs_kernel)."""
import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
os.environ.setdefault("ATDN_P_MIXED", "1")
import torch
import gpu_e2e
from atdn_vslam_b200 import gma
m, sd = gpu_e2e._gma()
dev = torch.device("cuda")
wts = m._weights(dev)
batch, h8, w8 = int(os.environ.get("NCU_BATCH", "27")), 47, 154
gma._SMALL_TILES = 0
plan = gma._Plan(batch, h8 * 8, w8 * 8, dev)
g = torch.Generator().manual_seed(31)
plan.hx.zero_()
plan.hx[..., 128:256] = torch.relu(torch.randn(batch, h8, w8, 128, generator=g)).half().cuda()
plan.hx[..., 256:384] = torch.relu(torch.randn(batch, h8, w8, 128, generator=g)).half().cuda()
for _ in range(2):
    m._attention(plan, wts)
    m._aggregate(plan, wts)
torch.cuda.synchronize()
print("hot", float(plan.p_hot.float().mean()) if plan.mixed else None)
