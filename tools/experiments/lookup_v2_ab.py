"""Round-2 GPU experiment (not a test): the opt-in lookup kernel (ATDN_LOOKUP_V2=1, csrc/corr_lookup_v2.cuh) against
the shipped one on the bench shape -- fp16 outputs must be BIT-IDENTICAL, then both are timed with CUDA events.

    gpurun -- 'python tools/experiments/lookup_v2_ab.py 0 > gpurun_out/lk_v1.txt; ATDN_LOOKUP_V2=1 python tools/experiments/lookup_v2_ab.py 1 > gpurun_out/lk_v2.txt; cmp gpurun_out/lk_v1.bin gpurun_out/lk_v2.bin && echo IDENTICAL'

The switch is read once per process (static in atdn_corr_lookup), hence two processes; each writes its output tensor
to gpurun_out/lk_v{1,2}.bin and prints the time per launch and the algorithmic GB/s.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from atdn_vslam_b200 import ops            # noqa: E402
from atdn_vslam_b200.ops import View       # noqa: E402

tag = "v2" if (len(sys.argv) > 1 and sys.argv[1] == "1") else "v1"
assert (os.environ.get("ATDN_LOOKUP_V2") == "1") == (tag == "v2"), "set ATDN_LOOKUP_V2=1 for the v2 run only"
B, H8, W8 = int(os.environ.get("LK_BATCH", "27")), 47, 154
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
fm = (torch.randn(2 * B, H8, W8, 256, device=dev, generator=g) * 0.5).half()
levels = ops.alloc_pyramid(B, H8, W8, dev, half_levels=4)
ops.corr_pyramid_build(View(fm[:B]), View(fm[B:]), levels)
ys, xs = torch.meshgrid(torch.arange(H8, device=dev), torch.arange(W8, device=dev), indexing="ij")
coords = torch.stack([xs, ys], -1).float()[None].repeat(B, 1, 1, 1) + torch.randn(B, H8, W8, 2, device=dev, generator=g) * 6
coords = coords.contiguous()
out = torch.zeros(B, H8, W8, 384, dtype=torch.float16, device=dev)
ops.corr_lookup(levels, coords, out16=View(out))
torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
out[..., :324].contiguous().cpu().numpy().tofile(f"gpurun_out/lk_{tag}.bin")
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
s.record()
for _ in range(reps):
    ops.corr_lookup(levels, coords, out16=View(out))
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / reps
nbytes = B * H8 * W8 * (100 * 4 * 2 + 8 + 324 * 2)
print(f"{tag}: {ms * 1e3:.1f} us per launch of {B} pairs, {nbytes / ms / 1e6:.0f} GB/s algorithmic, checksum {out[..., :324].float().sum().item():.6e}")
