"""GPU experiment (not a test): time the correlation lookup on the bench shape (batch 54, 47 x 154) and print the
algorithmic GB/s (4 levels x 10 x 10 fp16 texels + coords read, 324 fp16 written per query)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from atdn_vslam_b200 import ops            # noqa: E402
from atdn_vslam_b200.ops import View       # noqa: E402

B, H8, W8 = int(os.environ.get("LK_BATCH", "54")), 47, 154
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
fm = (torch.randn(2 * B, H8, W8, 256, device=dev, generator=g) * 0.5).half()
levels = ops.alloc_pyramid(B, H8, W8, dev, half_levels=4)
ops.corr_pyramid_build(View(fm[:B]), View(fm[B:]), levels)
ys, xs = torch.meshgrid(torch.arange(H8, device=dev), torch.arange(W8, device=dev), indexing="ij")
for spread in (1.5, 6.0):
    coords = torch.stack([xs, ys], -1).float()[None].repeat(B, 1, 1, 1) + torch.randn(B, H8, W8, 2, device=dev, generator=g) * spread
    coords = coords.contiguous()
    out = torch.zeros(B, H8, W8, 328, dtype=torch.float16, device=dev)
    ops.corr_lookup(levels, coords, out16=View(out))
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    s.record()
    for _ in range(reps):
        ops.corr_lookup(levels, coords, out16=View(out))
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    nbytes = B * H8 * W8 * (100 * 4 * 2 + 8 + 324 * 2)
    print(f"lookup (coords spread {spread} texels): {ms * 1e3:.1f} us per launch of {B} pairs, {nbytes / ms / 1e6:.0f} GB/s algorithmic, "
          f"checksum {out[..., :324].float().sum().item():.6e}", flush=True)
