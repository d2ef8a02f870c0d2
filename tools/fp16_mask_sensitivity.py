"""CPU experiment (test infrastructure): flow error when the convex-upsampling mask (576 channels, written once per
pair in fp32 today: 0.9 GB per batch of 54 pairs, read back by atdn_convex_upsample) is rounded to fp16 before the
softmax, as the reference's own fp16-autocast CUDA path does (network.py:122).  fp32 oracle, 376x1232, 12 iterations."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                   # noqa: E402

from atdn_vslam_b200 import synth              # noqa: E402
from oracle import gma_oracle as G             # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(os.cpu_count() or 8)
fr = synth.frame_sequence(2, 376, 1232, seed=11)
sd = synth.gma_state_dict()
_, base = G.raftgma_forward(sd, fr[0:1], fr[1:2], iters=12, aten_ops=True)
orig = G.upsample_flow
G.upsample_flow = lambda flow, mask: orig(flow, mask.half().float())
_, up = G.raftgma_forward(sd, fr[0:1], fr[1:2], iters=12, aten_ops=True)
G.upsample_flow = orig
epe = (up - base).pow(2).sum(1).sqrt()
print(f"fp16 upsampling mask: added EPE mean {epe.mean():.3e}  p99 {epe.flatten().quantile(0.99):.3e}  max {epe.max():.3e}")
