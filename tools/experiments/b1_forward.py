"""One eager batch-1 RAFTGMA.forward + ATDNVO.forward (the reference's per-frame call shape) for an ncu launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b1_launches.csv python tools/experiments/b1_forward.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import gpu_e2e                                    # noqa: E402
from atdn_vslam_b200 import synth                 # noqa: E402
from atdn_vslam_b200.odometry import ATDNVO       # noqa: E402

m, _ = gpu_e2e._gma()
m.capture_forward = False
vo = ATDNVO()
vo.load_state_dict(synth.atdnvo_state_dict())
vo = vo.to("cuda").eval()
vo.capture_forward = False
fr = synth.frame_sequence(3, 376, 1232).cuda()
for t in range(2):
    torch.cuda.nvtx.range_push(f"pair{t}")
    _, up = m(fr[t:t + 1], fr[t + 1:t + 2], iters=12, test_mode=True)
    rot, tr = vo(up)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
