"""GPU experiment: cost of ATDNVO.recurrent_scan (persistent LSTM scan + batched input / head products) vs sequence length."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from atdn_vslam_b200 import synth
from atdn_vslam_b200.odometry import ATDNVO
vo = ATDNVO(); vo.load_state_dict(synth.atdnvo_state_dict()); vo = vo.to("cuda").eval()
for t in (54, 270, 432, 1080, 4540):
    f = torch.randn(t, 512, device="cuda") * 0.5
    vo.reset_lstm(); vo.recurrent_scan(f); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        vo.recurrent_scan(f)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(f"recurrent_scan T={t}: {ms:.3f} ms = {1e3 * ms / t:.2f} us per pair", flush=True)
