"""CPU experiment (test infrastructure): how much does storing the correlation pyramid in fp16 move the final
flow?  Runs the fp32 oracle on a 376x1232 synthetic pair with (a) fp32 pyramid, (b) every pyramid level rounded
to fp16, (c) lookup output rounded to fp16 (what the CUDA path already does before convc1), (d) both."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from atdn_vslam_b200 import synth
from oracle import gma_oracle as G

torch.set_num_threads(8)
sd = synth.gma_state_dict()
h, w = (376, 1232) if len(sys.argv) < 2 else (int(sys.argv[1]), int(sys.argv[2]))
fr = synth.frame_sequence(2, h, w, seed=11)
orig_pyr, orig_lookup = G.corr_pyramid, G.corr_lookup

def run(round_pyr, round_out):
    G.corr_pyramid = (lambda a, b: [l.half().float() for l in orig_pyr(a, b)]) if round_pyr else orig_pyr
    G.corr_lookup = (lambda p, c: orig_lookup(p, c).half().float()) if round_out else orig_lookup
    t = time.time()
    lo, up = G.raftgma_forward(sd, fr[0:1], fr[1:2], iters=12)
    G.corr_pyramid, G.corr_lookup = orig_pyr, orig_lookup
    return up, time.time() - t

base, dt = run(False, False)
print(f"baseline fp32: {dt:.1f}s |flow| mean {base.abs().mean():.2f}")
for name, rp, ro in (("pyramid fp16", True, False), ("lookup out fp16", False, True), ("both", True, True)):
    up, _ = run(rp, ro)
    epe = (up - base).pow(2).sum(1).sqrt()
    print(f"{name}: EPE mean {epe.mean():.3e} p99 {epe.flatten().quantile(0.99):.3e} max {epe.max():.3e}", flush=True)
