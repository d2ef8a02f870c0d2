"""CPU experiment (test infrastructure): flow error when groups of convolutions take e4m3 operands (tcgen05
kind::f8f6f4 runs at twice the fp16 rate).  Inside the fp32 oracle the input and the weight of the selected
convolutions are rounded to e4m3 (activations: one scale per tensor from its max; weights: one scale per output
channel), accumulation stays fp32.  Reference: the same oracle with fp16-rounded operands everywhere (what the CUDA
path computes today), so the numbers are the error ADDED by going from fp16 to fp8 operands.

    python tools/fp8_conv_sensitivity.py [H W]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                   # noqa: E402
import torch.nn.functional as F                # noqa: E402

from atdn_vslam_b200 import synth              # noqa: E402
from oracle import gma_oracle as G             # noqa: E402

torch.set_num_threads(os.cpu_count() or 8)
torch.set_grad_enabled(False)
h, w = (376, 1232) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
fr = synth.frame_sequence(2, h, w, seed=11)
sd = synth.gma_state_dict()
E4 = torch.float8_e4m3fn


def q8_tensor(x):
    s = 384.0 / x.abs().amax().clamp_min(1e-12)
    return (x * s).to(E4).float() / s


def q8_rows(wt):
    s = 384.0 / wt.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-12)
    return (wt * s).to(E4).float() / s


def make_conv(fp8_prefixes):
    def conv(x, sdict, name, stride=1, padding=0):
        wt = sdict[name + ".weight"]
        if any(name.startswith(p) for p in fp8_prefixes):
            x, wt = q8_tensor(x), q8_rows(wt)
        else:
            x, wt = x.half().float(), wt.half().float()
        return F.conv2d(x, wt, sdict.get(name + ".bias"), stride=stride, padding=padding)
    return conv


orig = G._conv
G._conv = make_conv(())
_, base = G.raftgma_forward(sd, fr[0:1], fr[1:2], iters=12, aten_ops=True)
G._conv = orig
_, f32 = G.raftgma_forward(sd, fr[0:1], fr[1:2], iters=12, aten_ops=True)
epe = (base - f32).pow(2).sum(1).sqrt()
print(f"{h}x{w}: fp16 conv operands vs fp32 oracle: EPE mean {epe.mean():.3e} (the floor the CUDA path sits on)")
GROUPS = {
    "fnet (feature encoder)": ("fnet.",),
    "cnet (context encoder)": ("cnet.",),
    "fnet + cnet": ("fnet.", "cnet."),
    "motion encoder": ("update_block.encoder.",),
    "SepConvGRU": ("update_block.gru.",),
    "flow + mask heads": ("update_block.flow_head.", "update_block.mask."),
}
for label, prefixes in GROUPS.items():
    G._conv = make_conv(prefixes)
    _, up = G.raftgma_forward(sd, fr[0:1], fr[1:2], iters=12, aten_ops=True)
    G._conv = orig
    epe = (up - base).pow(2).sum(1).sqrt()
    print(f"  e4m3 operands in {label:24s}: added EPE mean {epe.mean():.3e}  p99 {epe.flatten().quantile(0.99):.3e}  max {epe.max():.3e}", flush=True)
