"""CPU experiment (test infrastructure): can the CLVO pose encoder's convolutions (fp32 CUDA cores today: 4% of the
step at 39% of the FP32 FMA peak) run on tensor cores?  Inside the fp32 oracle the operands of every conv2d of the
encoder are rounded to TF32 (10-bit mantissa, kind::tf32), to fp16, or split into TF32 hi + lo terms with the three
significant products kept (3xTF32); the relative error of the pose (rot, tr) of three consecutive stateful steps is
compared with the 1e-4 tolerance of the north star.

    python tools/clvo_operand_sensitivity.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                   # noqa: E402
import torch.nn.functional as F                # noqa: E402

from atdn_vslam_b200 import synth              # noqa: E402
from oracle import clvo_oracle as O            # noqa: E402

torch.set_num_threads(os.cpu_count() or 8)
torch.set_grad_enabled(False)
sd = synth.atdnvo_state_dict()
flows = synth.synthetic_flows(3, seed=5)
orig_conv2d = F.conv2d


def tf32(x):                                   # round to nearest even on the 13 dropped mantissa bits
    i = x.contiguous().view(torch.int32)
    r = i + 0x0FFF + ((i >> 13) & 1)
    return (r & ~0x1FFF).view(torch.float32)


def run(mode):
    def conv2d(x, w, b=None, **kw):
        if mode == "tf32":
            return orig_conv2d(tf32(x), tf32(w), b, **kw)
        if mode == "fp16":
            return orig_conv2d(x.half().float(), w.half().float(), b, **kw)
        if mode == "3xtf32":
            xh, wh = tf32(x), tf32(w)
            xl, wl = tf32(x - xh), tf32(w - wh)
            return orig_conv2d(xh, wh, b, **kw) + orig_conv2d(xl, wh, None, **kw) + orig_conv2d(xh, wl, None, **kw)
        return orig_conv2d(x, w, b, **kw)
    F.conv2d = conv2d
    try:
        state = O.zero_state()
        out = []
        for t in range(flows.shape[0]):
            out.append(O.atdnvo_forward(sd, flows[t:t + 1], state))
    finally:
        F.conv2d = orig_conv2d
    return out


ref = run("fp32")
for mode in ("tf32", "fp16", "3xtf32"):
    got = run(mode)
    rr = max(((g[0] - r[0]).norm() / r[0].norm()).item() for g, r in zip(got, ref))
    rt = max(((g[1] - r[1]).norm() / r[1].norm()).item() for g, r in zip(got, ref))
    print(f"{mode:7s} operands in the encoder convs: max relative error over 3 steps  rot {rr:.2e}  tr {rt:.2e}   (tolerance 1e-4)", flush=True)
