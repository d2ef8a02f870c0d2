import sys, torch
sys.path.insert(0, '/root/repo')
from atdn_vslam_b200 import _lib as L, synth, ops
from atdn_vslam_b200.odometry import ATDNVO
import contextlib
vo = ATDNVO(); vo.load_state_dict(synth.atdnvo_state_dict()); vo = vo.to("cuda").eval()
flows = torch.randn(27, 2, 376, 1232, device="cuda") * 10
vo.encode(flows); torch.cuda.synchronize()
ev = []
class P:
    @contextlib.contextmanager
    def __call__(self, label, f, b):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); yield; e.record(); ev.append((label, s, e))
# wrap conv32 to record shapes
orig = ops.conv32.__wrapped__
shapes = []
def conv32(x, w, bias, y, **kw):
    shapes.append((tuple(x.shape), tuple(w.shape), kw.get("stride", 1)))
    return orig(x, w, bias, y, **kw)
import functools
@functools.wraps(conv32)
def wrapped(*a, **kw):
    with L.PROFILER("conv32", 0, 0):
        return conv32(*a, **kw)
ops.conv32 = wrapped
import atdn_vslam_b200.odometry as od
L.PROFILER = P()
vo.encode(flows); torch.cuda.synchronize()
conv_ev = [e for e in ev if e[0] == "conv32"]
for (lab, s, e), sh in zip(conv_ev, shapes):
    x, w, st = sh
    oh, ow = (x[2] + 2 * (w[2] // 2) - w[2]) // st + 1, (x[3] + 2 * (w[3] // 2) - w[3]) // st + 1
    fl = 2.0 * x[0] * oh * ow * w[0] * w[1] * w[2] * w[3]
    ms = s.elapsed_time(e)
    print(f"{ms:7.3f} ms  in {x} w {w} s{st}  {fl / ms / 1e9:8.1f} GFLOP/s")
print("others:", [(l, round(s.elapsed_time(e), 3)) for l, s, e in ev if l != "conv32"])
