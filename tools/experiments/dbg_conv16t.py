import sys, time, torch
sys.path.insert(0, '/root/repo')
import torch.nn.functional as F
from atdn_vslam_b200 import ops
from atdn_vslam_b200.odometry import new_map
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator().manual_seed(0)
for (cin, k, s, pad, h, w, affine) in [(16, 3, 1, 1, 40, 156, False), (16, 3, 2, 1, 40, 156, False), (2, 7, 2, 3, 64, 160, True),
                                       (16, 3, 1, 1, 47, 154, False), (16, 3, 2, 1, 47, 154, False), (16, 3, 1, 1, 12, 39, False), (16, 3, 2, 1, 24, 77, False),
                                       (16, 3, 1, 1, 188, 616, False), (16, 3, 2, 1, 188, 616, False), (2, 7, 2, 3, 376, 1232, True)]:
    b = 3
    x = new_map(b, cin, h, w, "cuda", True)
    x.copy_(torch.randn(b, cin, h, w, generator=g).cuda())
    wt = (torch.randn(16, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    bias = torch.randn(16, generator=g).cuda()
    sc = (torch.rand(cin, generator=g) + 0.5).cuda() if affine else None
    sh = torch.randn(cin, generator=g).cuda() if affine else None
    oh, ow = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    y = new_map(b, 16, oh, ow, "cuda", True)
    y.fill_(float("nan"))
    t0 = time.time()
    try:
        bs, bh = (torch.rand(16, generator=g) + 0.5).cuda(), torch.randn(16, generator=g).cuda()
        ops.conv32(x, wt, bias, y, stride=s, pad=pad, mish=True, in_scale=sc, in_shift=sh, bn_scale=bs, bn_shift=bh)
        torch.cuda.synchronize()
    except Exception as e:
        print("FAIL", cin, k, s, h, w, str(e)[:100], f"{time.time()-t0:.2f}s")
        break
    xin = x if not affine else x * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    ref = F.conv2d(xin.contiguous().double(), wt.double(), bias.double(), stride=s, padding=pad)
    ref = (F.mish(ref) * bs.double().view(1, -1, 1, 1) + bh.double().view(1, -1, 1, 1)).float()
    err = (y - ref).abs().max().item()
    print(f"cin={cin} k={k} s={s} {h}x{w}: max err {err:.3e} (ref max {ref.abs().max().item():.2f}) {time.time()-t0:.2f}s", flush=True)
