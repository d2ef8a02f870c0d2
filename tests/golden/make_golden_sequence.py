"""Golden SEQUENCE fixture: the UNMODIFIED reference ``NeuralSLAM`` (odometry mode, device=cpu) driven frame by
frame over a synthetic KITTI-shaped sequence, exactly as ``neural_slam.py:192-227`` runs it.

Run from the repo root (build container only; needs /root/reference, read-only):

    python tests/golden/make_golden_sequence.py            # writes tests/golden/sequence.npz

What is recorded, per frame pair t (frames t, t+1 of ``synth.frame_sequence(T, 376, 1241)`` -- the raw KITTI shape, so
the reference's own ``TF.resize(im, (376, 1232))`` is on the path):
  flow_lo [P,2,47,154]     coords1 - coords0 returned by the reference flow net (forward hook on ``RAFTGMA``)
  flow_up_s [P,2,47,154]   flow_up[:, :, 3::8, 5::8] (forward hook)
  rot, tr [P,3]            the stateful ``ATDNVO`` outputs (forward hook)
  poses [T,4,4]            ``NeuralSLAM.__call__`` return value after every frame (pose chain, host fp32)
  keyframes                frame indices at which ``len(slam)`` grew (``__decide_keyframe``)
  resized_sum [T]          checksum of the reference's resized frames (guards ``sequence.preprocess`` == ``TF.resize``)
plus the digests of the seeded weights.  The ATDNVO head of this fixture is scaled (``POSE_GAIN``) so that the
random-weight network moves far enough for the keyframe rule (10 degrees / 15 units) to fire several times.
"""
import os
import sys
import tempfile
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/GMA-1.0.0-py3-none-any.whl")
sys.path.insert(0, "/root/reference")
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

from atdn_vslam.slam_framework.neural_slam import NeuralSLAM      # noqa: E402
from atdn_vslam.utils.arguments import Arguments                   # noqa: E402

from atdn_vslam_b200 import synth                                  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "sequence.npz")
FRAMES = int(os.environ.get("GOLDEN_FRAMES", "21"))
torch.set_grad_enabled(False)


def main():
    gsd = synth.gma_state_dict(module_prefix=True)
    vsd = synth.atdnvo_state_dict(pose_gain=synth.SEQUENCE_POSE_GAIN)
    scratch = tempfile.mkdtemp(prefix="atdn_golden_")
    os.makedirs(os.path.join(scratch, "atdn_vslam", "checkpoints"))
    torch.save(gsd, os.path.join(scratch, "atdn_vslam", "checkpoints", "gma-kitti.pth"))   # cwd-relative, gma_parameters.py:5
    vo_path = os.path.join(scratch, "vo.pth")
    torch.save(vsd, vo_path)
    os.chdir(scratch)

    args = Arguments()
    args.device = "cpu"
    args.keyframes_path = os.path.join(scratch, "keyframes")
    slam = NeuralSLAM(args, odometry_weights=vo_path)
    slam.start_odometry()

    rec = {"flow_lo": [], "flow_up_s": [], "rot": [], "tr": [], "resized_sum": []}
    flow_net = slam._NeuralSLAM__flow_net
    vo_net = slam._NeuralSLAM__odometry_net

    def flow_hook(mod, inp, out):
        lo, up = out
        rec["flow_lo"].append(lo[0].clone().numpy())
        rec["flow_up_s"].append(up[0, :, 3::8, 5::8].clone().numpy())
        if not rec["resized_sum"]:
            rec["resized_sum"].append(float(inp[0].double().sum()))
        rec["resized_sum"].append(float(inp[1].double().sum()))

    def vo_hook(mod, inp, out):
        rec["rot"].append(out[0][0].clone().numpy())
        rec["tr"].append(out[1][0].clone().numpy())

    flow_net.module.register_forward_hook(flow_hook)
    vo_net.register_forward_hook(vo_hook)

    frames = synth.frame_sequence(FRAMES, 376, 1241, seed=synth.FRAME_SEED)
    poses, keyframes = [], []
    t0 = time.time()
    for t in range(FRAMES):
        n0 = len(slam)
        pose = slam(frames[t])
        poses.append(pose.clone().numpy())
        if len(slam) > n0:
            keyframes.append(t)
        print(f"frame {t}: {time.time() - t0:.1f} s, keyframes {keyframes}", flush=True)
    # the keyframe images the reference wrote (uint8 [3,376,1232]): checksum of each
    kf_sums = []
    for i in range(len(keyframes)):
        img = torch.load(os.path.join(args.keyframes_path, "rgb", f"{i:06d}.pth"))
        assert img.dtype == torch.uint8 and tuple(img.shape) == (3, 376, 1232), (img.dtype, img.shape)
        kf_sums.append(int(img.long().sum()))
    np.savez_compressed(
        OUT, gma_digest=np.array(synth.state_dict_digest(gsd)), vo_digest=np.array(synth.state_dict_digest(vsd)),
        frames=np.array(FRAMES), frame_seed=np.array(synth.FRAME_SEED), pose_gain=np.array(synth.SEQUENCE_POSE_GAIN),
        flow_lo=np.stack(rec["flow_lo"]), flow_up_s=np.stack(rec["flow_up_s"]), rot=np.stack(rec["rot"]), tr=np.stack(rec["tr"]),
        poses=np.stack(poses), keyframes=np.array(keyframes, dtype=np.int64), keyframe_image_sums=np.array(kf_sums, dtype=np.int64),
        resized_sum=np.array(rec["resized_sum"]))
    print("rot", np.stack(rec["rot"])[:4], "\ntr", np.stack(rec["tr"])[:4])
    print("keyframes at frames", keyframes, "->", OUT, f"{os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
