"""Host-side pose assembly: the caller-side contract of the hot path.

Mirrors ``atdn_vslam/utils/transforms.py:25-51`` (matrix2euler, 'yxz'), ``:54-94`` (euler2matrix),
``:97-119`` (transform) and the chaining / keyframe rule of
``atdn_vslam/slam_framework/neural_slam.py:204-215, 288-302`` -- but takes the relative poses of MANY
pairs from one device->host copy instead of ~12 implicit syncs per frame.  fp32 like the reference.
"""
from __future__ import annotations

import math

import torch


def euler2matrix(r):
    """Euler angles (yxz convention) [3] -> rotation matrix [3,3] fp32."""
    r = r.detach().to("cpu", torch.float32)
    c1, c2, c3 = torch.cos(r[0]), torch.cos(r[1]), torch.cos(r[2])
    s1, s2, s3 = torch.sin(r[0]), torch.sin(r[1]), torch.sin(r[2])
    return torch.tensor([[c1 * c3 + s1 * s2 * s3, c3 * s1 * s2 - c1 * s3, c2 * s1],
                         [c2 * s3, c2 * c3, -s2],
                         [c1 * s2 * s3 - c3 * s1, c1 * c3 * s2 + s1 * s3, c1 * c2]], dtype=torch.float32)


def matrix2euler(R):
    """Rotation matrix [3,3] -> Euler angles (yxz) [3]."""
    alpha = torch.atan2(R[0, 2], R[2, 2])
    beta = torch.atan2(-R[1, 2], torch.sqrt(1 - R[1, 2] ** 2))
    gamma = torch.atan2(R[1, 0], R[1, 1])
    return torch.tensor([alpha, beta, gamma])


def transform(rot, tr):
    """Euler vector + translation -> homogeneous [4,4] fp32 (on the host)."""
    mat = torch.eye(4, dtype=torch.float32)
    mat[:3, :3] = euler2matrix(rot)
    mat[:3, 3] = tr.detach().to("cpu", torch.float32)
    return mat


class PoseChain:
    """Sequential pose state of ``NeuralSLAM``: current pose, propagation matrix since the last
    keyframe, keyframe rule (||euler|| > 10 deg or ||t|| > 15)."""

    def __init__(self, rot_threshold_deg=10.0, translation_threshold=15.0):
        self.rotation_threshold = (rot_threshold_deg / 180) * math.pi
        self.translation_threshold = translation_threshold
        self.current_pose = torch.eye(4, dtype=torch.float32)
        self.propagation = torch.eye(4, dtype=torch.float32)
        self.poses = [self.current_pose.clone()]
        self.keyframes = [0]            # frame 0 is always registered (neural_slam.py:218-225)
        self._frame = 0

    def push(self, rot, tr):
        """Advance by one relative pose; returns True when the new frame is a keyframe."""
        m = transform(rot, tr)
        self.current_pose = self.current_pose @ m
        self.propagation = self.propagation @ m
        self._frame += 1
        rotation = matrix2euler(self.propagation[:3, :3])
        translation = self.propagation[:3, -1]
        is_key = bool(torch.norm(rotation) > self.rotation_threshold) or \
            bool(torch.norm(translation) > self.translation_threshold)
        if is_key:
            self.propagation = torch.eye(4, dtype=torch.float32)
            self.keyframes.append(self._frame)
        self.poses.append(self.current_pose.clone())
        return is_key

    def extend(self, rots, trs):
        """Whole-sequence form: ONE device->host copy of all relative poses, then the native host loop
        ``atdn_pose_chain`` (same fp32 formulas; 270 poses in microseconds instead of ~0.1 ms each through
        per-element torch calls).  Only valid on a fresh chain; ``push`` remains for incremental use."""
        rots = rots.detach().to("cpu", torch.float32).contiguous()
        trs = trs.detach().to("cpu", torch.float32).contiguous()
        if self._frame != 0:
            for t in range(rots.shape[0]):
                self.push(rots[t], trs[t])
            return torch.stack(self.poses), list(self.keyframes)
        import ctypes as C
        from . import _lib as L
        n = rots.shape[0]
        poses = torch.empty(n + 1, 4, 4, dtype=torch.float32)
        is_key = torch.empty(n + 1, dtype=torch.int32)
        cs = torch.cat([torch.cos(rots), torch.sin(rots)], 1).contiguous()     # torch's cos/sin: bit-identical to transform()
        code = L.load().atdn_pose_chain(C.c_void_p(rots.data_ptr()), C.c_void_p(cs.data_ptr()), C.c_void_p(trs.data_ptr()), C.c_int64(n),
                                        C.c_float(self.rotation_threshold), C.c_float(self.translation_threshold),
                                        C.c_void_p(poses.data_ptr()), C.c_void_p(is_key.data_ptr()))
        if code != 0:
            raise RuntimeError(f"atdn_pose_chain failed with code {code}: {L.load().atdn_last_error().decode(errors='replace')}")
        self.poses = list(poses)
        self.keyframes = torch.nonzero(is_key).flatten().tolist()
        self._frame = n
        self.current_pose = poses[-1].clone()
        # propagation since the last keyframe (so that push() can continue the chain)
        prop = torch.eye(4, dtype=torch.float32)
        last = self.keyframes[-1]
        if last < n:
            prop = torch.linalg.inv(poses[last].double()).float() @ poses[-1]
        self.propagation = prop
        return poses, list(self.keyframes)
