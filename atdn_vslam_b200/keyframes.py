"""On-disk keyframe store and flow pre-computation in the reference's formats.

The store is the interface between the reference's odometry, mapping and relocalisation modes and between runs
(``atdn_vslam/slam_framework/neural_slam.py:77-106, 149-153, 212-225``):

  <keyframes_path>/rgb/%06d.pth   uint8 image tensor ``[3,376,1232]`` (``torch.save(im.to("cpu").byte(), ...)``)
  <keyframes_path>/poses.pth      float32 ``[K,12]``: the first three rows of every 4x4 keyframe pose, flattened
  <keyframes_path>/MappingVAE_weights.pth   state dict of the mapping net (written by ``__create_map``, not here)

The reference blocks the frame loop on ``torch.save`` for every keyframe (``:214``); here the files are written by
a background thread from a host copy, so the GPU stream never waits on the file system, and the uint8 images
stay available in memory for relocalisation (the reference ``torch.load``s them back per query, ``:393``).

``write_flows`` is the producer of the training pipeline's pre-computed flows
(``atdn_vslam/odometry/datasets.py:113-123, 175-189``): ``flows2/<sequence>/%06d.pt``, fp16 ``[1,2,376,W]``.
"""
from __future__ import annotations

import glob
import os
import queue
import threading

import torch


class Frame:
    """``atdn_vslam/slam_framework/frame.py``: same attribute names (``rgb_file_name``, ``pose``, ``embedding``)."""

    def __init__(self, rgb_file_name, pred_pose, code=None):
        self.rgb_file_name = rgb_file_name
        self.pose = pred_pose
        self.embedding = code


def rgb_file_name(base_path, index):
    """``neural_slam.py:212-213``: zero-padded to six digits."""
    s = str(index)
    return os.path.join(base_path, "rgb", "0" * (6 - len(s)) + s + ".pth")


class KeyframeStore:
    """Keyframes of one run: ``frames`` (list of :class:`Frame`), files in the reference layout under ``path``."""

    def __init__(self, path, keep_images=True, fresh=False):
        """``fresh``: cold start like ``neural_slam.py:108-123`` -- stale ``rgb/*`` files and ``poses.pth`` of an earlier
        run under ``path`` are removed, so a run with fewer keyframes cannot pair old images with new poses."""
        self.path = path
        os.makedirs(os.path.join(path, "rgb"), exist_ok=True)
        if fresh:
            for f in glob.glob(os.path.join(path, "rgb", "*")):
                os.remove(f)
            if os.path.exists(os.path.join(path, "poses.pth")):
                os.remove(os.path.join(path, "poses.pth"))
        self.frames = []
        self.keep_images = keep_images
        self._images = []                      # uint8 host copies (None when keep_images is off)
        self._queue = queue.Queue()
        self._error = None
        self._writer = threading.Thread(target=self._drain, daemon=True)
        self._writer.start()

    # -- writer thread ----------------------------------------------------------------------------------
    def _drain(self):
        while True:
            item = self._queue.get()
            try:
                if item is None:
                    return
                name, tensor = item
                torch.save(tensor, name)
            except Exception as e:              # surfaced by flush(): a lost keyframe must not pass silently
                self._error = e
            finally:
                self._queue.task_done()

    def flush(self):
        """Block until every queued file is on disk; re-raise a writer error."""
        self._queue.join()
        if self._error is not None:
            e, self._error = self._error, None
            raise RuntimeError(f"keyframe writer failed: {e}") from e

    def close(self):
        self.flush()
        self._queue.put(None)
        self._writer.join()

    # -- registration -----------------------------------------------------------------------------------
    def add(self, image, pose, embedding=None):
        """Register ``image`` (``[3,H,W]`` or ``[1,3,H,W]``, float 0..255 or uint8, any device) with its 4x4 pose.
        The uint8 conversion truncates like ``.byte()`` in the reference.  Returns the keyframe index."""
        img = image.detach()
        img = img[0] if img.dim() == 4 else img
        host = img.to("cpu").byte().contiguous()      # ``im.to("cpu").byte()`` (:214, :224)
        index = len(self.frames)
        name = rgb_file_name(self.path, index)
        self._queue.put((name, host))
        self.frames.append(Frame(name, pose.detach().to("cpu", torch.float32).clone(), embedding))
        self._images.append(host if self.keep_images else None)
        return index

    def add_sequence(self, frames, poses, keyframe_indices):
        """Register the keyframes an ``OdometryPipeline.run`` selected: ``frames [T,3,H,W]`` (already at the SLAM
        size), ``poses [T,4,4]``, ``keyframe_indices`` (frame numbers, ascending)."""
        return [self.add(frames[int(t)], poses[int(t)]) for t in keyframe_indices]

    def image(self, index, device="cpu"):
        """uint8 keyframe image ``[3,H,W]``: from memory when kept, else from its file (after a flush)."""
        img = self._images[index]
        if img is None:
            self.flush()
            img = torch.load(self.frames[index].rgb_file_name)
        return img.to(device)

    # -- poses.pth ----------------------------------------------------------------------------------------
    def save_poses(self):
        """``end_odometry`` (:147-153): ``[K,12]`` = first three rows of every pose, flattened."""
        if not self.frames:
            raise RuntimeError("no keyframes registered")
        poses = torch.stack([f.pose.flatten()[:12] for f in self.frames], dim=0)
        self.flush()
        torch.save(poses, os.path.join(self.path, "poses.pth"))
        return poses

    @classmethod
    def load(cls, path, keep_images=False):
        """Start-up of the mapping / relocalisation modes (:77-106): poses.pth + sorted ``rgb/*``."""
        store = cls(path, keep_images=keep_images)
        homogenous = torch.tensor([0.0, 0.0, 0.0, 1.0]).view(1, 1, 4)
        poses = torch.load(os.path.join(path, "poses.pth"))
        poses = torch.cat([poses.view(len(poses), 3, 4), homogenous.repeat(len(poses), 1, 1)], dim=1)
        files = sorted(glob.glob(os.path.join(path, "rgb", "*")))
        if len(files) != len(poses):
            raise RuntimeError(f"{len(files)} keyframe images but {len(poses)} poses under {path}")
        for f, p in zip(files, poses):
            store.frames.append(Frame(f, p))
            store._images.append(torch.load(f) if keep_images else None)
        return store

    def __len__(self):
        return len(self.frames)


@torch.no_grad()
def write_flows(flow_net, frames, out_dir, batch_pairs=27, iters=12, start_index=0):
    """Pre-compute the flows the odometry training consumes (``odometry/datasets.py:113-123``): for every pair
    (t, t+1) of ``frames [T,3,376,W]`` (device tensor, W a multiple of 8) write ``out_dir/%06d.pt`` holding the fp16
    ``[1,2,376,W]`` up-sampled flow.  The flow net runs on batches of consecutive pairs (feature net once per frame);
    files are written by a background thread.  Returns the number of flows written."""
    os.makedirs(out_dir, exist_ok=True)
    q = queue.Queue(maxsize=4)
    err = []

    def drain():
        while True:
            item = q.get()
            if item is None:
                return
            first, flows = item
            try:
                for i in range(flows.shape[0]):
                    s = str(first + i)
                    torch.save(flows[i:i + 1].clone(), os.path.join(out_dir, "0" * (6 - len(s)) + s + ".pt"))
            except Exception as e:
                err.append(e)

    writer = threading.Thread(target=drain, daemon=True)
    writer.start()
    t, s, n = frames.shape[0], 0, 0
    try:
        while s < t - 1:
            e = min(t - 1, s + batch_pairs)
            _, up = flow_net.forward_frames(frames[s:e + 1], iters=iters, test_mode=True)
            q.put((start_index + s, up.half().cpu()))
            n += e - s
            s = e
    finally:                      # a failing flow net must not leave the writer blocked on the queue
        q.put(None)
        writer.join()
    if err:
        raise RuntimeError(f"flow writer failed: {err[0]}") from err[0]
    return n
