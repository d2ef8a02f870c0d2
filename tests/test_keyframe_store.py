"""Keyframe store / flow files in the reference's on-disk formats (neural_slam.py:77-106, 147-153, 212-225;
odometry/datasets.py:113-123).  Host logic only: runs without a GPU."""
import glob
import os

import pytest
import torch

from atdn_vslam_b200.keyframes import Frame, KeyframeStore, rgb_file_name


def _poses(k):
    out = []
    for i in range(k):
        p = torch.eye(4)
        p[:3, 3] = torch.tensor([0.5 * i, -0.1 * i, 15.5 * i])
        p[0, 1], p[1, 0] = 0.01 * i, -0.01 * i
        out.append(p)
    return out


def test_file_names_follow_the_reference():
    assert rgb_file_name("/x", 0).endswith("rgb/000000.pth")
    assert rgb_file_name("/x", 123).endswith("rgb/000123.pth")
    assert rgb_file_name("/x", 1234567).endswith("rgb/1234567.pth")     # '0' * negative = '' in the reference too


def test_store_round_trip_in_reference_format(tmp_path):
    g = torch.Generator().manual_seed(0)
    imgs = [torch.rand(3, 16, 24, generator=g) * 255 for _ in range(5)]
    poses = _poses(5)
    store = KeyframeStore(str(tmp_path))
    assert store.add(imgs[0], poses[0]) == 0
    assert store.add(imgs[1].unsqueeze(0), poses[1]) == 1                # [1,3,H,W] like im2 after the padder
    for i in range(2, 5):
        store.add(imgs[i], poses[i])
    saved = store.save_poses()
    store.close()
    # what the reference's start-up code does (neural_slam.py:77-84, 93-100)
    homogenous = torch.tensor([0.0, 0.0, 0.0, 1.0]).view(1, 1, 4)
    p = torch.load(os.path.join(str(tmp_path), "poses.pth"))
    assert p.shape == (5, 12) and torch.equal(p, saved)
    p = torch.cat([p.view(len(p), 3, 4), homogenous.repeat(len(p), 1, 1)], dim=1)
    files = sorted(glob.glob(os.path.join(str(tmp_path), "rgb", "*")))
    assert [os.path.basename(f) for f in files] == [f"{i:06d}.pth" for i in range(5)]
    for i, f in enumerate(files):
        rgb = torch.load(f)
        assert rgb.dtype == torch.uint8 and rgb.shape == (3, 16, 24)
        assert torch.equal(rgb, imgs[i].byte())                          # .byte() truncation, as in the reference
        assert torch.equal(p[i], poses[i])
    # our loader restores the same frames
    again = KeyframeStore.load(str(tmp_path), keep_images=True)
    assert len(again) == 5
    for i in range(5):
        assert isinstance(again.frames[i], Frame) and again.frames[i].embedding is None
        assert torch.equal(again.frames[i].pose, poses[i])
        assert torch.equal(again.image(i), imgs[i].byte())
    again.close()


def test_add_sequence_and_lazy_images(tmp_path):
    frames = (torch.arange(6 * 3 * 8 * 8) % 256).float().view(6, 3, 8, 8)
    poses = torch.stack(_poses(6))
    store = KeyframeStore(str(tmp_path), keep_images=False)
    assert store.add_sequence(frames, poses, [0, 2, 5]) == [0, 1, 2]
    assert torch.equal(store.image(1), frames[2].byte())                 # read back from disk after a flush
    assert torch.equal(store.frames[2].pose, poses[5])
    store.close()


def test_mismatched_store_is_rejected(tmp_path):
    store = KeyframeStore(str(tmp_path))
    store.add(torch.zeros(3, 4, 4), torch.eye(4))
    store.add(torch.zeros(3, 4, 4), torch.eye(4))
    store.save_poses()
    store.close()
    os.remove(rgb_file_name(str(tmp_path), 1))
    with pytest.raises(RuntimeError):
        KeyframeStore.load(str(tmp_path))
    with pytest.raises(RuntimeError):
        KeyframeStore(str(tmp_path / "empty")).save_poses()


def test_fresh_store_removes_stale_files(tmp_path):
    """Cold start (neural_slam.py:108-123): a second run with fewer keyframes must not see the first run's files."""
    import torch
    from atdn_vslam_b200.keyframes import KeyframeStore
    a = KeyframeStore(str(tmp_path))
    for i in range(3):
        a.add(torch.full((3, 8, 8), float(i)), torch.eye(4))
    a.save_poses()
    a.close()
    b = KeyframeStore(str(tmp_path), fresh=True)
    b.add(torch.zeros(3, 8, 8), torch.eye(4))
    b.save_poses()
    b.close()
    assert len(KeyframeStore.load(str(tmp_path))) == 1
