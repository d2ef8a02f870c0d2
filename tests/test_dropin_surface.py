"""Drop-in surface of the host layer against the LIVE reference.

Runs only where /root/reference exists (the build container); the reference is imported unmodified
(GMA wheel zip-imported, matplotlib stubbed as in tests/golden/make_golden.py).  No compute on our
side: constructors, signatures, attribute names and state-dict keys/shapes only.
"""
import inspect
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "atdn_vslam")), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    added = [os.path.join(REF, "GMA-1.0.0-py3-none-any.whl"), REF]
    for p in added:
        sys.path.insert(0, p)
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    from GMA.core.network import RAFTGMA
    from GMA.core.corr import CorrBlock
    from atdn_vslam.utils.gma_parameters import GMA_Parameters
    from atdn_vslam.odometry.network import ATDNVO
    from atdn_vslam.localization.network import MappingVAE
    from atdn_vslam.utils import transforms
    ns = types.SimpleNamespace(RAFTGMA=RAFTGMA, CorrBlock=CorrBlock, GMA_Parameters=GMA_Parameters, ATDNVO=ATDNVO,
                               MappingVAE=MappingVAE, transforms=transforms)
    yield ns
    for p in added:
        sys.path.remove(p)


def _params(fn):
    return [(n, p.default) for n, p in inspect.signature(fn).parameters.items() if n != "self"]


def test_raftgma_signature_args_and_state_dict(ref):
    from atdn_vslam_b200.gma import RAFTGMA, CorrBlock
    assert _params(RAFTGMA.__init__) == _params(ref.RAFTGMA.__init__)                 # network.py:26
    assert _params(RAFTGMA.forward) == _params(ref.RAFTGMA.forward)                   # network.py:72
    assert _params(CorrBlock.__init__) == _params(ref.CorrBlock.__init__)             # corr.py:16
    assert _params(CorrBlock.__call__) == _params(ref.CorrBlock.__call__)             # corr.py:32
    torch.manual_seed(0)
    theirs, ours = ref.RAFTGMA(ref.GMA_Parameters()), RAFTGMA(ref.GMA_Parameters())   # the reference's own args object
    a, b = theirs.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    for attr in ("corr_levels", "corr_radius", "dropout"):                            # mutated by the constructor, :33-37
        assert getattr(ours.args, attr) == getattr(theirs.args, attr)
    # DataParallel-style checkpoint (neural_slam.py:51-52) loads into ours, ours loads into the reference
    ours.load_state_dict({"module." + k: v for k, v in a.items()})
    theirs.load_state_dict(ours.state_dict())


def test_atdnvo_signature_attributes_and_state_dict(ref):
    from atdn_vslam_b200.odometry import ATDNVO
    assert _params(ATDNVO.__init__) == _params(ref.ATDNVO.__init__)                   # odometry/network.py:11
    assert _params(ATDNVO.forward) == _params(ref.ATDNVO.forward)                     # :122
    theirs, ours = ref.ATDNVO(), ATDNVO()
    a, b = theirs.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys()) and all(a[k].shape == b[k].shape for k in a)
    for attr in ("suffix", "batch_size", "device", "lstm1_h", "lstm1_c", "lstm2_h", "lstm2_c"):   # read by train/evaluate scripts
        assert hasattr(ours, attr) and hasattr(theirs, attr), attr
    assert ours.suffix == theirs.suffix and ours.batch_size == theirs.batch_size
    assert tuple(ours.lstm1_h.shape) == tuple(theirs.lstm1_h.shape)
    assert callable(ours.reset_lstm) and ours.to("cpu") is ours                       # .to() returns self (:155-160)
    ours.load_state_dict(a)
    theirs.load_state_dict(b)


def test_mapping_encoder_accepts_the_vae_checkpoint(ref):
    from atdn_vslam_b200.localization import MappingEncoder
    theirs, ours = ref.MappingVAE(), MappingEncoder()
    a, b = theirs.state_dict(), ours.state_dict()
    assert set(b.keys()) <= set(a.keys()) and all(a[k].shape == b[k].shape for k in b)
    ours.load_state_dict(a)                                                           # decoder tensors are ignored


def test_pose_helpers_equal_the_reference_bitwise(ref):
    from atdn_vslam_b200 import poses
    g = torch.Generator().manual_seed(3)
    for _ in range(20):
        rot, tr = torch.randn(3, generator=g) * 0.2, torch.randn(3, generator=g) * 3
        m = ref.transforms.transform(rot, tr)
        assert torch.equal(poses.transform(rot, tr), m)
        assert torch.equal(poses.matrix2euler(m[:3, :3]), ref.transforms.matrix2euler(m[:3, :3]))
