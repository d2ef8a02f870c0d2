"""Generate golden fixtures by running the UNMODIFIED reference in the build container.

Run from the repo root:  python tests/golden/make_golden.py
Needs /root/reference (read-only).  The reference is imported as is: the GMA wheel is zip-imported
and ``matplotlib`` (imported by atdn_vslam/utils/helpers.py:2, unused on the path, not installed)
is stubbed.  Seeded weights come from ``atdn_vslam_b200.synth`` and are loaded into the reference
modules through their own ``load_state_dict``; a digest of the weights is stored so that the tests
can detect RNG drift.  Also asserts that ``oracle/`` reproduces the reference (prints max errors).
"""
import os
import sys
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

from GMA.core.network import RAFTGMA                      # noqa: E402
from GMA.core.corr import CorrBlock                       # noqa: E402
from atdn_vslam.utils.gma_parameters import GMA_Parameters  # noqa: E402
from atdn_vslam.odometry.network import ATDNVO            # noqa: E402
from atdn_vslam.localization.network import MappingVAE    # noqa: E402
from atdn_vslam.utils.transforms import transform as ref_transform, matrix2euler as ref_m2e  # noqa: E402

from atdn_vslam_b200 import synth                         # noqa: E402
from oracle import gma_oracle, clvo_oracle                # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_grad_enabled(False)


def err(a, b):
    return float((a - b).abs().max())


def main():
    torch.manual_seed(0)
    gsd = synth.gma_state_dict()
    gma = RAFTGMA(GMA_Parameters())
    gma.load_state_dict(gsd)
    gma.eval()

    # ---- small pair: every intermediate we test kernels against -------------------------------
    frames = synth.frame_sequence(3, 128, 160, seed=11, max_shift=3.0)
    im1, im2 = frames[0:1], frames[1:2]
    lo, up = gma(im1, im2, iters=4, test_mode=True)
    o_lo, o_up, inter = gma_oracle.raftgma_forward(gsd, im1, im2, iters=4, return_intermediates=True)
    print("small pair  flow_lo err", err(lo, o_lo), " flow_up err", err(up, o_up), " |flow| mean", float(up.abs().mean()))
    assert err(up, o_up) < 2e-3

    # reference CorrBlock on the oracle's fmaps: pyramid + one lookup at non-integer coords
    cb = CorrBlock(inter["fmap1"], inter["fmap2"], radius=4)
    g = torch.Generator().manual_seed(5)
    coords = gma_oracle.coords_grid(1, 16, 20) + 3.0 * torch.randn(1, 2, 16, 20, generator=g)
    coords[0, :, 0, 0] = torch.tensor([-7.3, 2.2])           # far outside -> zeros
    coords[0, :, 15, 19] = torch.tensor([19.0, 15.0])          # exactly on the last texel
    ref_look = cb(coords)
    pyr = gma_oracle.corr_pyramid(inter["fmap1"], inter["fmap2"])
    for l in range(4):
        assert err(cb.corr_pyramid[l].reshape(pyr[l].shape), pyr[l]) < 1e-4, l
    o_look = gma_oracle.corr_lookup(pyr, coords)
    print("lookup err", err(ref_look, o_look), " loops err", err(gma_oracle.corr_lookup_loops(pyr, coords), ref_look))
    assert err(ref_look, o_look) < 1e-4

    np.savez_compressed(
        os.path.join(OUT, "gma_small.npz"),
        digest=np.array(synth.state_dict_digest(gsd)), im1=im1.numpy().astype(np.uint8), im2=im2.numpy().astype(np.uint8),
        flow_lo=lo.numpy(), flow_up=up.numpy(), fmap1=inter["fmap1"].numpy(), fmap2=inter["fmap2"].numpy(),
        coords=coords.numpy(), lookup=ref_look.numpy(), pyr3=cb.corr_pyramid[3].numpy(),
        net0=inter["net0"].numpy(), inp=inter["inp"].numpy())

    # ---- non-test-mode list of predictions (API shape check) ------------------------------------
    preds = gma(im1, im2, iters=2, test_mode=False)
    o_preds = gma_oracle.raftgma_forward(gsd, im1, im2, iters=2, test_mode=False)
    assert len(preds) == 2 and err(preds[1], o_preds[1]) < 2e-3

    # ---- full-size pair (SLAM-internal 376x1232), iters=12: the headline parity case ------------
    frames = synth.frame_sequence(2, 376, 1232, seed=synth.FRAME_SEED)
    lo, up = gma(frames[0:1], frames[1:2], iters=12, test_mode=True)
    o_lo, o_up = gma_oracle.raftgma_forward(gsd, frames[0:1], frames[1:2], iters=12)
    print("full pair   flow_lo err", err(lo, o_lo), " flow_up err", err(up, o_up), " |flow| mean", float(up.abs().mean()))
    assert err(up, o_up) < 5e-3
    np.savez_compressed(os.path.join(OUT, "gma_full.npz"), digest=np.array(synth.state_dict_digest(gsd)),
                        flow_lo=lo.numpy(), flow_up_s4=up[:, :, ::4, ::4].numpy().astype(np.float32),
                        frame_sum=np.array([float(frames[0].double().sum()), float(frames[1].double().sum())]))

    # ---- ATDNVO: 3 consecutive calls (stateful LSTM) ---------------------------------------------
    vsd = synth.atdnvo_state_dict()
    vo = ATDNVO()
    vo.load_state_dict(vsd)
    vo.eval()
    vo.reset_lstm()
    flows = synth.synthetic_flows(3, seed=3).unsqueeze(1)
    state = clvo_oracle.zero_state()
    rots, trs, feats = [], [], []
    for t in range(3):
        r, tr = vo(flows[t])
        o_r, o_tr = clvo_oracle.atdnvo_forward(vsd, flows[t], state)
        print("atdnvo step", t, "rot err", err(r, o_r), "tr err", err(tr, o_tr), r.tolist(), tr.tolist())
        assert err(r, o_r) < 1e-5 and err(tr, o_tr) < 1e-5
        rots.append(r); trs.append(tr); feats.append(clvo_oracle.atdnvo_encode(vsd, flows[t]))
    np.savez_compressed(os.path.join(OUT, "atdnvo.npz"), digest=np.array(synth.state_dict_digest(vsd)),
                        rot=torch.cat(rots).numpy(), tr=torch.cat(trs).numpy(), feat=torch.cat(feats).numpy(),
                        flow_seed=np.array(3))
    # pose assembly
    m = ref_transform(rots[0].squeeze(), trs[0].squeeze())
    assert torch.equal(m, clvo_oracle.transform(rots[0].squeeze(), trs[0].squeeze()))
    assert torch.equal(ref_m2e(m[:3, :3]), clvo_oracle.matrix2euler(m[:3, :3]))

    # ---- MappingVAE encoder -----------------------------------------------------------------------
    esd = synth.vae_state_dict()
    vae = MappingVAE()
    missing = vae.load_state_dict(esd, strict=False)
    assert all(k.startswith("decoder.") for k in missing.missing_keys), missing
    vae.eval()
    img = synth.frame_sequence(1, 376, 1232, seed=21)
    mu = vae(img)[0]
    o_mu = clvo_oracle.vae_embed(esd, img)
    print("vae mu", tuple(mu.shape), "err", err(mu, o_mu))
    assert err(mu, o_mu) < 1e-4
    np.savez_compressed(os.path.join(OUT, "vae.npz"), digest=np.array(synth.state_dict_digest(esd)), mu=mu.numpy())
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
