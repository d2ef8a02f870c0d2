#!/usr/bin/env python
"""One eager batch of the hot path with an NVTX push/pop range around every C-ABI launch (range name = the
bench.py kernel label), so that ncu can pick kernels by label:

  ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "gru_q/" --nvtx-include "attn_pv/" \
      -o gpurun_out/x python tools/ncu_batch.py [batch_pairs] [iters]
"""
import contextlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atdn_vslam_b200 import _lib as L, synth                      # noqa: E402
from atdn_vslam_b200.gma import RAFTGMA                            # noqa: E402
from atdn_vslam_b200.odometry import ATDNVO                        # noqa: E402
from atdn_vslam_b200.sequence import OdometryPipeline, preprocess  # noqa: E402


class Args:
    mixed_precision, num_heads, position_only, position_and_content = True, 1, False, False

    def __contains__(self, k):
        return hasattr(self, k)


@contextlib.contextmanager
def nvtx(label, flops, nbytes):
    torch.cuda.nvtx.range_push(label)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


def main():
    bp = int(sys.argv[1]) if len(sys.argv) > 1 else 27
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    dev = torch.device("cuda:0")
    torch.zeros(1, device=dev)
    if os.environ.get("ATDN_L2_GRAN"):      # experiment: cudaLimitMaxL2FetchGranularity (0x05) = 32 / 64 / 128
        import ctypes
        import glob
        rt = ctypes.CDLL(glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*"))[0] if glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) else "libcudart.so.12")
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["ATDN_L2_GRAN"])))
        got = ctypes.c_size_t(0)
        rt.cudaDeviceGetLimit(ctypes.byref(got), 5)
        print("cudaLimitMaxL2FetchGranularity ->", got.value, "rc", rc, flush=True)
    flow = RAFTGMA(Args())
    flow.load_state_dict(synth.gma_state_dict(module_prefix=True))
    flow = flow.to(dev).eval()
    vo = ATDNVO()
    vo.load_state_dict(synth.atdnvo_state_dict())
    vo = vo.to(dev).eval()
    pipe = OdometryPipeline(flow, vo, batch_pairs=bp, iters=iters, use_graphs=False)
    frames = preprocess(synth.frame_sequence(bp + 1, 376, 1241).to(dev))
    pipe.pair_features(frames)          # warm-up (plans, weights)
    torch.cuda.synchronize()
    L.PROFILER = nvtx
    pipe.pair_features(frames)
    torch.cuda.synchronize()
    L.PROFILER = None


if __name__ == "__main__":
    main()
