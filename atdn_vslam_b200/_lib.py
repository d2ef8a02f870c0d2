"""ctypes binding of libatdn_b200.so (include/atdn_b200.h).  Fails loudly: a missing library, a
missing symbol or a non-zero return code raises -- there is no PyTorch / CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libatdn_b200.so")

MODE_ROWS, MODE_PATCH = 0, 1
EPI_STORE16, EPI_STORE32, EPI_CORR, EPI_GRU_ZR, EPI_GRU_Q, EPI_PV, EPI_FLOW = range(7)
F_RELU, F_RESID, F_FLOWTAIL, F_TANH_LO, F_B_BATCHED, F_A_SHARED, F_PAIR, F_STATS, F_TILED32, F_PRE16, F_Z16, F_H16 = 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048
F_A_TILED = 4096
F_A_MIXED = 8192

EXPORTS = (
    "atdn_last_error", "atdn_version", "atdn_check_device", "atdn_tc_gemm", "atdn_corr_lookup",
    "atdn_stem_pack", "atdn_flow_pack", "atdn_inorm_stats", "atdn_inorm_apply",
    "atdn_convex_upsample", "atdn_coords_init", "atdn_conv32", "atdn_linear32",
    "atdn_lstm_cell", "atdn_keyframe_search", "atdn_attn_probs", "atdn_corr_pyramid", "atdn_clvo_lstm_scan", "atdn_pose_chain", "atdn_flow_head_gather", "atdn_inorm_finalize", "atdn_attn_harmonize", "atdn_resize_aa",
)


class TcDesc(C.Structure):
    _fields_ = [
        ("bn", C.c_int32), ("epi", C.c_int32), ("flags", C.c_int32), ("a_mode", C.c_int32), ("b_mode", C.c_int32),
        ("n_valid", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
        ("taps_h", C.c_int32), ("taps_w", C.c_int32), ("pad_h", C.c_int32), ("pad_w", C.c_int32), ("stride", C.c_int32),
        ("a_split_chunk", C.c_int32),
        ("a", C.c_void_p), ("a_dims", C.c_int64 * 4), ("a_strides", C.c_int64 * 3),
        ("a2", C.c_void_p), ("a2_dims", C.c_int64 * 4), ("a2_strides", C.c_int64 * 3),
        ("b", C.c_void_p), ("b_dims", C.c_int64 * 4), ("b_strides", C.c_int64 * 3),
        ("alpha", C.c_float), ("bias", C.c_void_p), ("out", C.c_void_p), ("out_pitch", C.c_int64),
        ("out_ch_off", C.c_int64), ("resid16", C.c_void_p), ("resid_pitch", C.c_int64), ("resid_ch_off", C.c_int64),
        ("h32", C.c_void_p), ("z32", C.c_void_p), ("rh16", C.c_void_p), ("aux32", C.c_void_p), ("gamma", C.c_void_p),
        ("lvl", C.c_void_p * 3), ("lvl_pitch", C.c_int32 * 4), ("corr_h", C.c_int32), ("corr_w", C.c_int32),
        ("mt", C.c_int32),
        ("out8", C.c_void_p), ("b8", C.c_void_p), ("a_hot", C.c_void_p),
    ]


class Conv32Desc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p),
        ("in_scale", C.c_void_p), ("in_shift", C.c_void_p),
        ("bn_scale", C.c_void_p), ("bn_shift", C.c_void_p), ("skip", C.c_void_p),
        ("bn2_scale", C.c_void_p), ("bn2_shift", C.c_void_p),
        ("batch", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32),
        ("k", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("mish", C.c_int32),
        ("x_pitch", C.c_int32), ("y_pitch", C.c_int32),
    ]


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m atdn_vslam_b200.build` "
            "(atdn_vslam_b200 has no PyTorch/CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise RuntimeError(f"{LIB_PATH} does not export {name}")
    lib.atdn_last_error.restype = C.c_char_p
    for name in EXPORTS[1:]:
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


PV_HOT_FRACTION = 0.0   # fp16 share of the mixed-precision P blocks of the last profiled attention (gma._attention)
LAUNCHES = 0          # kernels launched through the C ABI by this process (bench.py reports it)
PROFILER = None       # optional callable(label, flops, bytes) -> context manager, installed by bench.py


def check(code, what, launches=1):
    global LAUNCHES
    LAUNCHES += launches
    if code != 0:
        raise RuntimeError(f"{what} failed with code {code}: {load().atdn_last_error().decode(errors='replace')}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t, offset_elems=0):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr() + offset_elems * t.element_size())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("atdn_vslam_b200 kernels need CUDA tensors on an sm_100 device (no CPU fallback)")


def tc_label(d: TcDesc):
    """(label, flops, bytes) of one tensor-core launch, derived from its descriptor (profiling only).  Bytes are
    only given where HBM is the binding roofline: the P.V aggregation re-reads the fp16 probabilities every
    iteration (N x Np x 2 B per pair)."""
    nbytes = 0.0
    if d.a_mode == MODE_PATCH:
        cin = d.a_dims[0] + (d.a2_dims[0] if d.a_split_chunk else 0)
        flops = 2.0 * d.out_h * d.out_w * d.a_dims[3] * d.n_valid * d.taps_h * d.taps_w * cin   # algorithmic (n_valid, not the padded tile)
        name = {EPI_GRU_ZR: "gru_zr", EPI_GRU_Q: "gru_q", EPI_FLOW: "flow_head2"}.get(d.epi, f"conv{d.taps_h}x{d.taps_w}_{cin}to{d.n_valid}_s{d.stride}")
        if d.flags & F_TILED32:
            name = "gru_context_pre"
    else:
        batch = d.b_dims[3] if (d.flags & F_A_SHARED) else d.a_dims[3]
        flops = 2.0 * d.a_dims[1] * d.n_valid * min(d.a_dims[0], d.b_dims[0]) * batch
        if d.epi == EPI_PV:   # P [rows, pitch] + V^T [128, pitch] read, residual read + output written (fp16)
            p_bytes = 2.0 if not (d.flags & F_A_MIXED) else 1.0 + PV_HOT_FRACTION    # mixed: e4m3 blocks are 1 byte per probability
            nbytes = batch * (p_bytes * d.a_dims[1] * d.a_strides[0] + 2.0 * d.n_valid * d.b_strides[0] + 4.0 * d.a_dims[1] * d.n_valid)
        name = {EPI_CORR: "corr_gemm", EPI_PV: "attn_pv", EPI_STORE32: "attn_qk"}.get(
            d.epi, "to_v" if (d.flags & F_A_SHARED) else f"rows_k{d.a_dims[0]}to{d.n_valid}")
    return name, flops, nbytes


def tc_gemm(desc: TcDesc):
    if PROFILER is not None:
        name, flops, nbytes = tc_label(desc)
        with PROFILER(name, flops, nbytes):
            check(load().atdn_tc_gemm(C.byref(desc), stream_ptr()), "atdn_tc_gemm")
        return
    check(load().atdn_tc_gemm(C.byref(desc), stream_ptr()), "atdn_tc_gemm")


def _set(arr, values):
    for i, v in enumerate(values):
        arr[i] = int(v)
