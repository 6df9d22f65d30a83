"""ctypes binding of libopental_b200.so (the C ABI declared in include/opental_b200.h).

The library is the product: there is no Python/CPU fallback.  Importing this module never fails (so that
CPU-only tooling can inspect the package), but any attempt to *call* a kernel without the built library or
without a CUDA device raises `RuntimeError`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libopental_b200.so")

OTAL_OK = 0
ERR_NAMES = {-1: "OTAL_ERR_BAD_ARG", -2: "OTAL_ERR_CUDA", -3: "OTAL_ERR_DRIVER", -4: "OTAL_ERR_UNSUPPORTED"}


def _ints(*names):
    return [(n, c_int) for n in names]


def _ptrs(*names):
    return [(n, c_void_p) for n in names]


class ConvDesc(Structure):
    """Mirror of `otal_conv_desc` (include/opental_b200.h)."""

    _fields_ = (_ints("N", "T", "H", "W", "Cin", "Cout", "kt", "kh", "kw", "pt", "ph", "pw", "tT", "tH", "tW",
                      "sT", "sH", "sW", "nsplit", "relu", "accumulate", "dgrad", "y_f32_ncdhw",
                      "in_cstride", "in_coff", "out_cstride", "out_coff")
                + _ptrs("x_hi", "x_lo", "w_hi", "w_lo", "scale", "shift", "y_hi", "y_lo", "y_f32")
                + _ints("Cin2", "in2_cstride", "in2_coff") + _ptrs("x2_hi", "x2_lo", "w2_hi", "w2_lo") + _ints("ksplit"))


class Conv1aDesc(Structure):
    """Mirror of `otal_conv1a_desc`."""

    _fields_ = (_ints("N", "T", "H", "W", "Cout", "tT", "tH", "tW", "nsplit", "relu", "out_cstride", "out_coff")
                + _ptrs("x_hi", "x_lo", "w_hi", "w_lo", "scale", "shift", "y_hi", "y_lo"))


class WgradDesc(Structure):
    """Mirror of `otal_wgrad_desc`."""

    _fields_ = (_ints("N", "T", "H", "W", "Cin", "Cout", "kt", "kh", "kw", "pt", "ph", "pw", "sT", "sH", "sW",
                      "tT", "tH", "tW", "nsplit", "x_cstride", "x_coff", "d_cstride", "d_coff")
                + _ptrs("x_hi", "x_lo", "d_hi", "d_lo", "dw"))


class Conv1aWgradDesc(Structure):
    """Mirror of `otal_conv1a_wgrad_desc`."""

    _fields_ = (_ints("N", "T", "H", "W", "Cout", "tT", "tH", "tW", "nsplit", "d_cstride", "d_coff")
                + _ptrs("x_hi", "x_lo", "d_hi", "d_lo", "dw"))


class PoolDesc(Structure):
    """Mirror of `otal_pool_desc`."""

    _fields_ = (_ints("N", "T", "H", "W", "C", "kt", "kh", "kw", "st", "sh", "sw", "pt", "ph", "pw",
                      "in_cstride", "in_coff", "out_cstride", "out_coff",
                      "gout_cstride", "gout_coff", "gin_cstride", "gin_coff")
                + _ptrs("x_hi", "x_lo", "y_hi", "y_lo", "g_out", "g_in", "argmax"))


class MslDesc(Structure):
    """Mirror of `otal_msl_desc`."""

    _fields_ = (_ints("B", "P", "K", "G") + [("clip_length", c_float), ("overlap_thresh", c_float)] + _ints("use_ibm", "num_bins")
                + [("momentum", c_float)] + _ints("iou_aware") + [("act_weight", c_float), ("act_margin", c_float)]
                + _ints("prior_stride")
                + _ptrs("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act", "priors", "targets", "valid",
                        "weight_accum", "losses", "workspace")
                + _ints("flavour") + [("ibm_coeff", c_float), ("focal_alpha", c_float), ("focal_gamma", c_float),
                                      ("level_bounds", c_float * 16)]
                + _ints("reweight", "cls_all") + [("edl_focal_alpha", c_float), ("edl_focal_gamma", c_float)] + _ptrs("ghm_acc_sum"))


class GnDesc(Structure):
    """Mirror of `otal_gn_desc`."""

    _fields_ = (_ints("B", "C", "T", "groups") + [("eps", c_float)] + _ints("relu", "nseg") + [("seg_off", c_int * 8), ("seg_len", c_int * 8)]
                + _ptrs("x", "gamma", "beta", "mean", "rstd", "y", "p_hi", "p_lo") + _ints("p_cstride", "p_coff")
                + _ptrs("yt") + _ints("yt_off", "yt_T")
                + _ptrs("gy") + [("gy_bstride", c_longlong)] + _ptrs("gy2a", "gy2b") + _ints("gy2_off", "gy2_T")
                + _ptrs("gx", "d_hi", "d_lo", "dgamma", "dbeta", "dbias"))


class RowsDesc(Structure):
    """Mirror of `otal_rows_desc`."""

    _fields_ = (_ints("B", "C", "Td", "npairs", "nsrc") + [("src", c_void_p * 8), ("src_T", c_int * 8)]
                + _ptrs("table", "dst", "p_hi", "p_lo"))


class HeadoutDesc(Structure):
    """Mirror of `otal_headout_desc`."""

    _fields_ = (_ints("B", "S", "P", "n") + _ptrs("sep_idx", "level_id", "mult") + [("scale", c_void_p * 8), ("dscale", c_void_p * 8),
                ("raw", c_void_p * 4), ("cpad", c_int * 4), ("cout", c_int * 4), ("mode", c_int * 4), ("out", c_void_p * 4),
                ("gout", c_void_p * 4), ("d_hi", c_void_p * 4), ("d_lo", c_void_p * 4), ("bias", c_void_p * 4), ("dbias", c_void_p * 4)])


# name -> (restype, argtypes).  tests/test_abi.py checks this table against include/opental_b200.h.
SIGNATURES = {
    "otal_last_error": (c_char_p, []),
    "otal_abi_version": (c_int, []),
    "otal_abi_sizeof": (c_int, [c_char_p]),
    "otal_bmp_forward_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_bmp_backward_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_bmp_forward_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_bmp_backward_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_bmp_forward_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_bmp_backward_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_conv_igemm_fwd": (c_int, [POINTER(ConvDesc), c_void_p]),
    "otal_conv1a_fwd": (c_int, [POINTER(Conv1aDesc), c_void_p]),
    "otal_conv1a_fwd_u8": (c_int, [POINTER(Conv1aDesc), c_void_p]),
    "otal_conv1a_fwd_u8_halo": (c_int, [POINTER(Conv1aDesc), c_void_p]),
    "otal_conv1a_wgrad_u8": (c_int, [POINTER(Conv1aWgradDesc), c_void_p]),
    "otal_conv1a_wgrad_u8_halo": (c_int, [POINTER(Conv1aWgradDesc), c_void_p]),
    "otal_clip_ingest_u8_raw": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_border_class_sums": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_conv_wgrad": (c_int, [POINTER(WgradDesc), c_void_p]),
    "otal_conv1a_wgrad": (c_int, [POINTER(Conv1aWgradDesc), c_void_p]),
    "otal_maxpool_fwd": (c_int, [POINTER(PoolDesc), c_void_p]),
    "otal_maxpool_bwd": (c_int, [POINTER(PoolDesc), c_void_p]),
    "otal_clip_ingest": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_maxpool_bwd_relu_bn_split": (c_int, [POINTER(PoolDesc), c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                               c_void_p]),
    "otal_clip_ingest_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_relu_bn_bwd_split": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_int,
                                       c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_float, c_float, c_float, c_float,
                               c_float, c_float, c_int, c_void_p]),
    "otal_adam_step_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_float, c_float, c_float, c_float,
                                   c_float, c_float, c_void_p, c_void_p]),
    "otal_ncl_to_nlc_split": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_segment_mean": (c_int, [c_void_p, POINTER(c_longlong), c_int, c_void_p, c_void_p]),
    "otal_rows_combine": (c_int, [POINTER(RowsDesc), c_void_p]),
    "otal_head_gather_fwd": (c_int, [POINTER(HeadoutDesc), c_void_p]),
    "otal_head_gather_bwd": (c_int, [POINTER(HeadoutDesc), c_void_p, c_void_p]),
    "otal_ncl_to_nlc_split_ex": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_boundary_bce_fwd_ex": (c_int, [c_void_p, c_int, c_void_p, c_longlong, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "otal_boundary_bce_bwd_ex": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "otal_groupnorm_relu_fwd_ex": (c_int, [POINTER(GnDesc), c_void_p]),
    "otal_groupnorm_relu_bwd_ex": (c_int, [POINTER(GnDesc), c_void_p]),
    "otal_split_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "otal_merge_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "otal_msl_workspace_floats": (c_longlong, [c_int, c_int, c_int]),
    "otal_msl_forward": (c_int, [POINTER(MslDesc), c_void_p]),
    "otal_msl_backward": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "otal_groupnorm_relu_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                        c_float, c_int, c_int, POINTER(c_int), POINTER(c_int), c_void_p]),
    "otal_groupnorm_relu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                        c_int, c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), c_void_p]),
    "otal_make_segments": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                                   c_void_p]),
    "otal_make_segments_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_float,
                                      c_void_p]),
    "otal_dirichlet_uncertainty": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    "otal_boundary_bce_fwd": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "otal_boundary_bce_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "otal_decode_scores": (c_int, [c_void_p] * 13 + [c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "otal_softnms": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_float, c_void_p]),
    "otal_ncdhw_to_ndhwc_split": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
}

STRUCT_MIRRORS = {"otal_conv_desc": ConvDesc, "otal_conv1a_desc": Conv1aDesc, "otal_wgrad_desc": WgradDesc,
                  "otal_conv1a_wgrad_desc": Conv1aWgradDesc, "otal_pool_desc": PoolDesc, "otal_msl_desc": MslDesc,
                  "otal_gn_desc": GnDesc, "otal_rows_desc": RowsDesc, "otal_headout_desc": HeadoutDesc}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m opental_b200.build` "
                "(opental_b200 has no CPU or PyTorch fallback)"
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.otal_abi_version() != 1:
            raise RuntimeError("libopental_b200.so ABI version mismatch: rebuild the library")
        for cname, mirror in STRUCT_MIRRORS.items():
            if lib.otal_abi_sizeof(cname.encode()) != ctypes.sizeof(mirror):
                raise RuntimeError(f"{cname}: the ctypes mirror ({ctypes.sizeof(mirror)} bytes) does not match the library "
                                   f"({lib.otal_abi_sizeof(cname.encode())} bytes): rebuild the library or update _lib.py")
        _lib = lib
    return _lib


def last_error() -> str:
    return load().otal_last_error().decode()


def check(rc: int, what: str) -> None:
    if rc != OTAL_OK:
        raise RuntimeError(f"{what} failed: {ERR_NAMES.get(rc, rc)}: {last_error()}")


LAUNCHES = {}      # entry point -> number of successful calls (each launches at least one CUDA kernel of this library)
TRACE = None       # developer tracing (tools/step_profile.py): a list collects (entry point, label, start, end) events
LABEL = None       # free-form description of the next call (shape, role), set by ops.* when tracing


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise RuntimeError on a non-zero code."""
    global LABEL
    if TRACE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(getattr(load(), name)(*args), name)
        e1.record()
        TRACE.append((name, LABEL, e0, e1))
        LABEL = None
    else:
        check(getattr(load(), name)(*args), name)
    LAUNCHES[name] = LAUNCHES.get(name, 0) + 1


def launch_count() -> int:
    return sum(LAUNCHES.values())
