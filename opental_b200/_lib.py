"""ctypes binding of libopental_b200.so (the C ABI declared in include/opental_b200.h).

The library is the product: there is no Python/CPU fallback.  Importing this module never fails (so that
CPU-only tooling can inspect the package), but any attempt to *call* a kernel without the built library or
without a CUDA device raises `RuntimeError`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libopental_b200.so")

OTAL_OK = 0
ERR_NAMES = {-1: "OTAL_ERR_BAD_ARG", -2: "OTAL_ERR_CUDA", -3: "OTAL_ERR_DRIVER", -4: "OTAL_ERR_UNSUPPORTED"}


class ConvDesc(Structure):
    """Mirror of `otal_conv_desc` (include/opental_b200.h)."""

    _fields_ = [
        ("N", c_int), ("T", c_int), ("H", c_int), ("W", c_int),
        ("Cin", c_int), ("Cout", c_int),
        ("kt", c_int), ("kh", c_int), ("kw", c_int),
        ("pt", c_int), ("ph", c_int), ("pw", c_int),
        ("tT", c_int), ("tH", c_int), ("tW", c_int),
        ("nsplit", c_int), ("relu", c_int),
        ("in_cstride", c_int), ("in_coff", c_int),
        ("out_cstride", c_int), ("out_coff", c_int),
        ("x_hi", c_void_p), ("x_lo", c_void_p),
        ("w_hi", c_void_p), ("w_lo", c_void_p),
        ("scale", c_void_p), ("shift", c_void_p),
        ("y_hi", c_void_p), ("y_lo", c_void_p),
        ("y_f32", c_void_p),
    ]


# name -> (restype, argtypes).  tests/test_abi.py checks this table against include/opental_b200.h.
SIGNATURES = {
    "otal_last_error": (c_char_p, []),
    "otal_abi_version": (c_int, []),
    "otal_bmp_forward_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_bmp_backward_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_bmp_forward_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_bmp_backward_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "otal_conv_igemm_fwd": (c_int, [POINTER(ConvDesc), c_void_p]),
    "otal_split_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "otal_merge_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "otal_ncdhw_to_ndhwc_split": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
}

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
        _lib = lib
    return _lib


def last_error() -> str:
    return load().otal_last_error().decode()


def check(rc: int, what: str) -> None:
    if rc != OTAL_OK:
        raise RuntimeError(f"{what} failed: {ERR_NAMES.get(rc, rc)}: {last_error()}")


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise RuntimeError on a non-zero code."""
    check(getattr(load(), name)(*args), name)
