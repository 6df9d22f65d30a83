"""Tensor-level wrappers around the C ABI: allocate outputs with torch, pass raw pointers and the current stream.

PyTorch is plumbing here (device memory, streams); all arithmetic happens in libopental_b200.so.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import ConvDesc


def _require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("opental_b200 kernels need CUDA tensors (there is no CPU fallback)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


# ----------------------------------------------------------------------------------------------------------
# BoundaryMaxPooling
# ----------------------------------------------------------------------------------------------------------
_BMP_SUFFIX = {torch.float32: "f32", torch.float64: "f64"}


def _bmp_check(input: torch.Tensor, segments: torch.Tensor, *more: torch.Tensor) -> str:
    # error behaviour of boundary_max_pooling_cuda.cpp:4-6 (CHECK_CUDA / CHECK_CONTIGUOUS -> RuntimeError)
    for name, t in (("input", input), ("segments", segments), *[("grad_output", m) for m in more]):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")
    if input.dtype not in _BMP_SUFFIX:
        raise RuntimeError(f"boundary_max_pooling: unsupported dtype {input.dtype} (float32 / float64 only)")
    if segments.dtype != input.dtype:
        raise RuntimeError("segments must have the same dtype as input")
    if input.dim() != 3 or segments.dim() != 3 or segments.size(2) != 4:
        raise RuntimeError("expected input [B,C,T] and segments [B,K,4]")
    if segments.size(0) != input.size(0):
        raise RuntimeError("segments batch size must equal input batch size")
    return _BMP_SUFFIX[input.dtype]


def bmp_forward(input: torch.Tensor, segments: torch.Tensor) -> torch.Tensor:
    sfx = _bmp_check(input, segments)
    B, C, T = input.shape
    K = segments.size(1)
    out = torch.empty((B, C, K), dtype=input.dtype, device=input.device)
    with torch.cuda.device(input.device):
        _lib.call(f"otal_bmp_forward_{sfx}", input.data_ptr(), segments.data_ptr(), out.data_ptr(), B, C, T, K, _stream())
    return out


def bmp_backward(grad_output: torch.Tensor, input: torch.Tensor, segments: torch.Tensor,
                 compat_tscale_bug: bool = False) -> torch.Tensor:
    sfx = _bmp_check(input, segments, grad_output)
    B, C, T = input.shape
    K = segments.size(1)
    if tuple(grad_output.shape) != (B, C, K):
        raise RuntimeError("grad_output must be [B,C,K]")
    grad_in = torch.empty((B, C, T), dtype=grad_output.dtype, device=input.device)
    with torch.cuda.device(input.device):
        _lib.call(f"otal_bmp_backward_{sfx}", grad_output.data_ptr(), input.data_ptr(), segments.data_ptr(),
                  grad_in.data_ptr(), B, C, T, K, int(bool(compat_tscale_bug)), _stream())
    return grad_in


# ----------------------------------------------------------------------------------------------------------
# bf16 hi/lo planes
# ----------------------------------------------------------------------------------------------------------
@dataclass
class Planes:
    """An activation / weight tensor stored as bf16 hi (+ optional lo) planes, x ~= hi + lo."""

    hi: torch.Tensor
    lo: torch.Tensor | None

    @property
    def shape(self):
        return self.hi.shape

    def float(self) -> torch.Tensor:
        out = torch.empty(self.hi.shape, dtype=torch.float32, device=self.hi.device)
        _lib.call("otal_merge_bf16", self.hi.data_ptr(), _ptr(self.lo), out.data_ptr(), self.hi.numel(), _stream())
        return out


def split_bf16(x: torch.Tensor, with_lo: bool = True) -> Planes:
    _require_cuda(x)
    x = x.contiguous().float()
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if with_lo else None
    _lib.call("otal_split_bf16", x.data_ptr(), hi.data_ptr(), _ptr(lo), x.numel(), _stream())
    return Planes(hi, lo)


def clip_to_ndhwc(x: torch.Tensor, cpad: int = 8, with_lo: bool = True) -> Planes:
    """NCDHW fp32 clip -> NDHWC bf16 planes with channels zero-padded to `cpad`."""
    _require_cuda(x)
    x = x.contiguous().float()
    N, C, T, H, W = x.shape
    hi = torch.empty((N, T, H, W, cpad), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi) if with_lo else None
    _lib.call("otal_ncdhw_to_ndhwc_split", x.data_ptr(), hi.data_ptr(), _ptr(lo), N, C, T, H, W, cpad, _stream())
    return Planes(hi, lo)


# ----------------------------------------------------------------------------------------------------------
# implicit-GEMM convolution
# ----------------------------------------------------------------------------------------------------------
def pick_tile_box(T: int, H: int, W: int) -> tuple[int, int, int]:
    """128-position tile box (tT, tH, tW), powers of two, minimising padded positions; ties -> wider W."""
    best = None
    for lw in range(8):
        for lh in range(8 - lw):
            lt = 7 - lw - lh
            tw, th, tt = 1 << lw, 1 << lh, 1 << lt
            padded = (-(-T // tt) * tt) * (-(-H // th) * th) * (-(-W // tw) * tw)
            key = (padded, -tw, -th)
            if best is None or key < best[0]:
                best = (key, (tt, th, tw))
    return best[1]


def pack_conv_weight(w: torch.Tensor, with_lo: bool = True) -> Planes:
    """[Cout, Cin, kt, kh, kw] (or [Cout, Cin, k] for conv1d) fp32 -> [taps, Cout, Cin] bf16 planes."""
    if w.dim() == 3:
        w = w[:, :, :, None, None]
    Cout, Cin = w.shape[:2]
    wt = w.detach().float().permute(2, 3, 4, 0, 1).reshape(-1, Cout, Cin).contiguous()
    return split_bf16(wt, with_lo)


def conv_igemm(x: Planes, w: Planes, *, kernel: tuple[int, int, int], pad_front: tuple[int, int, int],
               scale: torch.Tensor | None = None, shift: torch.Tensor | None = None, relu: bool = False,
               in_slice: tuple[int, int] | None = None, out: Planes | None = None,
               out_slice: tuple[int, int] | None = None, out_f32: torch.Tensor | None = None,
               want_planes: bool = True, tile: tuple[int, int, int] | None = None) -> Planes | None:
    """y = relu?(conv(x, w) * scale + shift).  x: NDHWC planes [N,T,H,W,Cx]; w: [taps,Cout,Cin] planes.

    in_slice = (offset, Cin) reads a channel slice of x; out/out_slice = write into a slice of an existing buffer.
    """
    _require_cuda(x.hi, w.hi)
    N, T, H, W, Cx = x.hi.shape
    taps, Cout, Cin = w.hi.shape
    kt, kh, kw = kernel
    assert taps == kt * kh * kw
    in_coff, cin_used = in_slice if in_slice is not None else (0, Cx)
    assert cin_used == Cin, f"weight Cin {Cin} != input slice {cin_used}"
    nsplit = 3 if (x.lo is not None and w.lo is not None) else 1
    if out is None and want_planes:
        hi = torch.empty((N, T, H, W, Cout), dtype=torch.bfloat16, device=x.hi.device)
        out = Planes(hi, torch.empty_like(hi) if nsplit == 3 else None)
    if out is not None:
        out_cstride = out.hi.shape[-1]
        out_coff = out_slice[0] if out_slice is not None else 0
        if out_slice is not None:
            assert out_slice[1] == Cout
    else:
        assert out_f32 is not None
        out_cstride = out_f32.shape[-1]
        out_coff = out_slice[0] if out_slice is not None else 0
    if out_f32 is not None:
        assert out_f32.shape[-1] == out_cstride and out_f32.dtype == torch.float32
    tT, tH, tW = tile if tile is not None else pick_tile_box(T, H, W)
    d = ConvDesc(N=N, T=T, H=H, W=W, Cin=Cin, Cout=Cout, kt=kt, kh=kh, kw=kw,
                 pt=pad_front[0], ph=pad_front[1], pw=pad_front[2], tT=tT, tH=tH, tW=tW,
                 nsplit=nsplit, relu=int(relu), in_cstride=Cx, in_coff=in_coff,
                 out_cstride=out_cstride, out_coff=out_coff,
                 x_hi=x.hi.data_ptr(), x_lo=_ptr(x.lo) if nsplit == 3 else None,
                 w_hi=w.hi.data_ptr(), w_lo=_ptr(w.lo) if nsplit == 3 else None,
                 scale=_ptr(scale), shift=_ptr(shift),
                 y_hi=_ptr(out.hi) if out is not None else None,
                 y_lo=_ptr(out.lo) if (out is not None and nsplit == 3) else None,
                 y_f32=_ptr(out_f32))
    _lib.call("otal_conv_igemm_fwd", ctypes.byref(d), _stream())
    return out
