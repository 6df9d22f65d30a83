"""Tensor-level wrappers around the C ABI: allocate outputs with torch, pass raw pointers and the current stream.

PyTorch is plumbing here (device memory, streams); all arithmetic happens in libopental_b200.so.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import os

import torch

OVERLAP_WGRAD = os.environ.get("OTAL_NO_WGRAD_OVERLAP") is None      # see fork() / join() below

from . import _lib
from ._lib import Conv1aDesc, Conv1aWgradDesc, ConvDesc, GnDesc, HeadoutDesc, MslDesc, PoolDesc, RowsDesc, WgradDesc


class KernelProfile:
    """Optional per-kernel timing with CUDA events on the launching stream (bench.py's live roofline measurement).
    Disabled by default: `with ops.PROFILE.enabled(): ...` brackets every tensor-core conv launch with two events and
    records its algorithmic FLOPs (2 * output positions * Cout * Cin * taps, fp32 semantics — bf16x3 executes 3x that)."""

    def __init__(self):
        self.on = False
        self.records = []      # (kernel, start_event, end_event, flops)

    def enabled(self):
        prof = self

        class _Ctx:
            def __enter__(self_):
                global OVERLAP_WGRAD
                prof.on = True
                prof.records = []
                # a kernel's roofline time is measured with the kernel running alone: no side-stream overlap while profiling
                self_.overlap, OVERLAP_WGRAD = OVERLAP_WGRAD, False
                return prof

            def __exit__(self_, *a):
                global OVERLAP_WGRAD
                prof.on = False
                OVERLAP_WGRAD = self_.overlap

        return _Ctx()

    def begin(self):
        if not self.on:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def end(self, kernel: str, start, flops: float):
        if start is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.records.append((kernel, start, ev, flops))

    def summary(self) -> dict:
        """kernel -> dict(launches, ms, flops); call after torch.cuda.synchronize()."""
        out = {}
        for k, s, e, f in self.records:
            if f < 2e9:     # head-sized launches: their event brackets measure host enqueue gaps in eager mode, not the kernel
                k = k + " (launches < 2 GFLOP: 1-D head)"
            d = out.setdefault(k, dict(launches=0, ms=0.0, flops=0.0))
            d["launches"] += 1
            d["ms"] += s.elapsed_time(e)
            d["flops"] += f
        return out


PROFILE = KernelProfile()

# NVTX ranges around the stages of a step, for an Nsight Systems timeline of the EAGER path (a graph replay is one
# node; tools/graph_timeline.py shows its inside).  Host-side markers only, off unless OTAL_NVTX=1.
NVTX = os.environ.get("OTAL_NVTX", "0") == "1"


def nvtx_push(name: str) -> None:
    if NVTX and torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)


def nvtx_pop() -> None:
    if NVTX and torch.cuda.is_available():
        torch.cuda.nvtx.range_pop()


def _require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("opental_b200 kernels need CUDA tensors (there is no CPU fallback)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ----------------------------------------------------------------------------------------------------------
# fork / join: a layer's weight gradient does not depend on its data gradient.  `fork()` returns a side stream that has
# waited for everything enqueued on the current stream so far; `join()` makes the current stream wait for it.  Callers
# join before any tensor the side stream reads can be released (the caching allocator re-uses a block in stream order of
# the stream it was allocated on).  Works under CUDA-graph capture (the side stream joins the capture through the event).
# ----------------------------------------------------------------------------------------------------------
_SIDE: dict = {}


def fork() -> "torch.cuda.Stream":
    dev = torch.cuda.current_device()
    side = _SIDE.get(dev)
    if side is None:
        side = _SIDE[dev] = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    return side


def side_stream() -> "torch.cuda.Stream":
    dev = torch.cuda.current_device()
    side = _SIDE.get(dev)
    if side is None:
        side = _SIDE[dev] = torch.cuda.Stream(device=dev)
    return side


def join() -> None:
    side = _SIDE.get(torch.cuda.current_device())
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)


_SIDE2: dict = {}


class parallel_branch:
    """`with ops.parallel_branch():` runs the enclosed launches on a SECOND side stream that has waited for everything enqueued on the
    current stream so far; `ops.join_branch()` makes the current stream wait for it.  For two independent halves of a schedule (the
    loc / conf halves of the detection head): their kernels are far smaller than the GPU.  Same memory rule as fork / join: every
    tensor the branch touches stays referenced until the join.  Captures into a CUDA graph as a parallel branch."""

    def __enter__(self):
        dev = torch.cuda.current_device()
        side = _SIDE2.get(dev)
        if side is None:
            side = _SIDE2[dev] = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        self._ctx = torch.cuda.stream(side)
        self._ctx.__enter__()
        return side

    def __exit__(self, *exc):
        return self._ctx.__exit__(*exc)


def join_branch() -> None:
    side = _SIDE2.get(torch.cuda.current_device())
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


# ----------------------------------------------------------------------------------------------------------
# BoundaryMaxPooling
# ----------------------------------------------------------------------------------------------------------
_BMP_SUFFIX = {torch.float32: "f32", torch.float64: "f64", torch.float16: "f16"}      # the reference's dispatch set


def _bmp_check(input: torch.Tensor, segments: torch.Tensor, *more: torch.Tensor) -> str:
    # error behaviour of boundary_max_pooling_cuda.cpp:4-6 (CHECK_CUDA / CHECK_CONTIGUOUS -> RuntimeError)
    for name, t in (("input", input), ("segments", segments), *[("grad_output", m) for m in more]):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")
    if input.dtype not in _BMP_SUFFIX:
        raise RuntimeError(f"boundary_max_pooling: unsupported dtype {input.dtype} (float32 / float64 / float16, as the reference)")
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
# implicit-GEMM convolution (forward / dgrad / wgrad)
# ----------------------------------------------------------------------------------------------------------
def pick_tile_box(T: int, H: int, W: int, log2_positions: int = 7) -> tuple[int, int, int]:
    """Tile box (tT, tH, tW) of 2**log2_positions positions, powers of two, minimising padded positions; ties ->
    wider W.  128 positions for the forward/dgrad M tile, 64 for the wgrad K tile."""
    best = None
    n = log2_positions
    for lw in range(n + 1):
        for lh in range(n + 1 - lw):
            lt = n - lw - lh
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


def _out_extent(size: int, stride: int) -> int:
    return -(-size // stride)


def conv_igemm(x: Planes, w: Planes, *, kernel: tuple[int, int, int], pad_front: tuple[int, int, int],
               stride: tuple[int, int, int] = (1, 1, 1),
               scale: torch.Tensor | None = None, shift: torch.Tensor | None = None, relu: bool = False,
               in_slice: tuple[int, int] | None = None, out: Planes | None = None,
               out_slice: tuple[int, int] | None = None, out_f32: torch.Tensor | None = None,
               want_planes: bool = True, tile: tuple[int, int, int] | None = None,
               accumulate: bool = False, dgrad: bool = False, f32_ncdhw: bool = False,
               x2: Planes | None = None, w2: Planes | None = None, in2_slice: tuple[int, int] | None = None) -> Planes | None:
    """y = relu?(conv(x, w) * scale + shift).  x: NDHWC planes [N,T,H,W,Cx]; w: [taps,Cout,Cin] planes.

    in_slice = (offset, Cin) reads a channel slice of x; out/out_slice = write into a slice of an existing buffer.
    dgrad=True: x is the output gradient [.., forward Cout], w the FORWARD weights [taps, fwd Cout, fwd Cin]; the
    result is the input gradient [.., fwd Cin] (pad_front must be k-1-forward pad).
    """
    _require_cuda(x.hi, w.hi)
    N, T, H, W, Cx = x.hi.shape
    kt, kh, kw = kernel
    if dgrad:
        taps, Cin, Cout = w.hi.shape
    else:
        taps, Cout, Cin = w.hi.shape
    assert taps == kt * kh * kw
    in_coff, cin_used = in_slice if in_slice is not None else (0, Cx)
    assert cin_used == Cin, f"weight Cin {Cin} != input slice {cin_used}"
    nsplit = 3 if (x.lo is not None and w.lo is not None) else 1
    To, Ho, Wo = (_out_extent(a, s) for a, s in zip((T, H, W), stride))
    if out is None and want_planes:
        hi = torch.empty((N, To, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.hi.device)
        out = Planes(hi, torch.empty_like(hi) if nsplit == 3 else None)
    if out is not None:
        out_cstride = out.hi.shape[-1]
        assert tuple(out.hi.shape[:4]) == (N, To, Ho, Wo)
    else:
        assert out_f32 is not None
        out_cstride = out_f32.shape[1] if f32_ncdhw else out_f32.shape[-1]
    out_coff = out_slice[0] if out_slice is not None else 0
    if out_slice is not None:
        assert out_slice[1] == Cout
    if out_f32 is not None and f32_ncdhw:
        # channel-major destination [N, Ctot, To, Ho, Wo] (1-D head: [B, C, T] viewed as [B, C, T, 1, 1])
        assert out is None and out_f32.dtype == torch.float32 and out_f32.is_contiguous()
        assert out_f32.shape[0] == N and out_f32.numel() == N * out_cstride * To * Ho * Wo
    elif out_f32 is not None:
        assert out_f32.shape[-1] == out_cstride and out_f32.dtype == torch.float32 and out_f32.is_contiguous()
        assert tuple(out_f32.shape[:4]) == (N, To, Ho, Wo)
    tT, tH, tW = tile if tile is not None else pick_tile_box(To, Ho, Wo)
    d = ConvDesc(N=N, T=T, H=H, W=W, Cin=Cin, Cout=Cout, kt=kt, kh=kh, kw=kw,
                 pt=pad_front[0], ph=pad_front[1], pw=pad_front[2], tT=tT, tH=tH, tW=tW,
                 sT=stride[0], sH=stride[1], sW=stride[2],
                 nsplit=nsplit, relu=int(relu), accumulate=int(accumulate), dgrad=int(dgrad),
                 y_f32_ncdhw=int(f32_ncdhw), in_cstride=Cx, in_coff=in_coff, out_cstride=out_cstride, out_coff=out_coff,
                 x_hi=x.hi.data_ptr(), x_lo=_ptr(x.lo) if nsplit == 3 else None,
                 w_hi=w.hi.data_ptr(), w_lo=_ptr(w.lo) if nsplit == 3 else None,
                 scale=_ptr(scale), shift=_ptr(shift),
                 y_hi=_ptr(out.hi) if out is not None else None,
                 y_lo=_ptr(out.lo) if (out is not None and nsplit == 3) else None,
                 y_f32=_ptr(out_f32))
    flops2 = 0.0
    if x2 is not None:
        # second K segment (1x1 only): y = [x | x2] . [w ; w2]
        assert w2 is not None and taps == 1 and tuple(x2.hi.shape[:4]) == (N, T, H, W)
        c2 = w2.hi.shape[1] if dgrad else w2.hi.shape[2]
        off2, used2 = in2_slice if in2_slice is not None else (0, x2.hi.shape[-1])
        assert used2 == c2 and (w2.hi.shape[2] if dgrad else w2.hi.shape[1]) == Cout
        d.Cin2, d.in2_cstride, d.in2_coff = c2, x2.hi.shape[-1], off2
        d.x2_hi, d.x2_lo = x2.hi.data_ptr(), (_ptr(x2.lo) if nsplit == 3 else None)
        d.w2_hi, d.w2_lo = w2.hi.data_ptr(), (_ptr(w2.lo) if nsplit == 3 else None)
        flops2 = 2.0 * N * To * Ho * Wo * Cout * c2
    t0 = PROFILE.begin()
    if _lib.TRACE is not None:
        _lib.LABEL = (f"{'dgrad' if dgrad else 'fwd'} N{N} {T}x{H}x{W} Cin{Cin} Cout{Cout} k{kt}{kh}{kw} s{stride[0]}{stride[1]}{stride[2]} "
                      f"x{nsplit} tile{tT}x{tH}x{tW}{' f32' if out_f32 is not None else ''}{' acc' if accumulate else ''}",
                      2.0 * N * To * Ho * Wo * Cout * Cin * taps + flops2)
    _lib.call("otal_conv_igemm_fwd", ctypes.byref(d), _stream())
    PROFILE.end("conv_igemm_kernel", t0, 2.0 * N * To * Ho * Wo * Cout * Cin * taps + flops2)
    return out


def conv_wgrad(x: Planes, d: Planes, dw: torch.Tensor, *, kernel: tuple[int, int, int],
               pad_front: tuple[int, int, int], stride: tuple[int, int, int] = (1, 1, 1),
               in_slice: tuple[int, int] | None = None, d_slice: tuple[int, int] | None = None) -> None:
    """dw[tap, co, ci] += sum_p d[p, co] * x[s*p + tap - pad, ci]  (fp32, [taps, Cout, Cin], contiguous)."""
    _require_cuda(x.hi, d.hi, dw)
    N, T, H, W, Cx = x.hi.shape
    taps, Cout, Cin = dw.shape
    kt, kh, kw = kernel
    assert taps == kt * kh * kw and dw.dtype == torch.float32 and dw.is_contiguous()
    x_coff, cin_used = in_slice if in_slice is not None else (0, Cx)
    d_coff, cout_used = d_slice if d_slice is not None else (0, d.hi.shape[-1])
    assert cin_used == Cin and cout_used == Cout
    To, Ho, Wo = (_out_extent(a, s) for a, s in zip((T, H, W), stride))
    assert tuple(d.hi.shape[:4]) == (N, To, Ho, Wo)
    nsplit = 3 if (x.lo is not None and d.lo is not None) else 1
    tT, tH, tW = pick_tile_box(To, Ho, Wo, 6)
    desc = WgradDesc(N=N, T=T, H=H, W=W, Cin=Cin, Cout=Cout, kt=kt, kh=kh, kw=kw,
                     pt=pad_front[0], ph=pad_front[1], pw=pad_front[2], sT=stride[0], sH=stride[1], sW=stride[2],
                     tT=tT, tH=tH, tW=tW, nsplit=nsplit, x_cstride=Cx, x_coff=x_coff,
                     d_cstride=d.hi.shape[-1], d_coff=d_coff,
                     x_hi=x.hi.data_ptr(), x_lo=_ptr(x.lo) if nsplit == 3 else None,
                     d_hi=d.hi.data_ptr(), d_lo=_ptr(d.lo) if nsplit == 3 else None, dw=dw.data_ptr())
    t0 = PROFILE.begin()
    if _lib.TRACE is not None:
        _lib.LABEL = (f"wgrad N{N} {T}x{H}x{W} Cin{Cin} Cout{Cout} k{kt}{kh}{kw} s{stride[0]}{stride[1]}{stride[2]} x{nsplit}",
                      2.0 * N * To * Ho * Wo * Cout * Cin * taps)
    _lib.call("otal_conv_wgrad", ctypes.byref(desc), _stream())
    PROFILE.end("conv_wgrad_kernel", t0, 2.0 * N * To * Ho * Wo * Cout * Cin * taps)


# ----------------------------------------------------------------------------------------------------------
# Conv3d_1a_7x7 (stride 2, 3 channels) in the folded layout
# ----------------------------------------------------------------------------------------------------------
CLIP_CPAD = 4          # channel slots per pixel of the padded clip
CLIP_WIN = 8           # pixels per window (7 W taps + 1 zero-weight slot)
CLIP_WPAD = 8          # padded row = W + 8 pixels (2 left, 6 right)


def clip_ingest(x: torch.Tensor, with_lo: bool = True) -> Planes:
    """NCDHW fp32 clip [N,3,T,H,W] -> W-padded [N,T,H,W+8,4] bf16 planes, the input of conv1a_fwd."""
    _require_cuda(x)
    x = x.contiguous().float()
    N, C, T, H, W = x.shape
    assert W % 2 == 0 and C <= CLIP_CPAD
    hi = torch.empty((N, T, H, W + CLIP_WPAD, CLIP_CPAD), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi) if with_lo else None
    _lib.call("otal_clip_ingest", x.data_ptr(), hi.data_ptr(), _ptr(lo), N, C, T, H, W, _stream())
    return Planes(hi, lo)


def clip_ingest_u8(px: torch.Tensor, crop: int = 96, offsets: torch.Tensor | None = None, with_lo: bool = True,
                   frame_map: torch.Tensor | None = None, raw: bool = False) -> Planes:
    """uint8 frames [N,T,Hs,Ws,3] -> W-padded planes [N,T,crop,crop+8,4] of the normalised crop ((x/255)*2-1, bit-identical).
    offsets: optional int32 device tensor [N,3] = (row offset, column offset, mirror flag); default centre crop.
    frame_map: optional int32 device tensor [N,T], output frame t <- source frame frame_map[n,t] (SSL cut-paste).
    raw: ONE plane holding the pixel values 0..255 (exact in bf16) — the operand of conv1a_fwd / conv1a_wgrad with u8=True."""
    _require_cuda(px)
    assert px.dtype == torch.uint8 and px.dim() == 5 and px.shape[-1] == 3 and px.is_contiguous()
    N, T, Hs, Ws, _ = px.shape
    if offsets is not None:
        assert offsets.dtype == torch.int32 and offsets.is_cuda and tuple(offsets.shape) == (N, 3) and offsets.is_contiguous()
    if frame_map is not None:
        assert frame_map.dtype == torch.int32 and frame_map.is_cuda and tuple(frame_map.shape) == (N, T) and frame_map.is_contiguous()
    hi = torch.empty((N, T, crop, crop + CLIP_WPAD, CLIP_CPAD), dtype=torch.bfloat16, device=px.device)
    if raw:
        _lib.call("otal_clip_ingest_u8_raw", px.data_ptr(), _ptr(offsets), _ptr(frame_map), hi.data_ptr(), N, T, Hs, Ws, crop, crop,
                  _stream())
        return Planes(hi, None)
    lo = torch.empty_like(hi) if with_lo else None
    _lib.call("otal_clip_ingest_u8", px.data_ptr(), _ptr(offsets), _ptr(frame_map), hi.data_ptr(), _ptr(lo), N, T, Hs, Ws, crop,
              crop, _stream())
    return Planes(hi, lo)


def pack_conv1a_weight(w: torch.Tensor, with_lo: bool = True) -> Planes:
    """[Cout, 3, 7, 7, 7] fp32 -> [49 (dt,dh), Cout, 32] planes, row element dw*4 + c (zero for dw == 7, c == 3)."""
    Cout, C, kt, kh, kw = w.shape
    assert (kt, kh, kw) == (7, 7, 7) and C <= CLIP_CPAD
    f = torch.zeros((kt, kh, Cout, 8, CLIP_CPAD), dtype=torch.float32, device=w.device)
    f[:, :, :, :kw, :C] = w.detach().float().permute(2, 3, 0, 4, 1)
    return split_bf16(f.reshape(kt * kh, Cout, 8 * CLIP_CPAD), with_lo)


def unpack_conv1a_wgrad(dw: torch.Tensor, C: int = 3) -> torch.Tensor:
    """[49, Cout, 32] folded weight gradient -> [Cout, C, 7, 7, 7]."""
    Cout = dw.shape[1]
    return dw.reshape(7, 7, Cout, 8, CLIP_CPAD)[:, :, :, :7, :C].permute(2, 4, 0, 1, 3).contiguous()


# Conv3d_1a forward on raw pixels through the resident-halo kernel (csrc/conv1a_halo.cu); OTAL_CONV1A_HALO=0: the generic kernel
CONV1A_HALO = os.environ.get("OTAL_CONV1A_HALO", "1") != "0"


def pack_conv1a_weight_cat(w: Planes) -> torch.Tensor:
    """[49, Cout, 32] hi / lo planes -> ONE bf16 tensor [49, 2*Cout, 32]: per tap the hi rows, then the lo rows (the
    N-concatenated B operand of otal_conv1a_fwd_u8_halo as a single TMA box)."""
    return torch.cat([w.hi, w.lo], dim=1).contiguous()


def conv1a_fwd(x: Planes, w: Planes, W: int, *, scale: torch.Tensor | None, shift: torch.Tensor | None,
               relu: bool = True, out: Planes | None = None, out_slice: tuple[int, int] | None = None, u8: bool = False,
               w_cat: torch.Tensor | None = None) -> Planes:
    """u8: x is the raw-pixel plane (clip_ingest_u8(raw=True)), scale / shift come from conv1a_u8_scale_shift; with `w_cat`
    (pack_conv1a_weight_cat) and 64 output channels the resident-halo kernel runs."""
    N, T, H, Wp_, C4 = x.hi.shape
    taps, Cout, K = w.hi.shape
    assert Wp_ == W + CLIP_WPAD and C4 == CLIP_CPAD and taps == 49 and K == CLIP_WIN * CLIP_CPAD
    nsplit = 3 if ((x.lo is not None or u8) and w.lo is not None) else 1
    if u8:
        assert nsplit == 3 and shift is not None and tuple(shift.shape) == (4, 4, 4, Cout) and shift.is_contiguous()
    To, Ho, Wo = -(-T // 2), -(-H // 2), W // 2
    if out is None:
        hi = torch.empty((N, To, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.hi.device)
        out = Planes(hi, torch.empty_like(hi) if nsplit == 3 else None)
    tT, tH, tW = pick_tile_box(To, Ho, Wo)
    d = Conv1aDesc(N=N, T=T, H=H, W=W, Cout=Cout, tT=tT, tH=tH, tW=tW, nsplit=nsplit, relu=int(relu),
                   out_cstride=out.hi.shape[-1], out_coff=out_slice[0] if out_slice else 0,
                   x_hi=x.hi.data_ptr(), x_lo=_ptr(x.lo) if nsplit == 3 and not u8 else None,
                   w_hi=w.hi.data_ptr(), w_lo=_ptr(w.lo) if nsplit == 3 else None,
                   scale=_ptr(scale), shift=_ptr(shift), y_hi=out.hi.data_ptr(),
                   y_lo=_ptr(out.lo) if nsplit == 3 else None)
    t0 = PROFILE.begin()
    if _lib.TRACE is not None:
        _lib.LABEL = (f"conv1a fwd N{N} {T}x{H}x{W} Cout{Cout} x{nsplit}", 2.0 * N * To * Ho * Wo * Cout * 3 * 343)
    if u8 and CONV1A_HALO and w_cat is not None and Cout == 64 and T % 2 == 0 and H % 2 == 0 and W % 2 == 0 and min(T, H, W) >= 6:
        assert tuple(w_cat.shape) == (49, 2 * Cout, CLIP_WIN * CLIP_CPAD) and w_cat.dtype == torch.bfloat16 and w_cat.is_contiguous()
        d.w_hi, d.w_lo = w_cat.data_ptr(), None
        _lib.call("otal_conv1a_fwd_u8_halo", ctypes.byref(d), _stream())
    else:
        _lib.call("otal_conv1a_fwd_u8" if u8 else "otal_conv1a_fwd", ctypes.byref(d), _stream())
    PROFILE.end("conv_igemm_kernel", t0, 2.0 * N * To * Ho * Wo * Cout * 3 * 343)   # algorithmic: 3 channels, 7^3 taps
    return out


# the resident-halo weight-gradient kernel of the raw-uint8 Conv3d_1a (csrc/conv1a_wgrad_halo.cu); OTAL_CONV1A_WGRAD_HALO=0: the generic one
CONV1A_WGRAD_HALO = os.environ.get("OTAL_CONV1A_WGRAD_HALO", "1") != "0"


def conv1a_wgrad(x: Planes, d: Planes, dw: torch.Tensor, W: int, d_slice: tuple[int, int] | None = None, u8: bool = False) -> None:
    """dw [49, Cout, 32] fp32 += folded weight gradient of Conv3d_1a.  u8: x is the raw-pixel plane and dw receives the
    gradient against the pixel values (conv1a_u8_weight_grad turns it into the gradient of the reference's conv)."""
    N, T, H, Wp_, C4 = x.hi.shape
    taps, Cout, K = dw.shape
    assert Wp_ == W + CLIP_WPAD and C4 == CLIP_CPAD and taps == 49 and K == CLIP_WIN * CLIP_CPAD
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    nsplit = 3 if ((x.lo is not None or u8) and d.lo is not None) else 1
    assert not u8 or nsplit == 3
    To, Ho, Wo = -(-T // 2), -(-H // 2), W // 2
    assert tuple(d.hi.shape[:4]) == (N, To, Ho, Wo)
    tT, tH, tW = pick_tile_box(To, Ho, Wo, 6)
    desc = Conv1aWgradDesc(N=N, T=T, H=H, W=W, Cout=Cout, tT=tT, tH=tH, tW=tW, nsplit=nsplit,
                           d_cstride=d.hi.shape[-1], d_coff=d_slice[0] if d_slice else 0,
                           x_hi=x.hi.data_ptr(), x_lo=_ptr(x.lo) if nsplit == 3 and not u8 else None,
                           d_hi=d.hi.data_ptr(), d_lo=_ptr(d.lo) if nsplit == 3 else None, dw=dw.data_ptr())
    t0 = PROFILE.begin()
    if _lib.TRACE is not None:
        _lib.LABEL = (f"conv1a wgrad N{N} {T}x{H}x{W} Cout{Cout} x{nsplit}", 2.0 * N * To * Ho * Wo * Cout * 3 * 343)
    halo = u8 and CONV1A_WGRAD_HALO and Cout == 64 and T % 2 == 0 and H % 2 == 0 and W % 2 == 0 and min(T, H, W) >= 6
    _lib.call(("otal_conv1a_wgrad_u8_halo" if halo else "otal_conv1a_wgrad_u8") if u8 else "otal_conv1a_wgrad", ctypes.byref(desc), _stream())
    PROFILE.end("conv_wgrad_kernel", t0, 2.0 * N * To * Ho * Wo * Cout * 3 * 343)


# ---- Conv3d_1a on raw uint8 pixels: the host-side algebra around otal_conv1a_fwd_u8 / otal_conv1a_wgrad_u8 (pure torch on
# [Cout,3,7,7,7]-sized tensors; device-agnostic so the CPU tests can pin it against F.conv3d) -------------------------------
U8_SCALE = 2.0 / 255.0


def border_class_masks(device=None) -> torch.Tensor:
    """[4 classes, 7 taps] 0/1: which taps of a 7-tap, stride-2, front-pad-2 window lie inside an even-sized image for an
    output index of class 0 (interior), 1 (o == 0: taps 0,1 outside), 2 (o == n-2: tap 6 outside), 3 (o == n-1: taps 4..6)."""
    m = torch.ones(4, 7, device=device)
    m[1, :2] = 0
    m[2, 6:] = 0
    m[3, 4:] = 0
    return m


def border_classes(n: int, device=None) -> torch.Tensor:
    """Class of every output index 0..n-1 (n >= 3)."""
    assert n >= 3
    c = torch.zeros(n, dtype=torch.long, device=device)
    c[0], c[n - 2], c[n - 1] = 1, 2, 3
    return c


def conv1a_u8_scale_shift(w: torch.Tensor, bn_scale: torch.Tensor, bn_shift: torch.Tensor):
    """(scale [Cout], shift table [4,4,4,Cout]) for conv1a_fwd(u8=True).  With x = (2/255) u - 1 inside the image and the
    reference's zero padding of x outside (i3d_backbone.py:59-79): bn(conv(x, W)) = bn_scale * (2/255) * conv(u, W)
    + bn_shift - bn_scale * S_class, S_class = sum of W over the channels and the taps that are inside the image."""
    m = border_class_masks(w.device).to(w.dtype)
    ws = w.detach().sum(1)                                                        # [Cout,7,7,7]
    s = torch.einsum("othw,at,bh,cw->abco", ws, m, m, m)                          # [4,4,4,Cout]
    return bn_scale * U8_SCALE, (bn_shift - bn_scale * s).contiguous()


def conv1a_u8_weight_grad(dw_raw: torch.Tensor, class_sums: torch.Tensor | None = None, C: int = 3) -> torch.Tensor:
    """dW [Cout,C,7,7,7] of the reference's conv from the raw-pixel gradient dw_raw [49,Cout,32] (conv1a_wgrad(u8=True)):
    dW[co,ci,tap] = (2/255) * sum_p D[p,co] u[p+tap,ci] - R[co,tap],   R[co,tap] = sum_{p: tap inside the image} D[p,co].
    R comes for free: the raw plane carries 1.0 in channel slot 3 of every in-image pixel (otal_clip_ingest_u8_raw), so
    R[co,(dt,dh,dw)] = dw_raw[dt*7+dh, co, dw*4+3].  With `class_sums` ([4,4,4,Cout], border_class_sums) R is formed from the
    border-class sums of the output gradient instead (the round-1 form, kept for planes without the ones slot)."""
    if class_sums is None:
        Cout = dw_raw.shape[1]
        r = dw_raw.reshape(7, 7, Cout, 8, CLIP_CPAD)[:, :, :, :7, 3].permute(2, 0, 1, 3)      # [Cout,7,7,7]
    else:
        m = border_class_masks(dw_raw.device).to(dw_raw.dtype)
        r = torch.einsum("abco,at,bh,cw->othw", class_sums, m, m, m)              # [Cout,7,7,7]
    return unpack_conv1a_wgrad(dw_raw, C) * U8_SCALE - r[:, None]


def border_class_sums(d: Planes, d_slice: tuple[int, int] | None = None) -> torch.Tensor:
    """[4,4,4,C] fp32: sums of the NDHWC gradient planes d (hi + lo) over the positions of each border class."""
    N, To, Ho, Wo, cs = d.hi.shape
    coff, C = d_slice if d_slice else (0, cs)
    out = torch.zeros((4, 4, 4, C), dtype=torch.float32, device=d.hi.device)
    _lib.call("otal_border_class_sums", d.hi.data_ptr(), _ptr(d.lo), out.data_ptr(), N, To, Ho, Wo, C, cs, coff, _stream())
    return out


def ncl_to_nlc_planes(x: torch.Tensor, cpad: int | None = None, *, ttot: int | None = None, dilate: int = 1,
                      offset: int = 0, with_lo: bool = True) -> Planes:
    """[B,C,T] fp32 -> channels-last planes [B,Ttot,1,1,Cpad]; element (b,c,t) at position offset + t*dilate."""
    _require_cuda(x)
    x = x.contiguous().float()
    B, C, T = x.shape
    cpad = cpad or (C + 7) // 8 * 8
    ttot = ttot or T
    dense = cpad == C and dilate == 1 and ttot == T
    mk = torch.empty if dense else torch.zeros
    hi = mk((B, ttot, 1, 1, cpad), dtype=torch.bfloat16, device=x.device)
    lo = mk((B, ttot, 1, 1, cpad), dtype=torch.bfloat16, device=x.device) if with_lo else None
    _lib.call("otal_ncl_to_nlc_split", x.data_ptr(), hi.data_ptr(), _ptr(lo), B, C, T, cpad, ttot, dilate, offset, _stream())
    return Planes(hi, lo)


# ----------------------------------------------------------------------------------------------------------
# max pooling, ReLU/BN backward, Adam
# ----------------------------------------------------------------------------------------------------------
def _pool_desc(x: Planes, kernel, stride, pad_front, in_slice) -> tuple[PoolDesc, tuple[int, int, int, int], int]:
    N, T, H, W, Cx = x.hi.shape
    coff, C = in_slice if in_slice is not None else (0, Cx)
    d = PoolDesc(N=N, T=T, H=H, W=W, C=C, kt=kernel[0], kh=kernel[1], kw=kernel[2], st=stride[0], sh=stride[1],
                 sw=stride[2], pt=pad_front[0], ph=pad_front[1], pw=pad_front[2], in_cstride=Cx, in_coff=coff,
                 x_hi=x.hi.data_ptr(), x_lo=_ptr(x.lo))
    out_shape = (N, _out_extent(T, stride[0]), _out_extent(H, stride[1]), _out_extent(W, stride[2]))
    return d, out_shape, C


def maxpool_fwd(x: Planes, *, kernel, stride, pad_front, in_slice=None, out: Planes | None = None,
                out_slice=None, save_argmax: bool = False):
    """MaxPool3dSamePadding forward on planes.  save_argmax=True additionally returns the recorded arg-max bytes
    [N,To,Ho,Wo,C] that let maxpool_bwd scatter without re-reading the input."""
    _require_cuda(x.hi)
    d, oshape, C = _pool_desc(x, kernel, stride, pad_front, in_slice)
    if out is None:
        hi = torch.empty((*oshape, C), dtype=torch.bfloat16, device=x.hi.device)
        out = Planes(hi, torch.empty_like(hi) if x.lo is not None else None)
    assert tuple(out.hi.shape[:4]) == oshape
    d.out_cstride = out.hi.shape[-1]
    d.out_coff = out_slice[0] if out_slice else 0
    d.y_hi = out.hi.data_ptr()
    d.y_lo = _ptr(out.lo) if x.lo is not None else None
    arg = torch.empty((*oshape, C), dtype=torch.uint8, device=x.hi.device) if save_argmax else None
    d.argmax = _ptr(arg)
    if _lib.TRACE is not None:
        _lib.LABEL = (f"pool fwd {tuple(x.hi.shape)} C{C} k{kernel} s{stride}", 0.0)
    _lib.call("otal_maxpool_fwd", ctypes.byref(d), _stream())
    return (out, arg) if save_argmax else out


def maxpool_bwd(x: Planes, g_out: torch.Tensor, g_in: torch.Tensor, *, kernel, stride, pad_front, in_slice=None,
                gout_slice=None, gin_slice=None, argmax: torch.Tensor | None = None) -> None:
    """g_in[argmax window] += g_out  (fp32 NDHWC buffers; x = saved forward input planes, only read when `argmax`
    — the bytes recorded by maxpool_fwd(save_argmax=True) — is not given)."""
    _require_cuda(x.hi, g_out, g_in)
    d, oshape, C = _pool_desc(x, kernel, stride, pad_front, in_slice)
    assert tuple(g_out.shape[:4]) == oshape and tuple(g_in.shape[:4]) == tuple(x.hi.shape[:4])
    assert g_out.dtype == torch.float32 and g_in.dtype == torch.float32 and g_out.is_contiguous() and g_in.is_contiguous()
    d.gout_cstride = g_out.shape[-1]
    d.gout_coff = gout_slice[0] if gout_slice else 0
    d.gin_cstride = g_in.shape[-1]
    d.gin_coff = gin_slice[0] if gin_slice else 0
    d.g_out = g_out.data_ptr()
    d.g_in = g_in.data_ptr()
    if argmax is not None:
        assert argmax.dtype == torch.uint8 and argmax.is_contiguous() and tuple(argmax.shape) == (*oshape, C)
        d.argmax = argmax.data_ptr()
    if _lib.TRACE is not None:
        _lib.LABEL = (f"pool bwd {tuple(x.hi.shape)} C{C} k{kernel} s{stride}", 0.0)
    _lib.call("otal_maxpool_bwd", ctypes.byref(d), _stream())


def maxpool_bwd_relu_bn_split(x: Planes, argmax: torch.Tensor, g_out: torch.Tensor, scale: torch.Tensor | None, *, kernel, stride,
                              pad_front, g_add: torch.Tensor | None = None, with_lo: bool = True) -> Planes:
    """d = (maxpool_bwd(g_out) [+ g_add]) * [x > 0] * scale as bf16 planes shaped like x — the pool backward fused with the
    ReLU / frozen-BN backward of the layer that produced x (gather form: no atomics, no fp32 intermediate)."""
    _require_cuda(x.hi, argmax, g_out)
    d, oshape, C = _pool_desc(x, kernel, stride, pad_front, None)
    assert tuple(g_out.shape[:4]) == oshape and g_out.dtype == torch.float32 and g_out.is_contiguous()
    assert argmax.dtype == torch.uint8 and tuple(argmax.shape) == (*oshape, C) and C == x.hi.shape[-1]
    d.gout_cstride, d.gout_coff, d.g_out, d.argmax = g_out.shape[-1], 0, g_out.data_ptr(), argmax.data_ptr()
    if g_add is not None:
        assert g_add.dtype == torch.float32 and g_add.is_contiguous() and tuple(g_add.shape) == tuple(x.hi.shape)
    hi = torch.empty(x.hi.shape, dtype=torch.bfloat16, device=x.hi.device)
    lo = torch.empty_like(hi) if with_lo else None
    if _lib.TRACE is not None:
        _lib.LABEL = (f"pool bwd+relu_bn {tuple(x.hi.shape)} k{kernel} s{stride}", 0.0)
    _lib.call("otal_maxpool_bwd_relu_bn_split", ctypes.byref(d), _ptr(g_add), C, 0, _ptr(scale), hi.data_ptr(), _ptr(lo), C, 0, _stream())
    return Planes(hi, lo)


def relu_bn_bwd_split(g: torch.Tensor, y: Planes | None, scale: torch.Tensor | None, *, C: int | None = None,
                      g_slice=None, y_slice=None, relu: bool = True, with_lo: bool = True) -> Planes:
    """d = g * [y > 0] * scale  as bf16 planes [.., C] (dense)."""
    _require_cuda(g)
    assert g.dtype == torch.float32 and g.is_contiguous()
    Cg = g.shape[-1]
    g_coff, C_ = g_slice if g_slice is not None else (0, Cg)
    C = C or C_
    npos = g.numel() // Cg
    hi = torch.empty((*g.shape[:-1], C), dtype=torch.bfloat16, device=g.device)
    lo = torch.empty_like(hi) if with_lo else None
    y_cs = y.hi.shape[-1] if y is not None else 8
    y_co = y_slice[0] if y_slice else 0
    if _lib.TRACE is not None:
        _lib.LABEL = (f"relu_bn_bwd {tuple(g.shape)} C{C}", 0.0)
    _lib.call("otal_relu_bn_bwd_split", g.data_ptr(), _ptr(y.hi) if y is not None else None, _ptr(scale), hi.data_ptr(),
              _ptr(lo), npos, C, Cg, g_coff, y_cs, y_co, C, 0, int(relu and y is not None), _stream())
    return Planes(hi, lo)


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, *, lr: float, betas=(0.9, 0.999),
              eps: float = 1e-8, weight_decay: float = 0.0, grad_scale: float = 1.0, step: int = 1,
              step_dev: torch.Tensor | None = None) -> None:
    """In-place fused Adam (L2-in-gradient weight decay) over flat fp32 buffers.  step_dev: int32 [1] device tensor holding the
    step counter t >= 1 (then `step` is ignored): the launch can be captured into a CUDA graph."""
    _require_cuda(p, g, m, v)
    for t in (p, g, m, v):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == p.numel()
    if step_dev is not None:
        assert step_dev.dtype == torch.int32 and step_dev.device == p.device
        _lib.call("otal_adam_step_dev", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, betas[0], betas[1],
                  eps, weight_decay, grad_scale, step_dev.data_ptr(), _stream())
        return
    _lib.call("otal_adam_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, betas[0], betas[1],
              eps, weight_decay, grad_scale, step, _stream())


# ----------------------------------------------------------------------------------------------------------
# MultiSegmentLoss (single-CTA fused kernel)
# ----------------------------------------------------------------------------------------------------------
MSL_THUMOS, MSL_ANET, MSL_FOCAL = 0, 1, 2
MSL_RW_NONE, MSL_RW_IBM, MSL_RW_IB, MSL_RW_FOCAL, MSL_RW_GHM = 0, 1, 2, 3, 4


def msl_forward(loc, conf, prop_loc, prop_conf, center, act, prop_act, priors, targets, valid, weight_accum, *,
                clip_length: float, overlap_thresh: float, use_ibm: bool, momentum: float, iou_aware: bool,
                act_weight: float, act_margin: float, flavour: int = MSL_THUMOS, ibm_coeff: float = 10.0,
                focal_alpha: float = 0.25, focal_gamma: float = 2.0, level_bounds=(), reweight: int = 0, cls_all: bool = False,
                edl_focal_alpha: float = 0.25, edl_focal_gamma: float = 2.0, ghm_acc_sum: torch.Tensor | None = None,
                num_bins: int | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    """All 7 losses (+ N, PN, AN, PAN, loss_iouc) and the unit-gradient workspace.  See include/opental_b200.h.
    flavour: MSL_THUMOS (OpenTAL EDL), MSL_ANET (per-sample ActivityNet loss; level_bounds = ((left, right), ...) per pyramid
    level, priors [P,2] with the level in column 1) or MSL_FOCAL (closed-set softmax focal loss; act / prop_act None).
    reweight (MSL_THUMOS): MSL_RW_IBM / _IB / _FOCAL / _GHM select a re-weighting branch of EvidenceLoss (0: use_ibm decides);
    ghm_acc_sum = the fp64 [num_bins] GHM state; cls_all = no os_head (every prior a classification sample, class 0 = background)."""
    _require_cuda(loc, conf, prop_loc, prop_conf, center, priors, targets, valid)
    B, P, K = conf.shape
    G = targets.shape[1]
    ts = [t.contiguous() if t is not None else None for t in (loc, conf, prop_loc, prop_conf, center, act, prop_act)]
    for t in ts:
        assert t is None or t.dtype == torch.float32
    targets = targets.contiguous().float()
    valid = valid.contiguous()
    assert valid.element_size() == 1 and priors.dtype == torch.float32
    losses = torch.empty(16, dtype=torch.float32, device=loc.device)
    ws = torch.empty(int(_lib.load().otal_msl_workspace_floats(B, P, K)), dtype=torch.float32, device=loc.device)
    d = MslDesc(B=B, P=P, K=K, G=G, clip_length=clip_length, overlap_thresh=overlap_thresh, use_ibm=int(use_ibm),
                num_bins=(num_bins if num_bins is not None else (weight_accum.numel() if weight_accum is not None else 0)), momentum=momentum,
                iou_aware=int(iou_aware),
                act_weight=act_weight, act_margin=act_margin, prior_stride=priors.stride(0),
                loc=ts[0].data_ptr(), conf=ts[1].data_ptr(), prop_loc=ts[2].data_ptr(), prop_conf=ts[3].data_ptr(),
                center=ts[4].data_ptr(), act=_ptr(ts[5]), prop_act=_ptr(ts[6]), priors=priors.data_ptr(),
                targets=targets.data_ptr(), valid=valid.data_ptr(), weight_accum=_ptr(weight_accum),
                losses=losses.data_ptr(), workspace=ws.data_ptr(), flavour=int(flavour), ibm_coeff=ibm_coeff,
                focal_alpha=focal_alpha, focal_gamma=focal_gamma, reweight=int(reweight), cls_all=int(cls_all),
                edl_focal_alpha=edl_focal_alpha, edl_focal_gamma=edl_focal_gamma, ghm_acc_sum=_ptr(ghm_acc_sum))
    assert ghm_acc_sum is None or (ghm_acc_sum.dtype == torch.float64 and ghm_acc_sum.is_contiguous() and ghm_acc_sum.device == loc.device)
    flat = [float(v) for pair in level_bounds for v in pair]
    assert len(flat) <= 16, "msl: at most 8 pyramid levels"
    for i, v in enumerate(flat):
        d.level_bounds[i] = v
    _lib.call("otal_msl_forward", ctypes.byref(d), _stream())
    return losses, ws


def msl_backward(ws: torch.Tensor, grad_losses: torch.Tensor, B: int, P: int, K: int, with_act: bool):
    """Input gradients (loc, conf, prop_loc, prop_conf, center, act, prop_act) from the workspace of msl_forward."""
    dev = ws.device
    grad_losses = grad_losses.contiguous().float()
    assert grad_losses.numel() >= 7
    g_loc = torch.empty(B, P, 2, dtype=torch.float32, device=dev)
    g_ploc = torch.empty_like(g_loc)
    g_conf = torch.empty(B, P, K, dtype=torch.float32, device=dev)
    g_pconf = torch.empty_like(g_conf)
    g_center = torch.empty(B, P, dtype=torch.float32, device=dev)
    g_act = torch.empty(B, P, dtype=torch.float32, device=dev) if with_act else None
    g_pact = torch.empty(B, P, dtype=torch.float32, device=dev) if with_act else None
    _lib.call("otal_msl_backward", B, P, K, ws.data_ptr(), grad_losses.data_ptr(), g_loc.data_ptr(), g_conf.data_ptr(),
              g_ploc.data_ptr(), g_pconf.data_ptr(), g_center.data_ptr(), _ptr(g_act), _ptr(g_pact), _stream())
    return g_loc, g_conf, g_ploc, g_pconf, g_center, g_act, g_pact


# ----------------------------------------------------------------------------------------------------------
# GroupNorm + ReLU
# ----------------------------------------------------------------------------------------------------------
def _seg_arrays(segments):
    if not segments:
        return 0, None, None
    n = len(segments)
    return n, (ctypes.c_int * n)(*[int(o) for o, _ in segments]), (ctypes.c_int * n)(*[int(l) for _, l in segments])


class _GroupNormReLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, groups: int, eps: float, relu: bool, segments):
        _require_cuda(x, gamma, beta)
        x = x.contiguous()
        assert x.dtype == torch.float32 and x.dim() == 3
        B, C, T = x.shape
        y = torch.empty_like(x)
        nseg, so, sl = _seg_arrays(segments)
        stats = torch.empty(2, B * groups * max(nseg, 1), dtype=torch.float32, device=x.device)
        _lib.call("otal_groupnorm_relu_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), stats[0].data_ptr(),
                  stats[1].data_ptr(), B, C, T, groups, eps, int(relu), nseg, so, sl, _stream())
        ctx.save_for_backward(x, gamma, beta, stats)
        ctx.cfg = (groups, relu, segments)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gamma, beta, stats = ctx.saved_tensors
        groups, relu, segments = ctx.cfg
        B, C, T = x.shape
        gy = gy.contiguous()
        gx = torch.empty_like(x)
        dgb = torch.empty(B, 2, C, dtype=torch.float32, device=x.device)
        nseg, so, sl = _seg_arrays(segments)
        _lib.call("otal_groupnorm_relu_bwd", gy.data_ptr(), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), stats[0].data_ptr(),
                  stats[1].data_ptr(), gx.data_ptr(), dgb.data_ptr(), B, C, T, groups, int(relu), nseg, so, sl, _stream())
        d = dgb.sum(0) if B > 1 else dgb[0]
        return gx, d[0], d[1], None, None, None, None


def groupnorm_relu(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int = 32, eps: float = 1e-5,
                   relu: bool = True, segments=None) -> torch.Tensor:
    """relu(group_norm(x)) on [B,C,T] fp32 in one launch (forward) / one launch + a [B,2,C] batch reduction (backward).
    segments: optional tuple of (offset, length) column ranges normalised independently; other columns become 0."""
    return _GroupNormReLUFn.apply(x, gamma, beta, groups, eps, relu, tuple(segments) if segments else None)


def _gn_desc(x, gamma, beta, groups, eps, relu, segments, mean, rstd) -> "GnDesc":
    B, C, T = x.shape
    d = GnDesc(B=B, C=C, T=T, groups=groups, eps=eps, relu=int(relu), nseg=len(segments) if segments else 0,
               x=x.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(), mean=mean.data_ptr(), rstd=rstd.data_ptr())
    for i, (o, l) in enumerate(segments or ()):
        d.seg_off[i], d.seg_len[i] = int(o), int(l)
    return d


def groupnorm_relu_fwd_ex(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, groups: int = 32, eps: float = 1e-5,
                          relu: bool = True, segments=None, want_y: bool = False, planes: Planes | None = None, planes_coff: int = 0,
                          want_planes: bool = False, with_lo: bool = True, yt_range: tuple[int, int] | None = None):
    """Extended GroupNorm + ReLU forward on x [B,C,T] fp32 (see otal_groupnorm_relu_fwd_ex).  Returns (y | None, planes | None,
    yt | None, (mean, rstd)).  planes: an existing channels-last buffer [B,T,1,1,Ctot] to write at channel planes_coff, or
    want_planes=True for a fresh [B,T,1,1,C] one; yt_range = (offset, length): fp32 channels-last copy [B,length,C] of those columns."""
    _require_cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 3
    B, C, T = x.shape
    ns = len(segments) if segments else 1
    mean = torch.empty(B * groups * ns, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    d = _gn_desc(x, gamma, beta, groups, eps, relu, segments, mean, rstd)
    y = torch.empty_like(x) if want_y else None
    if planes is None and want_planes:
        hi = torch.empty((B, T, 1, 1, C), dtype=torch.bfloat16, device=x.device)
        planes = Planes(hi, torch.empty_like(hi) if with_lo else None)
    yt = None
    if yt_range is not None:
        yt = torch.empty((B, yt_range[1], C), dtype=torch.float32, device=x.device)
        d.yt, d.yt_off, d.yt_T = yt.data_ptr(), int(yt_range[0]), int(yt_range[1])
    d.y = _ptr(y)
    if planes is not None:
        assert planes.hi.shape[0] == B and planes.hi.shape[1] == T
        d.p_hi, d.p_lo, d.p_cstride, d.p_coff = planes.hi.data_ptr(), _ptr(planes.lo), planes.hi.shape[-1], int(planes_coff)
    _lib.call("otal_groupnorm_relu_fwd_ex", ctypes.byref(d), _stream())
    return y, planes, yt, (mean, rstd)


def groupnorm_relu_bwd_ex(gy: torch.Tensor | None, x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, stats, *,
                          dgamma: torch.Tensor, dbeta: torch.Tensor, dbias: torch.Tensor | None, groups: int = 32, relu: bool = True,
                          segments=None, gy_coff: int = 0, with_lo: bool = True, want_gx: bool = False,
                          gy2=None, gy2_off: int = 0) -> tuple[Planes, torch.Tensor | None]:
    """Extended GroupNorm + ReLU backward.  gy [B,Ctot,T] fp32, of which channels [gy_coff, gy_coff + C) are this layer's output
    gradient (or None when only gy2 carries gradient); gy2 = (a, b): channels-last [B,T2,C/2] gradients of the two channel halves for
    the columns [gy2_off, gy2_off + T2) (either may be None).  Accumulates into dgamma / dbeta / dbias [C]; returns (gx as
    channels-last planes [B,T,1,1,C], gx fp32 [B,C,T] | None)."""
    B, C, T = x.shape
    mean, rstd = stats
    d = _gn_desc(x, gamma, beta, groups, 0.0, relu, segments, mean, rstd)
    if gy is not None:
        assert gy.dtype == torch.float32 and gy.is_contiguous() and gy.shape[0] == B and gy.shape[2] == T
        d.gy = gy.data_ptr() + 4 * gy_coff * T
        d.gy_bstride = gy.shape[1] * T
    if gy2 is not None:
        a, b = gy2
        t2 = (a if a is not None else b).shape[1]
        for t in (a, b):
            assert t is None or (t.is_contiguous() and tuple(t.shape) == (B, t2, C // 2) and t.dtype == torch.float32)
        d.gy2a, d.gy2b, d.gy2_off, d.gy2_T = _ptr(a), _ptr(b), int(gy2_off), int(t2)
    hi = torch.empty((B, T, 1, 1, C), dtype=torch.bfloat16, device=x.device)
    planes = Planes(hi, torch.empty_like(hi) if with_lo else None)
    gx = torch.empty_like(x) if want_gx else None
    d.gx, d.d_hi, d.d_lo = _ptr(gx), hi.data_ptr(), _ptr(planes.lo)
    d.dgamma, d.dbeta, d.dbias = dgamma.data_ptr(), dbeta.data_ptr(), _ptr(dbias)
    _lib.call("otal_groupnorm_relu_bwd_ex", ctypes.byref(d), _stream())
    return planes, gx


def rows_combine(srcs: list[torch.Tensor], table: torch.Tensor, *, want_f32: bool = False, want_planes: bool = False,
                 with_lo: bool = True):
    """dst[b,c,j] = sum_k srcs[table[j,k,0]][b,c,table[j,k,1]] (table: device int32 [Td,npairs,2], source -1 = no term).
    Returns (dst fp32 [B,C,Td] | None, channels-last planes [B,Td,1,1,C] | None)."""
    B, C = srcs[0].shape[:2]
    Td, npairs = table.shape[:2]
    assert table.dtype == torch.int32 and table.is_contiguous() and len(srcs) <= 8
    d = RowsDesc(B=B, C=C, Td=Td, npairs=npairs, nsrc=len(srcs), table=table.data_ptr())
    for i, t in enumerate(srcs):
        assert t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape[:2]) == (B, C)
        d.src[i], d.src_T[i] = t.data_ptr(), t.shape[2]
    dst = torch.empty((B, C, Td), dtype=torch.float32, device=table.device) if want_f32 else None
    planes = None
    if want_planes:
        hi = torch.empty((B, Td, 1, 1, C), dtype=torch.bfloat16, device=table.device)
        planes = Planes(hi, torch.empty_like(hi) if with_lo else None)
        d.p_hi, d.p_lo = hi.data_ptr(), _ptr(planes.lo)
    d.dst = _ptr(dst)
    _lib.call("otal_rows_combine", ctypes.byref(d), _stream())
    return dst, planes


def _headout_desc(raws, couts, modes, biases, sep_idx, level_id, mult, scales, S, P) -> "HeadoutDesc":
    B = raws[0].shape[0]
    d = HeadoutDesc(B=B, S=S, P=P, n=len(raws), sep_idx=sep_idx.data_ptr(), level_id=_ptr(level_id), mult=_ptr(mult))
    for i, sc in enumerate(scales or ()):
        d.scale[i] = sc.data_ptr()
    for k, (r, co, m, bi) in enumerate(zip(raws, couts, modes, biases)):
        assert r.dtype == torch.float32 and r.is_contiguous() and r.shape[0] == B and r.shape[2] == S
        d.raw[k], d.cpad[k], d.cout[k], d.mode[k], d.bias[k] = r.data_ptr(), r.shape[1], int(co), int(m), _ptr(bi)
    return d


def head_gather_fwd(raws, couts, modes, biases, sep_idx, level_id=None, mult=None, scales=None) -> list[torch.Tensor]:
    """Raw head conv outputs [B,cpad,S] (+ bias) -> the reference's [B,P,cout] tensors (see otal_head_gather_fwd)."""
    B, S, P = raws[0].shape[0], raws[0].shape[2], sep_idx.numel()
    d = _headout_desc(raws, couts, modes, biases, sep_idx, level_id, mult, scales, S, P)
    outs = [torch.empty((B, P, co), dtype=torch.float32, device=raws[0].device) for co in couts]
    for k, o in enumerate(outs):
        d.out[k] = o.data_ptr()
    _lib.call("otal_head_gather_fwd", ctypes.byref(d), _stream())
    return outs


def head_gather_bwd(raws, couts, modes, biases, dbiases, gouts, outs, sep_idx, prior_of_col, level_id=None, mult=None, scales=None,
                    dscales=None, with_lo: bool = True) -> list[Planes]:
    """Gradients of the [B,P,cout] outputs -> zero-padded channels-last planes [B,S,1,1,cpad] of the raw conv outputs' gradient;
    bias gradients are accumulated into dbiases[k]; ScaleExp heads (mode 1) accumulate d scale into dscales[level]."""
    B, S, P = raws[0].shape[0], raws[0].shape[2], sep_idx.numel()
    d = _headout_desc(raws, couts, modes, biases, sep_idx, level_id, mult, scales, S, P)
    for i, ds in enumerate(dscales or ()):
        d.dscale[i] = ds.data_ptr()
    for k, db in enumerate(dbiases):
        d.dbias[k] = _ptr(db)
    gouts = list(gouts)
    res = []
    for k, (g, o, r) in enumerate(zip(gouts, outs, raws)):
        if g is not None:
            g = g.contiguous()
            assert g.dtype == torch.float32 and g.numel() == B * P * couts[k]
            gouts[k] = g                          # keep the contiguous copy alive until the launch is enqueued
        d.gout[k], d.out[k] = _ptr(g), _ptr(o)
        hi = torch.empty((B, S, 1, 1, r.shape[1]), dtype=torch.bfloat16, device=r.device)
        pl = Planes(hi, torch.empty_like(hi) if with_lo else None)
        d.d_hi[k], d.d_lo[k] = hi.data_ptr(), _ptr(pl.lo)
        res.append(pl)
    _lib.call("otal_head_gather_bwd", ctypes.byref(d), prior_of_col.data_ptr(), _stream())
    return res


def ncl_to_nlc_into(x: torch.Tensor, planes: Planes, coff: int, x_coff: int = 0, C: int | None = None) -> None:
    """Channels [x_coff, x_coff + C) of x [B,Ctot,T] fp32 -> channels [coff, coff + C) of the channels-last planes [B,T,1,1,Cdst]."""
    B, Ctot, T = x.shape
    C = Ctot - x_coff if C is None else C
    assert x.dtype == torch.float32 and x.is_contiguous() and planes.hi.shape[0] == B and planes.hi.shape[1] == T
    _lib.call("otal_ncl_to_nlc_split_ex", x.data_ptr() + 4 * x_coff * T, Ctot * T, planes.hi.data_ptr(), _ptr(planes.lo), B, C, T,
              planes.hi.shape[-1], int(coff), _stream())


# ----------------------------------------------------------------------------------------------------------
# detection-head helpers
# ----------------------------------------------------------------------------------------------------------
def make_segments(loc: torch.Tensor, prior: torch.Tensor, level_len: torch.Tensor, level_off: torch.Tensor, frame_num: float,
                  want_level: bool = False):
    """(seg_level | None, seg_concat, frame_seg), each [B,P,4] — see otal_make_segments in include/opental_b200.h."""
    _require_cuda(loc, prior, level_len, level_off)
    loc = loc.detach().contiguous()
    B, P, _ = loc.shape
    assert loc.dtype == torch.float32 and prior.numel() == P and level_len.dtype == torch.int32 and level_off.dtype == torch.int32
    seg_c = torch.empty(B, P, 4, dtype=torch.float32, device=loc.device)
    fseg = torch.empty_like(seg_c)
    seg_l = torch.empty_like(seg_c) if want_level else None
    _lib.call("otal_make_segments", loc.data_ptr(), prior.data_ptr(), level_len.data_ptr(), level_off.data_ptr(), _ptr(seg_l),
              seg_c.data_ptr(), fseg.data_ptr(), B, P, float(frame_num), _stream())
    return seg_l, seg_c, fseg


def dirichlet_uncertainty(logit: torch.Tensor) -> torch.Tensor:
    """K / sum(exp(clamp(logit, -10, 10)) + 1) over the last dim (no gradient: an inference-side score)."""
    _require_cuda(logit)
    x = logit.detach().contiguous().float()
    K = x.shape[-1]
    out = torch.empty(x.shape[:-1], dtype=torch.float32, device=x.device)
    _lib.call("otal_dirichlet_uncertainty", x.data_ptr(), out.data_ptr(), x.numel() // K, K, _stream())
    return out


# ----------------------------------------------------------------------------------------------------------
# inference post-processing
# ----------------------------------------------------------------------------------------------------------
def decode_scores(out: dict, offsets: torch.Tensor | None, clip_length: float, sample_fps: float):
    """decode_predictions (AFSD/thumos14/test.py:112-140) for every clip of a BDNet output dict.
    Returns (segments [B,P,2] seconds, scores [B,K,P], uncertainty [B,P], actionness [B,P])."""
    loc, conf = out["loc"].detach().contiguous(), out["conf"].detach().contiguous()
    _require_cuda(loc, conf)
    B, P, K = conf.shape
    ploc, pconf = out["prop_loc"].detach().contiguous(), out["prop_conf"].detach().contiguous()
    center = out["center"].detach().reshape(B, P).contiguous()
    act = out.get("act")
    pact = out.get("prop_act")
    act = act.detach().reshape(B, P).contiguous() if act is not None else None
    pact = pact.detach().reshape(B, P).contiguous() if pact is not None else None
    prior = out["priors"][:, 0].contiguous()
    if offsets is not None:
        offsets = offsets.to(device=loc.device, dtype=torch.float32).contiguous()
        assert offsets.numel() == B
    dev = loc.device
    seg = torch.empty(B, P, 2, dtype=torch.float32, device=dev)
    scores = torch.empty(B, K, P, dtype=torch.float32, device=dev)
    unct = torch.empty(B, P, dtype=torch.float32, device=dev)
    actn = torch.empty(B, P, dtype=torch.float32, device=dev)
    _lib.call("otal_decode_scores", loc.data_ptr(), ploc.data_ptr(), conf.data_ptr(), pconf.data_ptr(), center.data_ptr(), _ptr(act),
              _ptr(pact), prior.data_ptr(), _ptr(offsets), seg.data_ptr(), scores.data_ptr(), unct.data_ptr(), actn.data_ptr(), B, P, K,
              float(clip_length), float(sample_fps), _stream())
    return seg, scores, unct, actn


def softnms(segments: torch.Tensor, scores: torch.Tensor, sigma: float = 0.5, top_k: int = 1000, score_threshold: float = 0.001):
    """Gaussian soft-NMS per class (softnms_v2, AFSD/common/segment_utils.py:128-162).  segments [M,2] (shared by all classes)
    or [C,M,2]; scores [C,M] (entries below the threshold are ignored).  Returns (decayed scores [C,M], keep mask [C,M] bool,
    count [C] int32); the input scores are not modified."""
    _require_cuda(segments, scores)
    scores = scores.detach().float().contiguous().clone()
    segments = segments.detach().float().contiguous()
    C, M = scores.shape
    stride = 0 if segments.dim() == 2 else M * 2
    assert segments.shape[-2:] == (M, 2)
    keep = torch.empty(C, M, dtype=torch.uint8, device=scores.device)
    count = torch.empty(C, dtype=torch.int32, device=scores.device)
    _lib.call("otal_softnms", segments.data_ptr(), stride, scores.data_ptr(), keep.data_ptr(), count.data_ptr(), C, M, float(sigma),
              int(top_k), float(score_threshold), _stream())
    return scores, keep.bool(), count


# ----------------------------------------------------------------------------------------------------------
# boundary BCE (calc_bce_loss of the training scripts)
# ----------------------------------------------------------------------------------------------------------
class _BoundaryBCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target):
        """x [B,T,C] fp32; target [B,T] fp32 view with unit stride along T (a row of the [B,2,T] score maps)."""
        _require_cuda(x, target)
        B, T, C = x.shape
        # a channel slice of wider channels-last rows (the explicit head schedule hands out such views) is read in place
        sliced = (not x.is_contiguous()) and x.stride(2) == 1 and x.stride(0) == T * x.stride(1) and x.stride(1) >= C
        if not sliced:
            x = x.contiguous()
        assert target.shape == (B, T) and target.stride(1) == 1 and target.dtype == torch.float32 and x.dtype == torch.float32
        row_loss = torch.empty(B * T, dtype=torch.float32, device=x.device)
        coef = torch.empty(B * T, dtype=torch.float32, device=x.device)
        if sliced:
            _lib.call("otal_boundary_bce_fwd_ex", x.data_ptr(), x.stride(1), target.data_ptr(), target.stride(0), row_loss.data_ptr(),
                      coef.data_ptr(), B, T, C, _stream())
        else:
            _lib.call("otal_boundary_bce_fwd", x.data_ptr(), target.data_ptr(), target.stride(0), row_loss.data_ptr(), coef.data_ptr(),
                      B, T, C, _stream())
        ctx.save_for_backward(x, coef)
        ctx.sliced = sliced
        return row_loss.mean()

    @staticmethod
    def backward(ctx, g):
        x, coef = ctx.saved_tensors
        B, T, C = x.shape
        gx = torch.empty((B, T, C), dtype=torch.float32, device=x.device)
        g = g.contiguous().float().reshape(1)
        if ctx.sliced:
            _lib.call("otal_boundary_bce_bwd_ex", x.data_ptr(), x.stride(1), coef.data_ptr(), g.data_ptr(), gx.data_ptr(), B, T, C, _stream())
        else:
            _lib.call("otal_boundary_bce_bwd", x.data_ptr(), coef.data_ptr(), g.data_ptr(), gx.data_ptr(), B, T, C, _stream())
        return gx, None


class _BoundaryBCEMultiFn(torch.autograd.Function):
    """Several boundary maps at once: losses[i] = mean_{b,t} BCE(mean_c tanh(xs[i][b,t,c]), targets[i][b,t]).  One forward launch per
    map into shared row buffers, ONE launch for all the means; the backward reads its upstream gradients straight from the
    gradient vector (no per-map scalar tensors)."""

    @staticmethod
    def forward(ctx, n, *args):
        xs, targets = list(args[:n]), list(args[n:])
        _require_cuda(*xs, *targets)
        dev = xs[0].device
        rows = [x.shape[0] * x.shape[1] for x in xs]
        offs = [0]
        for r in rows:
            offs.append(offs[-1] + r)
        row_loss = torch.empty(offs[-1], dtype=torch.float32, device=dev)
        coef = torch.empty(offs[-1], dtype=torch.float32, device=dev)
        kept = []
        for i, (x, tg) in enumerate(zip(xs, targets)):
            B, T, C = x.shape
            sliced = (not x.is_contiguous()) and x.stride(2) == 1 and x.stride(0) == T * x.stride(1) and x.stride(1) >= C
            if not sliced:
                x = x.contiguous()
            assert tg.shape == (B, T) and tg.stride(1) == 1 and tg.dtype == torch.float32 and x.dtype == torch.float32
            _lib.call("otal_boundary_bce_fwd_ex", x.data_ptr(), x.stride(1), tg.data_ptr(), tg.stride(0), row_loss.data_ptr() + 4 * offs[i],
                      coef.data_ptr() + 4 * offs[i], B, T, C, _stream())
            kept.append(x)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        arr = (ctypes.c_longlong * (n + 1))(*offs)
        _lib.call("otal_segment_mean", row_loss.data_ptr(), arr, n, out.data_ptr(), _stream())
        ctx.kept, ctx.coef, ctx.offs = kept, coef, offs
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().float()
        grads = []
        for i, x in enumerate(ctx.kept):
            B, T, C = x.shape
            gx = torch.empty((B, T, C), dtype=torch.float32, device=x.device)
            _lib.call("otal_boundary_bce_bwd_ex", x.data_ptr(), x.stride(1), ctx.coef.data_ptr() + 4 * ctx.offs[i], g.data_ptr() + 4 * i,
                      gx.data_ptr(), B, T, C, _stream())
            grads.append(gx)
        ctx.kept = None
        return (None, *grads, *([None] * len(grads)))


def boundary_bce_multi(xs: list[torch.Tensor], targets: list[torch.Tensor]) -> torch.Tensor:
    """[len(xs)] vector of calc_bce_loss terms (thumos14/train.py:152-161), one per (map [B,T,C], target [B,T]) pair."""
    return _BoundaryBCEMultiFn.apply(len(xs), *xs, *targets)


def boundary_bce(x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mean_{b,t} BCE(mean_c tanh(x[b,t,c]), target[b,t]) — calc_bce_loss (thumos14/train.py:152-161) for one map."""
    return _BoundaryBCEFn.apply(x, target)
