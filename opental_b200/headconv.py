"""Head convolutions (Unit1D, head-side Unit3D) on the tcgen05 implicit-GEMM kernels, forward and backward.

Replaces `nn.Conv1d` / `nn.Conv3d` + `F.pad` inside `Unit1D.forward` (AFSD/common/layers.py:204-214) and the
'spatial_valid' `Unit3D.forward` (layers.py:143-175) for the 26 + 2 conv modules of CoarsePyramid
(AFSD/thumos14/BDNet.py:117-293), about 123 calls per forward.

  * A conv1d over [B,C,T] is the H = W = 1 case of the NDHWC implicit GEMM: the input is transposed + split into
    channels-last bf16 planes by one small kernel, the TMA zero fill supplies the "same" padding (stride-2 convs read
    parity views), the epilogue adds the bias and stores fp32 directly in the reference's [B,C,T] layout
    (coalesced along T).  The (1,6,6)/(1,3,3) full-extent spatial convs of the pyramid are 1x1 convs over
    Cin = kh*kw*C contiguous channels-last values — the backbone's NDHWC feature map is consumed without any copy.
  * dgrad = the same kernel reading the forward weights transposed (MN-major B, flipped taps); a stride-2 conv's
    dgrad runs on the zero-upsampled output gradient.  wgrad = the MN-major tcgen05 kernel accumulating straight into
    the flat gradient buffer the parameters' .grad alias.
  * All head conv weights live in ONE flat fp32 buffer in kernel order [tap][Cout_pad][Cin]; the nn.Parameters are
    strided views with the reference's shapes ([Cout,Cin,k], [Cout,C,1,kh,kw]); one split launch per forward makes
    every bf16 weight plane.  Cout is padded to a multiple of 8 (loc/conf/actionness heads: 2, 15, 1 channels).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .ops import Planes


def _same_pad_front(size: int, k: int, s: int) -> int:
    total = max(k - s, 0) if size % s == 0 else max(k - size % s, 0)
    return total // 2


class _Rec:
    def __init__(self, weight: nn.Parameter, bias: nn.Parameter | None, kind: str):
        self.weight, self.bias, self.kind = weight, bias, kind
        if kind == "conv1d":
            self.cout, self.cin, self.taps = weight.shape
            self.spatial = None
        else:  # 'valid3d': [Cout, C, 1, kh, kw] treated as one tap over kh*kw*C channels-last values
            self.cout, c, kt, kh, kw = weight.shape
            assert kt == 1
            self.cin, self.taps, self.spatial = kh * kw * c, 1, (kh, kw, c)
        self.cpad = (self.cout + 7) // 8 * 8
        self.numel = self.taps * self.cpad * self.cin
        self.off = -1


class HeadConvStore:
    """Flat packed storage of every head conv weight (+ gradients, + per-forward bf16 planes)."""

    def __init__(self, precision: str = "bf16x3"):
        self.recs: list[_Rec] = []
        self.by_id: dict[int, _Rec] = {}
        self.total = 0
        self.dev = None
        self.with_lo = precision == "bf16x3"
        self.planes: Planes | None = None

    def register(self, weight, bias, kind) -> _Rec:
        r = _Rec(weight, bias, kind)
        r.off = self.total
        self.total += (r.numel + 7) // 8 * 8
        self.recs.append(r)
        self.by_id[id(weight)] = r
        return r

    def _view(self, flat: torch.Tensor, r: _Rec) -> torch.Tensor:
        blk = flat[r.off:r.off + r.numel]
        if r.kind == "conv1d":
            return blk.view(r.taps, r.cpad, r.cin)[:, :r.cout].permute(1, 2, 0)          # [Cout,Cin,k]
        kh, kw, c = r.spatial
        return blk.view(r.cpad, kh, kw, c)[:r.cout].permute(0, 3, 1, 2).unsqueeze(2)      # [Cout,C,1,kh,kw]

    def block(self, flat: torch.Tensor, r: _Rec) -> torch.Tensor:
        return flat[r.off:r.off + r.numel].view(r.taps, r.cpad, r.cin)

    def ensure(self, device) -> None:
        first, last = self.recs[0], self.recs[-1]
        ok = self.dev == device and all(
            r.weight.device == device and r.weight.data_ptr() == self.flat_w.data_ptr() + 4 * r.off for r in (first, last))
        if ok:
            return
        self.flat_w = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.flat_g = torch.zeros(self.total, dtype=torch.float32, device=device)
        for r in self.recs:
            v = self._view(self.flat_w, r)
            v.copy_(r.weight.data.to(device))
            r.weight.data = v
            r.weight.grad = None
        self.dev = device

    def bind_grads(self) -> None:
        for r in self.recs:
            p = r.weight
            if not p.requires_grad:
                continue
            v = self._view(self.flat_g, r)
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                if p.grad is not None:
                    v.copy_(p.grad)
                else:
                    v.zero_()
                p.grad = v

    def prepare(self, device) -> None:
        """Once per forward: (re)establish the aliasing and refresh the bf16 weight planes."""
        self.ensure(device)
        self.planes = ops.split_bf16(self.flat_w, self.with_lo)

    def w(self, r: _Rec) -> Planes:
        sl = slice(r.off, r.off + r.numel)
        shape = (r.taps, r.cpad, r.cin)
        return Planes(self.planes.hi[sl].view(shape), self.planes.lo[sl].view(shape) if self.planes.lo is not None else None)


class _HeadConvFn(torch.autograd.Function):
    """y[B,Cout,To] = conv1d_same(x[B,Cin,T], w, stride) + bias  — or, for kind 'valid3d', x is the channels-last
    feature map [B,T,kh,kw,C] and y[B,Cout,T] the collapsed full-extent spatial conv."""

    @staticmethod
    def forward(ctx, x, weight, bias, store: HeadConvStore, rec: _Rec, stride: int):
        B = x.shape[0]
        if rec.kind == "conv1d":
            T = x.shape[2]
            xp = ops.ncl_to_nlc_planes(x, with_lo=store.with_lo)
        else:
            T = x.shape[1]
            x = x.contiguous()
            xp = ops.split_bf16(x.view(B, T, 1, 1, rec.cin), store.with_lo)
        k = rec.taps
        pf = _same_pad_front(T, k, stride)
        To = -(-T // stride)
        y = torch.empty((B, rec.cpad, To), dtype=torch.float32, device=x.device)
        fused_bias = bias is not None and rec.cpad == rec.cout
        ops.conv_igemm(xp, store.w(rec), kernel=(k, 1, 1), pad_front=(pf, 0, 0), stride=(stride, 1, 1),
                       shift=bias.detach() if fused_bias else None, out_f32=y, want_planes=False, f32_ncdhw=True)
        ctx.store, ctx.rec, ctx.stride, ctx.xp, ctx.geom = store, rec, stride, xp, (B, T, To, pf, tuple(x.shape))
        ctx.has_bias = bias is not None
        out = y if rec.cpad == rec.cout else y[:, :rec.cout]
        if bias is not None and not fused_bias:
            out = out + bias.detach().view(1, -1, 1)
        return out

    @staticmethod
    def backward(ctx, gy):
        store, rec, stride, xp = ctx.store, ctx.rec, ctx.stride, ctx.xp
        B, T, To, pf, xshape = ctx.geom
        k = rec.taps
        gy = gy.contiguous()
        dp = ops.ncl_to_nlc_planes(gy, rec.cpad, with_lo=store.with_lo)
        forked = False
        if rec.weight.requires_grad:
            store.bind_grads()
            if ops.OVERLAP_WGRAD and ctx.needs_input_grad[0]:
                # weight gradient on the side stream, next to the data gradient (both are far smaller than the GPU)
                forked = True
                with torch.cuda.stream(ops.fork()):
                    ops.conv_wgrad(xp, dp, store.block(store.flat_g, rec), kernel=(k, 1, 1), pad_front=(pf, 0, 0),
                                   stride=(stride, 1, 1))
            else:
                ops.conv_wgrad(xp, dp, store.block(store.flat_g, rec), kernel=(k, 1, 1), pad_front=(pf, 0, 0), stride=(stride, 1, 1))
        dp_w = dp                # keep the planes the side stream reads alive until the join
        gx = None
        if ctx.needs_input_grad[0]:
            if stride != 1:      # dgrad of a strided conv = stride-1 dgrad of the zero-upsampled gradient
                dp = ops.ncl_to_nlc_planes(gy, rec.cpad, ttot=T, dilate=stride, with_lo=store.with_lo)
            if rec.kind == "conv1d":
                gx = torch.empty(xshape, dtype=torch.float32, device=gy.device)
                ops.conv_igemm(dp, store.w(rec), kernel=(k, 1, 1), pad_front=(k - 1 - pf, 0, 0), out_f32=gx, want_planes=False,
                               dgrad=True, f32_ncdhw=True)
            else:
                gx = torch.empty((B, T, 1, 1, rec.cin), dtype=torch.float32, device=gy.device)
                ops.conv_igemm(dp, store.w(rec), kernel=(1, 1, 1), pad_front=(0, 0, 0), out_f32=gx, want_planes=False, dgrad=True)
                gx = gx.view(xshape)
        gb = gy.sum(dim=(0, 2)) if ctx.has_bias and ctx.needs_input_grad[2] else None
        if forked:
            ops.join()
        del dp_w
        return gx, None, gb, None, None, None


def head_conv(x, weight, bias, store: HeadConvStore, rec: _Rec, stride: int = 1):
    return _HeadConvFn.apply(x, weight, bias, store, rec, stride)
