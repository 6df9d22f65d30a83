"""I3D backbone (Inception-v1 3-D up to Mixed_5c) on the sm_100a kernels: forward AND backward.

Replaces `I3D_BackBone` / `InceptionI3d.extract_features` (AFSD/thumos14/BDNet.py:25-52,
AFSD/common/i3d_backbone.py:124-342) behind the same module interface and the same state_dict keys
(`_model.<EndPoint>[.<branch>].conv3d.weight`, `.bn.{weight,bias,running_mean,running_var,num_batches_tracked}`).

Design (not a translation of the reference's module-per-op graph):
  * activations live in HBM as NDHWC bf16 (hi, lo) planes and never leave that layout inside the backbone;
    every Unit3D (pad -> conv -> frozen BN -> ReLU, i3d_backbone.py:51-87) is ONE tcgen05 implicit-GEMM launch whose
    TMA zero fill is the padding and whose epilogue applies the folded BN + ReLU and writes the hi/lo planes
    straight into the channel slice of the inception concat buffer (no F.pad copy, no torch.cat copy);
  * all conv weights live in one flat fp32 buffer in the kernels' [tap][Cout][Cin] order; the nn.Parameters are
    strided views of it with the reference's [Cout,Cin,kt,kh,kw] shape, so state_dicts stay interchangeable while
    one split kernel per step produces every bf16 weight plane, the weight-gradient kernels accumulate straight
    into the flat gradient buffer, and the optimizer / all-reduce see one contiguous tensor;
  * backward is an explicit schedule (no autograd graph inside the backbone): ReLU/BN backward fused with the
    hi/lo split, dgrad = the forward kernel reading the weights transposed (MN-major B, flipped taps), wgrad = the
    MN-major tcgen05 kernel, max-pool gradient routing by recomputed argmax.
BatchNorm is frozen and in eval mode even while training (BDNet.py:39-49): only its folded scale/shift is used.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn

from . import ops
from .ops import Planes

BN_EPS = 1e-3   # i3d_backbone.py:43

# (name, kind, args): i3d_backbone.py:193-296
ENDPOINTS = [
    ("Conv3d_1a_7x7", "conv1a", dict(cin=3, cout=64)),
    ("MaxPool3d_2a_3x3", "pool", dict(k=(1, 3, 3), s=(1, 2, 2))),
    ("Conv3d_2b_1x1", "conv", dict(cin=64, cout=64, k=(1, 1, 1))),
    ("Conv3d_2c_3x3", "conv", dict(cin=64, cout=192, k=(3, 3, 3))),
    ("MaxPool3d_3a_3x3", "pool", dict(k=(1, 3, 3), s=(1, 2, 2))),
    ("Mixed_3b", "mixed", dict(cin=192, widths=(64, 96, 128, 16, 32, 32))),
    ("Mixed_3c", "mixed", dict(cin=256, widths=(128, 128, 192, 32, 96, 64))),
    ("MaxPool3d_4a_3x3", "pool", dict(k=(3, 3, 3), s=(2, 2, 2))),
    ("Mixed_4b", "mixed", dict(cin=480, widths=(192, 96, 208, 16, 48, 64))),
    ("Mixed_4c", "mixed", dict(cin=512, widths=(160, 112, 224, 24, 64, 64))),
    ("Mixed_4d", "mixed", dict(cin=512, widths=(128, 128, 256, 24, 64, 64))),
    ("Mixed_4e", "mixed", dict(cin=512, widths=(112, 144, 288, 32, 64, 64))),
    ("Mixed_4f", "mixed", dict(cin=528, widths=(256, 160, 320, 32, 128, 128))),
    ("MaxPool3d_5a_2x2", "pool", dict(k=(2, 2, 2), s=(2, 2, 2))),
    ("Mixed_5b", "mixed", dict(cin=832, widths=(256, 160, 320, 32, 128, 128))),
    ("Mixed_5c", "mixed", dict(cin=832, widths=(384, 192, 384, 48, 128, 128))),
]
BRANCHES = ("b0", "b1a", "b1b", "b2a", "b2b", "b3b")        # registration order, i3d_backbone.py:94-113


def same_pad_front(size: int, k: int, s: int) -> int:
    """Front part of the TF-style "same" padding (i3d_backbone.py:45-69): total//2."""
    total = max(k - s, 0) if size % s == 0 else max(k - size % s, 0)
    return total // 2


def _pads(shape, k, s=(1, 1, 1)):
    return tuple(same_pad_front(sz, kk, ss) for sz, kk, ss in zip(shape, k, s))


class _ConvRec:
    """One Unit3D: where its weight / BN live in the flat buffers."""

    def __init__(self, key, cin, cout, k):
        self.key, self.cin, self.cout, self.k = key, cin, cout, k
        self.taps = k[0] * k[1] * k[2]
        self.numel = self.taps * cin * cout
        self.w_off = -1     # element offset in the flat weight buffer
        self.bn_off = -1    # channel offset in the flat BN buffers


class _Holder(nn.Module):
    """Attribute container that only exists to give parameters the reference's state_dict names."""


class I3DBackbone(nn.Module):
    """Drop-in for `I3D_BackBone` (BDNet.py:25-52): forward(x [N,3,T,H,W] fp32) -> {'Mixed_4f', 'Mixed_5c'} NCDHW fp32.

    precision: 'bf16x3' (default; ~1e-5 relative, inside the 1e-3 budget) or 'bf16' (single pass, ~3e-3)."""

    final_endpoint = "Mixed_5c"

    def __init__(self, in_channels: int = 3, precision: str = "bf16x3", freeze_bn: bool = True,
                 freeze_bn_affine: bool = True):
        super().__init__()
        assert in_channels == 3, "the folded Conv3d_1a kernel is written for RGB clips"
        assert precision in ("bf16x3", "bf16")
        self.precision = precision
        self._freeze_bn = freeze_bn
        self._freeze_bn_affine = freeze_bn_affine
        self.convs: dict[str, _ConvRec] = {}
        self._model = _Holder()
        order: list[_ConvRec] = []
        for name, kind, a in ENDPOINTS:
            if kind == "conv1a":
                order.append(self._add_unit(name, a["cin"], a["cout"], (7, 7, 7)))
            elif kind == "conv":
                order.append(self._add_unit(name, a["cin"], a["cout"], a["k"]))
            elif kind == "mixed":
                w = a["widths"]
                setattr(self._model, name, _Holder())
                recs = {
                    "b0": self._add_unit(f"{name}.b0", a["cin"], w[0], (1, 1, 1)),
                    "b1a": self._add_unit(f"{name}.b1a", a["cin"], w[1], (1, 1, 1)),
                    "b1b": self._add_unit(f"{name}.b1b", w[1], w[2], (3, 3, 3)),
                    "b2a": self._add_unit(f"{name}.b2a", a["cin"], w[3], (1, 1, 1)),
                    "b2b": self._add_unit(f"{name}.b2b", w[3], w[4], (3, 3, 3)),
                    "b3b": self._add_unit(f"{name}.b3b", a["cin"], w[5], (1, 1, 1)),
                }
                # private flat order: the four concat branches first (their BN scales are then contiguous and line
                # up with the concat buffer), then the two bottlenecks (contiguous, line up with the mid buffer)
                order += [recs[b] for b in ("b0", "b1b", "b2b", "b3b", "b1a", "b2a")]
        w_off = bn_off = 0
        for r in order:
            r.w_off, r.bn_off = w_off, bn_off
            w_off += (r.numel + 7) // 8 * 8      # keep every weight block 32-byte aligned (bf16 planes: 16 B, TMA)
            bn_off += r.cout
        self._w_total, self._bn_total = w_off, bn_off
        self._flat_dev = None
        self._anchor = None
        self.on_backward_start = None     # optional callback (the trainer overlaps the head's all-reduce here): fired when
                                          # the LAST pending backbone backward of the step starts (see _BackboneFn)
        self.on_deep_done = None          # optional callback: fired when the weight gradients of Mixed_4b..Mixed_5c (the bulk of the
                                          # parameters: grad_split_offset() onwards in the flat buffer) are complete
        self._pending_bwd = 0             # backbone forwards of this step whose backward has not started yet
        self.crop_size = 96               # uint8 input path: crop extent (config dataset.training.crop_size)
        self.crop_offsets = None          # optional int32 [N,3] device tensor (row, column, mirror) per sample
        self._parked: list = []           # (event, tensors) of blocks whose side-stream weight gradients may still run
        self.frame_map = None             # optional int32 [N,T] device tensor: temporal gather in the ingest kernel (SSL cut-paste)
        # Conv3d_1a fwd + wgrad on the raw uint8 pixel values (uint8 input only) — one exact bf16 plane, one tensor-core pass
        # instead of 2 (fwd) / 3 (wgrad); ops.conv1a_u8_scale_shift / conv1a_u8_weight_grad.  Round-2 GPU A/B: 23.47 -> 22.99 ms
        # per step, parity 9e-6 / 2.4e-5 vs the bf16x3 form (tests/test_conv1a_u8_gpu.py).  OTAL_U8_CONV1A=0 switches it off.
        self.u8_conv1a = os.environ.get("OTAL_U8_CONV1A", "1") != "0"
        # The two bottleneck 1x1 convs of an inception block (b1a, b2a: same input, adjacent weights / BN scales / outputs by
        # construction) as ONE forward launch and ONE weight-gradient launch — b2a alone is a 16..48-channel conv, a
        # launch-latency-bound sliver.  Host-side only: same kernels, same memory layout (-0.18 ms per step).  OTAL_FUSE_B12A=0: off.
        self.fuse_b12a = os.environ.get("OTAL_FUSE_B12A", "1") != "0"
        self.reset_parameters()

    def grad_split_offset(self) -> int:
        """Offset (elements) inside the flat weight / gradient buffer of the first parameter of Mixed_4b: everything from there on
        is complete when `on_deep_done` fires (the buffer is laid out in forward order)."""
        return min(r.w_off for name, r in self.convs.items() if name.startswith("Mixed_4b"))

    # ------------------------------------------------------------------------------------------------ structure
    def _add_unit(self, path: str, cin: int, cout: int, k) -> _ConvRec:
        parent = self._model
        parts = path.split(".")
        for p in parts[:-1]:
            parent = getattr(parent, p)
        unit = _Holder()
        unit.conv3d = _Holder()
        unit.conv3d.register_parameter("weight", nn.Parameter(torch.empty(cout, cin, *k)))
        unit.bn = _Holder()
        unit.bn.register_parameter("weight", nn.Parameter(torch.ones(cout), requires_grad=not self._freeze_bn_affine))
        unit.bn.register_parameter("bias", nn.Parameter(torch.zeros(cout), requires_grad=not self._freeze_bn_affine))
        unit.bn.register_buffer("running_mean", torch.zeros(cout))
        unit.bn.register_buffer("running_var", torch.ones(cout))
        unit.bn.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        setattr(parent, parts[-1], unit)
        rec = _ConvRec(path, cin, cout, tuple(k))
        rec.unit = unit
        self.convs[path] = rec
        return rec

    def reset_parameters(self) -> None:
        """torch's default Conv3d init (kaiming_uniform a=sqrt(5)): the backbone keeps torch defaults unless
        pretrained weights are loaded (BDNet.py:442-444)."""
        for r in self.convs.values():
            nn.init.kaiming_uniform_(r.unit.conv3d.weight, a=math.sqrt(5))

    def train(self, mode: bool = True):
        # BN stays in eval mode with frozen affine parameters (BDNet.py:39-49); nothing else depends on the mode.
        return super().train(mode)

    # ------------------------------------------------------------------------------------------------ flat storage
    def _flatten(self, device) -> None:
        """(Re)build the flat buffers on `device` and re-point every parameter at its strided view."""
        old = {k: (r.unit.conv3d.weight.data, r.unit.bn.weight.data, r.unit.bn.bias.data,
                   r.unit.bn.running_mean.data, r.unit.bn.running_var.data) for k, r in self.convs.items()}
        self.flat_w = torch.zeros(self._w_total, dtype=torch.float32, device=device)
        self.flat_g = torch.zeros(self._w_total, dtype=torch.float32, device=device)
        self.flat_bn = torch.zeros(4, self._bn_total, dtype=torch.float32, device=device)   # weight, bias, mean, var
        for k, r in self.convs.items():
            w, bw, bb, bm, bv = old[k]
            view = self._wview(self.flat_w, r)
            view.copy_(w.to(device))
            r.unit.conv3d.weight.data = view
            r.unit.conv3d.weight.grad = None
            sl = slice(r.bn_off, r.bn_off + r.cout)
            for i, (t, holder, nm) in enumerate(((bw, r.unit.bn, "weight"), (bb, r.unit.bn, "bias"),
                                                 (bm, r.unit.bn, "running_mean"), (bv, r.unit.bn, "running_var"))):
                self.flat_bn[i, sl].copy_(t.to(device))
                getattr(holder, nm).data = self.flat_bn[i, sl]
        self._flat_dev = device
        self._anchor = torch.zeros(1, device=device, requires_grad=True)

    @staticmethod
    def _wview(flat: torch.Tensor, r: _ConvRec) -> torch.Tensor:
        """[Cout,Cin,kt,kh,kw]-shaped strided view of the packed [taps][Cout][Cin] block."""
        return flat[r.w_off:r.w_off + r.numel].view(*r.k, r.cout, r.cin).permute(3, 4, 0, 1, 2)

    def _packed(self, flat: torch.Tensor, r: _ConvRec) -> torch.Tensor:
        return flat[r.w_off:r.w_off + r.numel].view(r.taps, r.cout, r.cin)

    def _ensure_flat(self, device) -> None:
        first = next(iter(self.convs.values()))
        w = first.unit.conv3d.weight
        ok = (self._flat_dev == device and w.device == device
              and w.data_ptr() == self.flat_w.data_ptr() + 4 * first.w_off)
        if ok:
            # cheap spot check of the last parameter too (a partial .to() would break the aliasing)
            last = list(self.convs.values())[-1]
            ok = last.unit.conv3d.weight.data_ptr() == self.flat_w.data_ptr() + 4 * last.w_off
        if not ok:
            self._flatten(device)

    def flat_parameters(self, device=None) -> tuple[torch.Tensor, torch.Tensor]:
        """(flat weights, flat gradients) of all backbone conv weights — what the optimizer / all-reduce use."""
        self._ensure_flat(device or self._flat_dev or torch.device("cuda", torch.cuda.current_device()))
        return self.flat_w, self.flat_g

    def _bind_grads(self) -> None:
        """Make every conv weight's .grad the matching strided view of the flat gradient buffer."""
        for r in self.convs.values():
            p = r.unit.conv3d.weight
            want = self.flat_g.data_ptr() + 4 * r.w_off
            if p.grad is None or p.grad.data_ptr() != want:
                if p.grad is not None:
                    self._wview(self.flat_g, r).copy_(p.grad)
                else:
                    self._wview(self.flat_g, r).zero_()
                p.grad = self._wview(self.flat_g, r)

    # ------------------------------------------------------------------------------------------------ forward
    def _prepare(self):
        """Per-step operand preparation: bf16 weight planes (one split launch over the flat buffer) and folded BN."""
        with_lo = self.precision == "bf16x3"
        self._wp = ops.split_bf16(self.flat_w, with_lo)
        bw, bb, bm, bv = self.flat_bn
        self._scale = bw * torch.rsqrt(bv + BN_EPS)
        self._shift = bb - bm * self._scale
        r = self.convs["Conv3d_1a_7x7"]
        self._w1a = ops.pack_conv1a_weight(r.unit.conv3d.weight, with_lo)
        self._w1a_cat = ops.pack_conv1a_weight_cat(self._w1a) if (with_lo and self.u8_conv1a) else None

    def _w(self, r: _ConvRec) -> Planes:
        sl = slice(r.w_off, r.w_off + r.numel)
        return Planes(self._wp.hi[sl].view(r.taps, r.cout, r.cin),
                      self._wp.lo[sl].view(r.taps, r.cout, r.cin) if self._wp.lo is not None else None)

    def _ss(self, r: _ConvRec, width: int | None = None):
        n = width or r.cout
        return self._scale[r.bn_off:r.bn_off + n], self._shift[r.bn_off:r.bn_off + n]

    @staticmethod
    def _new(shape, like: Planes) -> Planes:
        hi = torch.empty(shape, dtype=torch.bfloat16, device=like.hi.device)
        return Planes(hi, torch.empty_like(hi) if like.lo is not None else None)

    def _conv(self, x: Planes, r: _ConvRec, out: Planes, out_off: int = 0, in_slice=None, out_f32=None):
        sc, sh = self._ss(r)
        ops.conv_igemm(x, self._w(r), kernel=r.k, pad_front=_pads(x.hi.shape[1:4], r.k), scale=sc, shift=sh, relu=True,
                       in_slice=in_slice, out=out, out_slice=(out_off, r.cout), out_f32=out_f32)

    def _w12(self, c: dict) -> Planes:
        """Weight planes of [b1a ; b2a] as ONE [1, w1a + w2a, Cin] operand (adjacent in the flat buffer by construction)."""
        r1, r2 = c["b1a"], c["b2a"]
        assert r1.w_off + r1.numel == r2.w_off and r1.bn_off + r1.cout == r2.bn_off and r1.cin == r2.cin
        sl = slice(r1.w_off, r2.w_off + r2.numel)
        return Planes(self._wp.hi[sl].view(1, r1.cout + r2.cout, r1.cin),
                      self._wp.lo[sl].view(1, r1.cout + r2.cout, r1.cin) if self._wp.lo is not None else None)

    def _mixed(self, name: str, x: Planes, saved: dict, out_f32: bool = False) -> Planes:
        c = {b: self.convs[f"{name}.{b}"] for b in BRANCHES}
        shape = x.hi.shape[:4]
        ctot = c["b0"].cout + c["b1b"].cout + c["b2b"].cout + c["b3b"].cout
        y = self._new((*shape, ctot), x)
        mid = self._new((*shape, c["b1a"].cout + c["b2a"].cout), x)
        f32 = torch.empty((*shape, ctot), dtype=torch.float32, device=x.hi.device) if out_f32 else None
        pooled, parg = ops.maxpool_fwd(x, kernel=(3, 3, 3), stride=(1, 1, 1), pad_front=_pads(shape[1:], (3, 3, 3)),
                                       save_argmax=True)
        o1, o2, o3 = c["b0"].cout, c["b0"].cout + c["b1b"].cout, c["b0"].cout + c["b1b"].cout + c["b2b"].cout
        self._conv(x, c["b0"], y, 0, out_f32=f32)
        if self.fuse_b12a:
            r1, w12 = c["b1a"], self._w12(c)
            n12 = w12.hi.shape[1]
            ops.conv_igemm(x, w12, kernel=(1, 1, 1), pad_front=(0, 0, 0), scale=self._scale[r1.bn_off:r1.bn_off + n12],
                           shift=self._shift[r1.bn_off:r1.bn_off + n12], relu=True, out=mid, out_slice=(0, n12))
        else:
            self._conv(x, c["b1a"], mid, 0)
            self._conv(x, c["b2a"], mid, c["b1a"].cout)
        self._conv(mid, c["b1b"], y, o1, in_slice=(0, c["b1a"].cout), out_f32=f32)
        self._conv(mid, c["b2b"], y, o2, in_slice=(c["b1a"].cout, c["b2a"].cout), out_f32=f32)
        self._conv(pooled, c["b3b"], y, o3, out_f32=f32)
        saved[name] = (x, y, mid, pooled, parg)
        if out_f32:
            saved[name + ".f32"] = f32
        return y

    def forward_planes(self, x: torch.Tensor, saved: dict) -> dict:
        """Runs the whole backbone; `saved` receives every tensor the backward schedule needs."""
        with_lo = self.precision == "bf16x3"
        self._ensure_flat(x.device)
        self._prepare()
        if x.dtype == torch.uint8:
            # the dataset's storage format: uint8 frames [N,T,Hs,Ws,3]; centre crop + normalisation happen in the ingest kernel
            assert x.is_cuda and x.dim() == 5 and x.shape[4] == 3, "expected CUDA uint8 frames [N,T,Hs,Ws,3]"
            W = self.crop_size
            u8 = bool(self.u8_conv1a and with_lo and x.shape[1] % 2 == 0 and x.shape[1] >= 6 and W >= 6)
            a = ops.clip_ingest_u8(x, W, self.crop_offsets, with_lo, frame_map=self.frame_map, raw=u8)
        else:
            assert x.is_cuda and x.dim() == 5 and x.shape[1] == 3, "expected a CUDA clip batch [N,3,T,H,W]"
            W = x.shape[4]
            u8 = False
            a = ops.clip_ingest(x, with_lo)
        saved["clip"] = (a, W, u8)
        r = self.convs["Conv3d_1a_7x7"]
        sc, sh = self._ss(r)
        if u8:
            sc, sh = ops.conv1a_u8_scale_shift(r.unit.conv3d.weight, sc, sh)
        cur = ops.conv1a_fwd(a, self._w1a, W, scale=sc, shift=sh, relu=True, u8=u8, w_cat=self._w1a_cat if u8 else None)
        saved["Conv3d_1a_7x7"] = cur
        for name, kind, arg in ENDPOINTS[1:]:
            if kind == "pool":
                nxt, parg = ops.maxpool_fwd(cur, kernel=arg["k"], stride=arg["s"],
                                            pad_front=_pads(cur.hi.shape[1:4], arg["k"], arg["s"]), save_argmax=True)
                saved[name] = (cur, nxt, parg)
            elif kind == "conv":
                r = self.convs[name]
                nxt = self._new((*cur.hi.shape[:4], r.cout), cur)
                self._conv(cur, r, nxt)
                saved[name] = (cur, nxt)
            else:
                nxt = self._mixed(name, cur, saved, out_f32=name in ("Mixed_4f", "Mixed_5c"))
            cur = nxt
        return {"Mixed_4f": saved["Mixed_4f.f32"], "Mixed_5c": saved["Mixed_5c.f32"]}

    def forward(self, x: torch.Tensor) -> dict:
        self._ensure_flat(x.device)
        f4, f5 = _BackboneFn.apply(self, x, self._anchor)
        # NDHWC storage viewed as the reference's NCDHW
        return {"Mixed_4f": f4.permute(0, 4, 1, 2, 3), "Mixed_5c": f5.permute(0, 4, 1, 2, 3)}

    # ------------------------------------------------------------------------------------------------ backward
    def _conv_bwd(self, r: _ConvRec, x: Planes, d: Planes, g_x: torch.Tensor | None, *, in_slice=None, d_slice=None,
                  gx_off: int = 0, accumulate: bool = False) -> None:
        """wgrad into the flat gradient buffer + (optionally) dgrad into the fp32 buffer g_x."""
        pads = _pads(x.hi.shape[1:4], r.k)
        if ops.OVERLAP_WGRAD:
            # the weight gradient runs on the side stream next to the data gradients / elementwise kernels of the main
            # stream (it fills their tails: every kernel here is persistent with one CTA per SM); the callers join before
            # x / d can be released
            with torch.cuda.stream(ops.fork()):
                ops.conv_wgrad(x, d, self._packed(self.flat_g, r), kernel=r.k, pad_front=pads, in_slice=in_slice, d_slice=d_slice)
        else:
            ops.conv_wgrad(x, d, self._packed(self.flat_g, r), kernel=r.k, pad_front=pads, in_slice=in_slice, d_slice=d_slice)
        if g_x is not None:
            ops.conv_igemm(d, self._w(r), kernel=r.k, pad_front=tuple(kk - 1 - p for kk, p in zip(r.k, pads)),
                           in_slice=d_slice, out_f32=g_x, out_slice=(gx_off, r.cin), want_planes=False, dgrad=True,
                           accumulate=accumulate)

    def _mixed_scale(self, name: str) -> torch.Tensor:
        """Folded BN scales of the four concat branches [b0|b1b|b2b|b3b] of a Mixed block (contiguous by design)."""
        c0 = self.convs[f"{name}.b0"]
        ctot = sum(self.convs[f"{name}.{b}"].cout for b in ("b0", "b1b", "b2b", "b3b"))
        return self._scale[c0.bn_off:c0.bn_off + ctot]

    def _mixed_bwd(self, name: str, saved: dict, g_y: torch.Tensor | None, d_y: Planes | None = None) -> torch.Tensor:
        """g_y: fp32 gradient w.r.t. the block output, or d_y: the already masked / scaled / split gradient planes
        (produced by the fused pool backward of the stage pool that follows the block)."""
        c = {b: self.convs[f"{name}.{b}"] for b in BRANCHES}
        x, y, mid, pooled, parg = saved.pop(name)
        with_lo = self.precision == "bf16x3"
        o1, o2, o3 = c["b0"].cout, c["b0"].cout + c["b1b"].cout, c["b0"].cout + c["b1b"].cout + c["b2b"].cout
        if d_y is None:
            d_y = ops.relu_bn_bwd_split(g_y, y, self._mixed_scale(name), with_lo=with_lo)
        dev = d_y.hi.device
        shape = x.hi.shape[:4]
        g_x = torch.empty((*shape, x.hi.shape[-1]), dtype=torch.float32, device=dev)
        g_mid = torch.empty((*shape, mid.hi.shape[-1]), dtype=torch.float32, device=dev)
        g_pool = torch.empty((*shape, x.hi.shape[-1]), dtype=torch.float32, device=dev)
        w1a, w2a = c["b1a"].cout, c["b2a"].cout
        self._conv_bwd(c["b0"], x, d_y, None, d_slice=(0, c["b0"].cout))                    # weight gradient only
        self._conv_bwd(c["b1b"], mid, d_y, g_mid, in_slice=(0, w1a), d_slice=(o1, c["b1b"].cout), gx_off=0)
        self._conv_bwd(c["b2b"], mid, d_y, g_mid, in_slice=(w1a, w2a), d_slice=(o2, c["b2b"].cout), gx_off=w1a)
        self._conv_bwd(c["b3b"], pooled, d_y, g_pool, d_slice=(o3, c["b3b"].cout))
        sc_m = self._scale[c["b1a"].bn_off:c["b1a"].bn_off + mid.hi.shape[-1]]   # [b1a|b2a] contiguous by design
        d_m = ops.relu_bn_bwd_split(g_mid, mid, sc_m, with_lo=with_lo)
        if self.fuse_b12a:
            r1, r2 = c["b1a"], c["b2a"]
            dw12 = self.flat_g[r1.w_off:r2.w_off + r2.numel].view(1, w1a + w2a, r1.cin)        # both gradient blocks, adjacent
            if ops.OVERLAP_WGRAD:
                with torch.cuda.stream(ops.fork()):
                    ops.conv_wgrad(x, d_m, dw12, kernel=(1, 1, 1), pad_front=(0, 0, 0))
            else:
                ops.conv_wgrad(x, d_m, dw12, kernel=(1, 1, 1), pad_front=(0, 0, 0))
        else:
            self._conv_bwd(c["b1a"], x, d_m, None, d_slice=(0, w1a))
            self._conv_bwd(c["b2a"], x, d_m, None, d_slice=(w1a, w2a))
        # data gradient of the three 1x1 convs that read x (b0, b1a, b2a) in ONE pass: K-concatenated
        # [d_y(b0) | d_m] . [W_b0 ; W_b1a ; W_b2a] — g_x is written once instead of one write + two read-modify-writes
        w12 = self._w12(c)
        ops.conv_igemm(d_y, self._w(c["b0"]), kernel=(1, 1, 1), pad_front=(0, 0, 0), in_slice=(0, c["b0"].cout), out_f32=g_x,
                       out_slice=(0, c["b0"].cin), want_planes=False, dgrad=True, x2=d_m, w2=w12)
        ops.maxpool_bwd(x, g_pool, g_x, kernel=(3, 3, 3), stride=(1, 1, 1), pad_front=_pads(shape[1:], (3, 3, 3)), argmax=parg)
        # the six weight gradients may still be running on the side stream: everything they read stays referenced until the
        # main stream has waited for them (one block later, see _retire)
        self._retire((x, y, mid, pooled, parg, d_y, d_m))
        return g_x

    def _retire(self, tensors) -> None:
        """Deferred join: the side stream's weight gradients of this block keep running under the main stream's next
        kernels (stage-pool backward, ReLU/BN backward, the next block's data gradients).  The tensors they read are
        parked with an event recorded behind them; the PREVIOUS block's entry is released once the main stream has
        waited for its event — so at most two blocks of gradient planes are held beyond their natural lifetime."""
        if not ops.OVERLAP_WGRAD:
            return
        ev = torch.cuda.Event()
        ev.record(ops.side_stream())
        self._parked.append((ev, tensors))
        while len(self._parked) > 1:
            old_ev, _ = self._parked.pop(0)
            torch.cuda.current_stream().wait_event(old_ev)

    def _retire_all(self) -> None:
        ops.join()
        self._parked.clear()

    def backward_planes(self, saved: dict, g4: torch.Tensor | None, g5: torch.Tensor | None) -> None:
        """Explicit backward schedule.  g4 / g5: fp32 NDHWC gradients w.r.t. Mixed_4f / Mixed_5c (or None)."""
        self._bind_grads()
        with_lo = self.precision == "bf16x3"
        dev = self.flat_w.device
        x5c = saved["Mixed_5c"][1]
        g = g5.contiguous() if g5 is not None else torch.zeros(x5c.hi.shape, dtype=torch.float32, device=dev)
        g4 = g4.contiguous() if g4 is not None else None          # added inside the fused backward of MaxPool3d_5a
        # A stage pool's backward is fused with the ReLU / BN backward of the layer in front of it (gather kernel): it
        # hands that layer its gradient planes `d_next` directly, no fp32 gradient of the pool input is materialised.
        names = [e[0] for e in ENDPOINTS]
        d_next = None
        for name, kind, arg in reversed(ENDPOINTS):
            if kind == "mixed":
                # (Mixed_4f: its gradient planes come from the fused MaxPool3d_5a backward, which already added g4)
                g = self._mixed_bwd(name, saved, g, d_next)
                d_next = None
                if name == "Mixed_4b" and self.on_deep_done is not None and self._pending_bwd == 0:
                    self._retire_all()           # joins the side stream: every weight gradient launched so far has been ordered
                    self.on_deep_done()
            elif kind == "pool":
                x, _, parg = saved.pop(name)
                prev_name, prev_kind, _ = ENDPOINTS[names.index(name) - 1]
                if prev_kind == "mixed":
                    sc = self._mixed_scale(prev_name)
                else:
                    sc, _ = self._ss(self.convs[prev_name])
                d_next = ops.maxpool_bwd_relu_bn_split(x, parg, g, sc, kernel=arg["k"], stride=arg["s"],
                                                       pad_front=_pads(x.hi.shape[1:4], arg["k"], arg["s"]),
                                                       g_add=g4 if name == "MaxPool3d_5a_2x2" else None, with_lo=with_lo)
                g = None
            elif kind == "conv":
                r = self.convs[name]
                x, y = saved.pop(name)
                if d_next is None:
                    sc, _ = self._ss(r)
                    d_next = ops.relu_bn_bwd_split(g, y, sc, with_lo=with_lo)
                d, d_next = d_next, None
                g = torch.empty((*x.hi.shape[:4], r.cin), dtype=torch.float32, device=dev)
                self._conv_bwd(r, x, d, g)
                self._retire((x, y, d))
            else:  # conv1a: weight gradient only, the clip needs no gradient (train.py:165)
                r = self.convs[name]
                y = saved.pop(name)
                a, W, u8 = saved.pop("clip")
                if d_next is None:
                    sc, _ = self._ss(r)
                    d_next = ops.relu_bn_bwd_split(g, y, sc, with_lo=with_lo)
                d, d_next = d_next, None
                dw = torch.zeros(49, r.cout, 8 * ops.CLIP_CPAD, dtype=torch.float32, device=dev)
                ops.conv1a_wgrad(a, d, dw, W, u8=u8)
                # packed layout of this block is [kt,kh,kw,Cout,Cin]; the parameter's .grad is its strided view
                if u8:
                    r.unit.conv3d.weight.grad.add_(ops.conv1a_u8_weight_grad(dw, None, r.cin))
                else:
                    r.unit.conv3d.weight.grad.add_(ops.unpack_conv1a_wgrad(dw, r.cin))
        self._retire_all()


class _BackboneFn(torch.autograd.Function):
    """Autograd boundary of the native backbone.  Weight gradients are accumulated by the kernels directly into the
    flat gradient buffer that the parameters' .grad alias, so the Function itself only returns None."""

    @staticmethod
    def forward(ctx, net: I3DBackbone, x: torch.Tensor, anchor: torch.Tensor):
        saved: dict = {}
        out = net.forward_planes(x, saved)
        ctx.net = net
        ctx.saved = saved if any(ctx.needs_input_grad) else None
        if ctx.saved is not None:
            net._pending_bwd += 1
        f4, f5 = out["Mixed_4f"], out["Mixed_5c"]
        return f4, f5

    @staticmethod
    def backward(ctx, g4, g5):
        saved = ctx.saved
        if saved is None:
            raise RuntimeError("backbone activations were not saved")
        ctx.saved = None
        saved.pop("Mixed_4f.f32", None)
        saved.pop("Mixed_5c.f32", None)
        # A step with the SSL pass has TWO backbone nodes; autograd runs the later-created one (the SSL branch) first, while
        # the main pass's head backward has not run yet.  Only when the last pending node starts has everything downstream
        # of the backbone — both passes — accumulated its gradient.
        ctx.net._pending_bwd = max(ctx.net._pending_bwd - 1, 0)
        if ctx.net.on_backward_start is not None and ctx.net._pending_bwd == 0:
            ctx.net.on_backward_start()
        ctx.net.backward_planes(saved, g4, g5)
        return None, None, None
