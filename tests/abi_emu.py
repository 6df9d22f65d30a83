"""A CPU stand-in for the training-path entry points of libopental_b200.so  —  TEST INFRASTRUCTURE (never imported by the product).

`install(monkeypatch)` replaces `opental_b200._lib.call` by a dispatcher that implements, with torch CPU ops on the host
memory behind the raw pointers, what include/opental_b200.h says each entry point does (NDHWC bf16 hi/lo planes, channel
slices of wider rows, tap-major weights, fused BN/ReLU epilogue, data-gradient mode, second K segment, arg-max recording
pools, the fused pool / ReLU / BN backward, clip ingest, the folded Conv3d_1a and its staged raw-uint8 form).  With it the
product's HOST code — opental_b200/backbone.py's forward and explicit backward schedule and the descriptor building in
opental_b200/ops.py — runs on a CPU-only box and can be compared with the oracle's I3D (tests/test_backbone_emulated_cpu.py).
Arithmetic: planes are read as hi + lo in fp32, the contraction runs in fp32 (torch), results are split back into planes —
the accuracy class of the bf16x3 kernels.  Nothing here is a fallback: the product still raises without a CUDA device."""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.nn.functional as F


def _view(ptr, n: int, np_dtype) -> torch.Tensor | None:
    """Flat torch view of n elements of host memory at `ptr` (shares storage)."""
    if not ptr:
        return None
    ptr = ptr.value if isinstance(ptr, ctypes.c_void_p) else int(ptr)
    buf = (ctypes.c_char * (n * np.dtype(np_dtype).itemsize)).from_address(ptr)
    return torch.from_numpy(np.frombuffer(buf, dtype=np_dtype))


def _bf16(ptr, n):          # bf16 plane as an int16 view; .view(torch.bfloat16) reinterprets
    v = _view(ptr, n, np.int16)
    return None if v is None else v.view(torch.bfloat16)


def _load(hi_ptr, lo_ptr, shape, coff, C) -> torch.Tensor:
    """fp32 [.., C] = (hi + lo)[..., coff:coff+C] of planes with rows of shape[-1] channels."""
    n = int(np.prod(shape))
    x = _bf16(hi_ptr, n).view(*shape)[..., coff:coff + C].float()
    if lo_ptr:
        x = x + _bf16(lo_ptr, n).view(*shape)[..., coff:coff + C].float()
    return x


def _store(x: torch.Tensor, hi_ptr, lo_ptr, shape, coff) -> None:
    n = int(np.prod(shape))
    C = x.shape[-1]
    hi = x.bfloat16()
    _bf16(hi_ptr, n).view(*shape)[..., coff:coff + C] = hi
    if lo_ptr:
        _bf16(lo_ptr, n).view(*shape)[..., coff:coff + C] = (x - hi.float()).bfloat16()


def _out(n, s):
    return -(-n // s)


def _conv(x, w, stride, pad_front):
    """x [N,T,H,W,Cin] fp32, w [Cout,Cin,kt,kh,kw]; output extent ceil(in / stride), implicit zero padding behind."""
    xin = x.permute(0, 4, 1, 2, 3)
    k = w.shape[2:]
    pads = []
    for dim in (2, 1, 0):                                           # F.pad order: W, H, T
        n, s, kk, pf = xin.shape[2 + dim], stride[dim], k[dim], pad_front[dim]
        back = max((_out(n, s) - 1) * s + kk - n - pf, 0)
        pads += [pf, back]
    y = F.conv3d(F.pad(xin, pads), w, stride=stride)
    return y[:, :, :_out(x.shape[1], stride[0]), :_out(x.shape[2], stride[1]), :_out(x.shape[3], stride[2])].permute(0, 2, 3, 4, 1)


class Emulator:
    def __init__(self):
        self.calls: dict[str, int] = {}

    def __call__(self, name: str, *args) -> None:
        self.calls[name] = self.calls.get(name, 0) + 1
        fn = getattr(self, name, None)
        if fn is None:
            raise NotImplementedError(f"abi_emu: {name} is not emulated")
        fn(*args)

    # ---------------------------------------------------------------------------------------------- layout kernels
    def otal_split_bf16(self, x, hi, lo, n, stream):
        v = _view(x, n, np.float32)
        _store(v.view(n, 1), hi, lo, (n, 1), 0)

    def otal_merge_bf16(self, hi, lo, x, n, stream):
        _view(x, n, np.float32).copy_(_load(hi, lo, (n, 1), 0, 1).view(n))

    def otal_clip_ingest(self, x, hi, lo, N, C, T, H, W, stream):
        v = _view(x, N * C * T * H * W, np.float32).view(N, C, T, H, W)
        out = torch.zeros(N, T, H, W + 8, 4)
        out[:, :, :, 2:W + 2, :C] = v.permute(0, 2, 3, 4, 1)
        _store(out, hi, lo, (N, T, H, W + 8, 4), 0)

    def _ingest_pixels(self, px, crop, fmap, N, T, Hs, Ws, H, W):
        p = _view(px, N * T * Hs * Ws * 3, np.uint8).view(N, T, Hs, Ws, 3)
        cr = _view(crop, N * 3, np.int32).view(N, 3) if crop else None
        fm = _view(fmap, N * T, np.int32).view(N, T) if fmap else None
        out = torch.zeros(N, T, H, W, 3)
        for n in range(N):
            oh, ow, flip = (int(v) for v in cr[n]) if cr is not None else ((Hs - H) // 2, (Ws - W) // 2, 0)
            src = p[n] if fm is None else p[n][fm[n].long().clamp(0, T - 1)]
            win = src[:, oh:oh + H, ow:ow + W, :].float()
            out[n] = win.flip(2) if flip else win
        return out

    def otal_clip_ingest_u8(self, px, crop, fmap, hi, lo, N, T, Hs, Ws, H, W, stream):
        u = self._ingest_pixels(px, crop, fmap, N, T, Hs, Ws, H, W)
        out = torch.zeros(N, T, H, W + 8, 4)
        out[:, :, :, 2:W + 2, :3] = (u / 255.0) * 2.0 - 1.0
        _store(out, hi, lo, (N, T, H, W + 8, 4), 0)

    def otal_clip_ingest_u8_raw(self, px, crop, fmap, out_ptr, N, T, Hs, Ws, H, W, stream):
        out = torch.zeros(N, T, H, W + 8, 4)
        out[:, :, :, 2:W + 2, :3] = self._ingest_pixels(px, crop, fmap, N, T, Hs, Ws, H, W)
        out[:, :, :, 2:W + 2, 3] = 1.0                                   # the ones slot (in-image indicator)
        _store(out, out_ptr, None, (N, T, H, W + 8, 4), 0)

    # ---------------------------------------------------------------------------------------------- Conv3d_1a (folded)
    @staticmethod
    def _w1a(d, Cout):
        """[Cout,3,7,7,7] from the folded planes [49][Cout][32] (element dw*4 + c)."""
        w = _load(d.w_hi, d.w_lo, (49, Cout, 32), 0, 32).view(7, 7, Cout, 8, 4)
        return w[:, :, :, :7, :3].permute(2, 4, 0, 1, 3).contiguous()

    @staticmethod
    def _classes(n):
        c = torch.zeros(n, dtype=torch.long)
        c[0], c[n - 2], c[n - 1] = 1, 2, 3
        return c

    def _conv1a_fwd(self, desc, u8, w=None):
        d = desc._obj
        N, T, H, W, Cout = d.N, d.T, d.H, d.W, d.Cout
        x = _load(d.x_hi, None if u8 else d.x_lo, (N, T, H, W + 8, 4), 0, 4)[:, :, :, 2:W + 2, :3]
        pf = tuple(2 if n % 2 == 0 else 3 for n in (T, H, W))
        y = _conv(x, self._w1a(d, Cout) if w is None else w, (2, 2, 2), pf)
        To, Ho, Wo = y.shape[1:4]
        sc = _view(d.scale, Cout, np.float32) if d.scale else torch.ones(Cout)
        if u8:
            tab = _view(d.shift, 64 * Cout, np.float32).view(4, 4, 4, Cout)
            sh = tab[self._classes(To)][:, self._classes(Ho)][:, :, self._classes(Wo)]            # [To,Ho,Wo,Cout]
        else:
            sh = _view(d.shift, Cout, np.float32) if d.shift else torch.zeros(Cout)
        y = y * sc + sh
        if d.relu:
            y = y.relu()
        _store(y, d.y_hi, d.y_lo, (N, To, Ho, Wo, d.out_cstride), d.out_coff)

    def otal_conv1a_fwd(self, desc, stream):
        self._conv1a_fwd(desc, False)

    def otal_conv1a_fwd_u8(self, desc, stream):
        self._conv1a_fwd(desc, True)

    def otal_conv1a_fwd_u8_halo(self, desc, stream):
        """Same operator; w_hi = packed [49][2*Cout][32] (hi rows, then lo rows per tap)."""
        d = desc._obj
        Cout = d.Cout
        wc = _bf16(d.w_hi, 49 * 2 * Cout * 32).view(49, 2 * Cout, 32).float()
        w = (wc[:, :Cout] + wc[:, Cout:]).view(7, 7, Cout, 8, 4)[:, :, :, :7, :3].permute(2, 4, 0, 1, 3).contiguous()
        self._conv1a_fwd(desc, True, w=w)

    def _conv1a_wgrad(self, desc, u8):
        d = desc._obj
        N, T, H, W, Cout = d.N, d.T, d.H, d.W, d.Cout
        To, Ho, Wo = _out(T, 2), _out(H, 2), W // 2
        x = _load(d.x_hi, None if u8 else d.x_lo, (N, T, H, W + 8, 4), 0, 4)[:, :, :, :, :]        # padded rows, 4 slots
        g = _load(d.d_hi, d.d_lo, (N, To, Ho, Wo, d.d_cstride), d.d_coff, Cout)
        # gradient against the folded weights: window of 8 pixels x 4 slots starting at padded column 2*w'
        w = torch.zeros(Cout, 4, 7, 7, 8, requires_grad=True)
        xin = x.permute(0, 4, 1, 2, 3)                                                            # [N,4,T,H,W+8]
        pf_t, pf_h = (2 if T % 2 == 0 else 3), (2 if H % 2 == 0 else 3)
        bt = max((To - 1) * 2 + 7 - T - pf_t, 0)
        bh = max((Ho - 1) * 2 + 7 - H - pf_h, 0)
        with torch.enable_grad():          # the product calls this from inside an autograd backward
            y = F.conv3d(F.pad(xin, [0, 0, pf_h, bh, pf_t, bt]), w, stride=2)[:, :, :To, :Ho, :Wo]
            (gw,) = torch.autograd.grad(y, w, g.permute(0, 4, 1, 2, 3))
        dw = _view(d.dw, 49 * Cout * 32, np.float32).view(7, 7, Cout, 8, 4)
        dw += gw.permute(2, 3, 0, 4, 1)

    def otal_conv1a_wgrad(self, desc, stream):
        self._conv1a_wgrad(desc, False)

    def otal_conv1a_wgrad_u8(self, desc, stream):
        self._conv1a_wgrad(desc, True)

    def otal_conv1a_wgrad_u8_halo(self, desc, stream):
        self._conv1a_wgrad(desc, True)

    def otal_border_class_sums(self, d_hi, d_lo, sums, N, To, Ho, Wo, C, cstride, coff, stream):
        g = _load(d_hi, d_lo, (N, To, Ho, Wo, cstride), coff, C).sum(0)
        out = _view(sums, 64 * C, np.float32).view(4, 4, 4, C)
        ct, ch, cw = self._classes(To), self._classes(Ho), self._classes(Wo)
        for a in range(4):
            for b in range(4):
                for c in range(4):
                    sel = (ct == a)[:, None, None] & (ch == b)[None, :, None] & (cw == c)[None, None, :]
                    if sel.any():
                        out[a, b, c] += g[sel].sum(0)

    # ---------------------------------------------------------------------------------------------- generic conv
    def otal_conv_igemm_fwd(self, desc, stream):
        d = desc._obj
        N, T, H, W, Cin, Cout = d.N, d.T, d.H, d.W, d.Cin, d.Cout
        k = (d.kt, d.kh, d.kw)
        taps = k[0] * k[1] * k[2]
        st = tuple(s or 1 for s in (d.sT, d.sH, d.sW))
        assert d.ksplit <= 1, "split-K was withdrawn (the library rejects it)"
        x = _load(d.x_hi, d.x_lo if d.nsplit == 3 else None, (N, T, H, W, d.in_cstride), d.in_coff, Cin)
        lo = d.w_lo if d.nsplit == 3 else None
        if d.dgrad:
            # w = FORWARD weights [taps][Cin (= fwd Cout)][Cout (= fwd Cin)], used transposed with flipped taps
            wf = _load(d.w_hi, lo, (taps, Cin, Cout), 0, Cout).view(*k, Cin, Cout)
            w = wf.flip(0, 1, 2).permute(4, 3, 0, 1, 2).contiguous()
        else:
            w = _load(d.w_hi, lo, (taps, Cout, Cin), 0, Cin).view(*k, Cout, Cin).permute(3, 4, 0, 1, 2).contiguous()
        y = _conv(x, w, st, (d.pt, d.ph, d.pw))
        if d.Cin2 > 0:
            x2 = _load(d.x2_hi, d.x2_lo if d.nsplit == 3 else None, (N, T, H, W, d.in2_cstride), d.in2_coff, d.Cin2)
            lo2 = d.w2_lo if d.nsplit == 3 else None
            w2 = (_load(d.w2_hi, lo2, (1, d.Cin2, Cout), 0, Cout)[0].t() if d.dgrad else _load(d.w2_hi, lo2, (1, Cout, d.Cin2), 0, d.Cin2)[0])
            y = y + x2 @ w2.t()
        To, Ho, Wo = y.shape[1:4]
        if d.scale:
            y = y * _view(d.scale, Cout, np.float32)
        if d.shift:
            y = y + _view(d.shift, Cout, np.float32)
        if d.relu:
            y = y.relu()
        if d.y_hi:
            _store(y, d.y_hi, d.y_lo if d.nsplit == 3 else None, (N, To, Ho, Wo, d.out_cstride), d.out_coff)
        if d.y_f32:
            if d.y_f32_ncdhw:
                dst = _view(d.y_f32, N * d.out_cstride * To * Ho * Wo, np.float32).view(N, d.out_cstride, To, Ho, Wo)
                sl = dst[:, d.out_coff:d.out_coff + Cout]
                yv = y.permute(0, 4, 1, 2, 3)
            else:
                dst = _view(d.y_f32, N * To * Ho * Wo * d.out_cstride, np.float32).view(N, To, Ho, Wo, d.out_cstride)
                sl = dst[..., d.out_coff:d.out_coff + Cout]
                yv = y
            if d.accumulate:
                sl += yv
            else:
                sl.copy_(yv)

    def otal_conv_wgrad(self, desc, stream):
        d = desc._obj
        N, T, H, W, Cin, Cout = d.N, d.T, d.H, d.W, d.Cin, d.Cout
        k = (d.kt, d.kh, d.kw)
        st = tuple(s or 1 for s in (d.sT, d.sH, d.sW))
        To, Ho, Wo = (_out(n, s) for n, s in zip((T, H, W), st))
        split = d.nsplit == 3
        x = _load(d.x_hi, d.x_lo if split else None, (N, T, H, W, d.x_cstride), d.x_coff, Cin)
        g = _load(d.d_hi, d.d_lo if split else None, (N, To, Ho, Wo, d.d_cstride), d.d_coff, Cout)
        w = torch.zeros(Cout, Cin, *k, requires_grad=True)
        with torch.enable_grad():
            (gw,) = torch.autograd.grad(_conv(x, w, st, (d.pt, d.ph, d.pw)), w, g)
        dw = _view(d.dw, k[0] * k[1] * k[2] * Cout * Cin, np.float32).view(*k, Cout, Cin)
        dw += gw.permute(2, 3, 4, 0, 1)

    # ---------------------------------------------------------------------------------------------- pools, ReLU / BN backward
    @staticmethod
    def _pool_geometry(d):
        k, s, pf = (d.kt, d.kh, d.kw), (d.st, d.sh, d.sw), (d.pt, d.ph, d.pw)
        ext = (d.T, d.H, d.W)
        out = tuple(_out(n, ss) for n, ss in zip(ext, s))
        back = tuple(max((o - 1) * ss + kk - n - p, 0) for o, ss, kk, n, p in zip(out, s, k, ext, pf))
        return k, s, pf, ext, out, back

    def otal_maxpool_fwd(self, desc, stream):
        d = desc._obj
        k, s, pf, (T, H, W), (To, Ho, Wo), back = self._pool_geometry(d)
        x = _load(d.x_hi, d.x_lo, (d.N, T, H, W, d.in_cstride), d.in_coff, d.C).permute(0, 4, 1, 2, 3)
        xp = F.pad(x, [pf[2], back[2], pf[1], back[1], pf[0], back[0]])                 # zero padding competes as 0
        y, idx = F.max_pool3d(xp, k, s, return_indices=True)
        y, idx = y[:, :, :To, :Ho, :Wo], idx[:, :, :To, :Ho, :Wo]
        _store(y.permute(0, 2, 3, 4, 1), d.y_hi, d.y_lo, (d.N, To, Ho, Wo, d.out_cstride), d.out_coff)
        if d.argmax:
            Hp, Wp = xp.shape[3], xp.shape[4]
            it, ih, iw = idx // (Hp * Wp), (idx // Wp) % Hp, idx % Wp
            ot = torch.arange(To).view(1, 1, -1, 1, 1) * s[0]
            oh = torch.arange(Ho).view(1, 1, 1, -1, 1) * s[1]
            ow = torch.arange(Wo).view(1, 1, 1, 1, -1) * s[2]
            tap = ((it - ot) * k[1] + (ih - oh)) * k[2] + (iw - ow)                     # window-relative, < 27
            _view(d.argmax, d.N * To * Ho * Wo * d.C, np.uint8).view(d.N, To, Ho, Wo, d.C).copy_(tap.permute(0, 2, 3, 4, 1).to(torch.uint8))

    def _pool_scatter(self, d, g_out):
        """fp32 [N,T,H,W,C]: g_out routed to the recorded arg-max positions (padding positions drop their gradient)."""
        k, s, pf, (T, H, W), (To, Ho, Wo), back = self._pool_geometry(d)
        assert d.argmax, "abi_emu: pool backward is emulated through the recorded arg-max only"
        tap = _view(d.argmax, d.N * To * Ho * Wo * d.C, np.uint8).view(d.N, To, Ho, Wo, d.C).long()
        dt, dh, dw = tap // (k[1] * k[2]), (tap // k[2]) % k[1], tap % k[2]
        t = torch.arange(To).view(1, -1, 1, 1, 1) * s[0] + dt - pf[0]
        h = torch.arange(Ho).view(1, 1, -1, 1, 1) * s[1] + dh - pf[1]
        w = torch.arange(Wo).view(1, 1, 1, -1, 1) * s[2] + dw - pf[2]
        ok = (t >= 0) & (t < T) & (h >= 0) & (h < H) & (w >= 0) & (w < W)
        n = torch.arange(d.N).view(-1, 1, 1, 1, 1).expand_as(tap)
        c = torch.arange(d.C).view(1, 1, 1, 1, -1).expand_as(tap)
        flat = (((n * T + t) * H + h) * W + w) * d.C + c
        out = torch.zeros(d.N * T * H * W * d.C)
        out.index_add_(0, flat[ok], g_out[ok])
        return out.view(d.N, T, H, W, d.C)

    def otal_maxpool_bwd(self, desc, stream):
        d = desc._obj
        k, s, pf, (T, H, W), (To, Ho, Wo), back = self._pool_geometry(d)
        g_out = _view(d.g_out, d.N * To * Ho * Wo * d.gout_cstride, np.float32).view(d.N, To, Ho, Wo, d.gout_cstride)
        g_in = _view(d.g_in, d.N * T * H * W * d.gin_cstride, np.float32).view(d.N, T, H, W, d.gin_cstride)
        g_in[..., d.gin_coff:d.gin_coff + d.C] += self._pool_scatter(d, g_out[..., d.gout_coff:d.gout_coff + d.C])

    def otal_maxpool_bwd_relu_bn_split(self, desc, g_add, add_cstride, add_coff, scale, d_hi, d_lo, d_cstride, d_coff, stream):
        d = desc._obj
        k, s, pf, (T, H, W), (To, Ho, Wo), back = self._pool_geometry(d)
        g_out = _view(d.g_out, d.N * To * Ho * Wo * d.gout_cstride, np.float32).view(d.N, To, Ho, Wo, d.gout_cstride)
        g = self._pool_scatter(d, g_out[..., d.gout_coff:d.gout_coff + d.C])
        if g_add:
            g = g + _view(g_add, d.N * T * H * W * add_cstride, np.float32).view(d.N, T, H, W, add_cstride)[..., add_coff:add_coff + d.C]
        x_hi = _bf16(d.x_hi, d.N * T * H * W * d.in_cstride).view(d.N, T, H, W, d.in_cstride)[..., d.in_coff:d.in_coff + d.C].float()
        g = g * (x_hi > 0)
        if scale:
            g = g * _view(scale, d.C, np.float32)
        _store(g, d_hi, d_lo, (d.N, T, H, W, d_cstride), d_coff)

    def otal_relu_bn_bwd_split(self, g, y_hi, scale, d_hi, d_lo, npos, C, g_cstride, g_coff, y_cstride, y_coff, d_cstride, d_coff, relu, stream):
        gv = _view(g, npos * g_cstride, np.float32).view(npos, g_cstride)[:, g_coff:g_coff + C]
        out = gv.clone()
        if relu and y_hi:
            out = out * (_bf16(y_hi, npos * y_cstride).view(npos, y_cstride)[:, y_coff:y_coff + C].float() > 0)
        if scale:
            out = out * _view(scale, C, np.float32)
        _store(out, d_hi, d_lo, (npos, d_cstride), d_coff)


    # ---------------------------------------------------------------------------------------------- detection head
    def otal_ncl_to_nlc_split(self, x, hi, lo, B, C, T, Cpad, Ttot, dilate, offset, stream):
        v = _view(x, B * C * T, np.float32).view(B, C, T).permute(0, 2, 1)                        # [B,T,C]
        n = B * Ttot * Cpad
        h = v.bfloat16()
        pos = offset + torch.arange(T) * dilate
        _bf16(hi, n).view(B, Ttot, Cpad)[:, pos, :C] = h
        if lo:
            _bf16(lo, n).view(B, Ttot, Cpad)[:, pos, :C] = (v - h.float()).bfloat16()

    @staticmethod
    def _segments(nseg, so, sl, T):
        return [(int(so[i]), int(sl[i])) for i in range(nseg)] if nseg else [(0, T)]

    def otal_groupnorm_relu_fwd(self, x, gamma, beta, y, mean, rstd, B, C, T, G, eps, relu, nseg, so, sl, stream):
        xv = _view(x, B * C * T, np.float32).view(B, C, T)
        yv = _view(y, B * C * T, np.float32).view(B, C, T)
        ga, be = _view(gamma, C, np.float32), _view(beta, C, np.float32)
        yv.zero_()
        for off, ln in self._segments(nseg, so, sl, T):
            out = F.group_norm(xv[:, :, off:off + ln], G, ga, be, eps)
            yv[:, :, off:off + ln] = out.relu() if relu else out

    def otal_groupnorm_relu_bwd(self, gy, x, gamma, beta, mean, rstd, gx, dgb, B, C, T, G, relu, nseg, so, sl, stream):
        gyv = _view(gy, B * C * T, np.float32).view(B, C, T)
        xv = _view(x, B * C * T, np.float32).view(B, C, T)
        gxv = _view(gx, B * C * T, np.float32).view(B, C, T)
        out = _view(dgb, B * 2 * C, np.float32).view(B, 2, C)
        gxv.zero_(), out.zero_()
        for b in range(B):                                      # per-sample partial sums of the parameter gradients
            for off, ln in self._segments(nseg, so, sl, T):
                with torch.enable_grad():
                    xs = xv[b:b + 1, :, off:off + ln].clone().requires_grad_(True)
                    ga = _view(gamma, C, np.float32).clone().requires_grad_(True)
                    be = _view(beta, C, np.float32).clone().requires_grad_(True)
                    o = F.group_norm(xs, G, ga, be, 1e-5)
                    o = o.relu() if relu else o
                    g1, g2, g3 = torch.autograd.grad(o, (xs, ga, be), gyv[b:b + 1, :, off:off + ln])
                gxv[b:b + 1, :, off:off + ln] = g1
                out[b, 0] += g2
                out[b, 1] += g3

    # ---------------------------------------------------------------------------------------------- explicit head schedule
    def otal_groupnorm_relu_fwd_ex(self, desc, stream):
        d = desc._obj
        B, C, T, G = d.B, d.C, d.T, d.groups
        xv = _view(d.x, B * C * T, np.float32).view(B, C, T)
        ga, be = _view(d.gamma, C, np.float32), _view(d.beta, C, np.float32)
        y = torch.zeros(B, C, T)
        for off, ln in self._segments(d.nseg, d.seg_off, d.seg_len, T):
            out = F.group_norm(xv[:, :, off:off + ln], G, ga, be, d.eps)
            y[:, :, off:off + ln] = out.relu() if d.relu else out
        if d.y:
            _view(d.y, B * C * T, np.float32).view(B, C, T).copy_(y)
        if d.p_hi:
            _store(y.permute(0, 2, 1).contiguous(), d.p_hi, d.p_lo, (B, T, d.p_cstride), d.p_coff)
        if d.yt:
            _view(d.yt, B * d.yt_T * C, np.float32).view(B, d.yt_T, C).copy_(y[:, :, d.yt_off:d.yt_off + d.yt_T].permute(0, 2, 1))

    def otal_groupnorm_relu_bwd_ex(self, desc, stream):
        d = desc._obj
        B, C, T, G = d.B, d.C, d.T, d.groups
        xv = _view(d.x, B * C * T, np.float32).view(B, C, T)
        gy = torch.zeros(B, C, T)
        if d.gy:
            base = d.gy.value if isinstance(d.gy, ctypes.c_void_p) else int(d.gy)
            for b in range(B):
                gy[b] = _view(base + 4 * b * d.gy_bstride, C * T, np.float32).view(C, T)
        for ptr, lo in ((d.gy2a, 0), (d.gy2b, C // 2)):
            if ptr:
                gy[:, lo:lo + C // 2, d.gy2_off:d.gy2_off + d.gy2_T] += _view(ptr, B * d.gy2_T * (C // 2), np.float32).view(B, d.gy2_T, C // 2).permute(0, 2, 1)
        gx = torch.zeros(B, C, T)
        dga, dbe = _view(d.dgamma, C, np.float32), _view(d.dbeta, C, np.float32)
        for off, ln in self._segments(d.nseg, d.seg_off, d.seg_len, T):
            with torch.enable_grad():
                xs = xv[:, :, off:off + ln].clone().requires_grad_(True)
                ga = _view(d.gamma, C, np.float32).clone().requires_grad_(True)
                be = _view(d.beta, C, np.float32).clone().requires_grad_(True)
                o = F.group_norm(xs, G, ga, be, 1e-5)
                o = o.relu() if d.relu else o
                g1, g2, g3 = torch.autograd.grad(o, (xs, ga, be), gy[:, :, off:off + ln])
            gx[:, :, off:off + ln] = g1
            dga += g2
            dbe += g3
        if d.dbias:
            _view(d.dbias, C, np.float32).add_(gx.sum(dim=(0, 2)))
        if d.gx:
            _view(d.gx, B * C * T, np.float32).view(B, C, T).copy_(gx)
        if d.d_hi:
            _store(gx.permute(0, 2, 1).contiguous(), d.d_hi, d.d_lo, (B, T, C), 0)

    def otal_rows_combine(self, desc, stream):
        d = desc._obj
        B, C, Td, NP = d.B, d.C, d.Td, d.npairs
        tab = _view(d.table, Td * NP * 2, np.int32).view(Td, NP, 2)
        srcs = [_view(d.src[i], B * C * d.src_T[i], np.float32).view(B, C, d.src_T[i]) for i in range(d.nsrc)]
        out = torch.zeros(B, C, Td)
        for j in range(Td):
            for k in range(NP):
                si, col = int(tab[j, k, 0]), int(tab[j, k, 1])
                if si >= 0:
                    out[:, :, j] += srcs[si][:, :, col]
        if d.dst:
            _view(d.dst, B * C * Td, np.float32).view(B, C, Td).copy_(out)
        if d.p_hi:
            _store(out.permute(0, 2, 1).contiguous(), d.p_hi, d.p_lo, (B, Td, C), 0)

    def _headout(self, d):
        B, S, P = d.B, d.S, d.P
        sep = _view(d.sep_idx, P, np.int32).long()
        lvl = _view(d.level_id, P, np.int32).long() if d.level_id else None
        mult = _view(d.mult, P, np.float32) if d.mult else None
        return B, S, P, sep, lvl, mult

    def otal_head_gather_fwd(self, desc, stream):
        d = desc._obj
        B, S, P, sep, lvl, mult = self._headout(d)
        for k in range(d.n):
            cp, co = d.cpad[k], d.cout[k]
            raw = _view(d.raw[k], B * cp * S, np.float32).view(B, cp, S)
            v = raw[:, :co, sep].permute(0, 2, 1)                                             # [B,P,co]
            if d.bias[k]:
                v = v + _view(d.bias[k], co, np.float32)
            if d.mode[k] == 1:
                sc = torch.stack([_view(d.scale[int(l)], 1, np.float32)[0] for l in lvl]).view(1, P, 1)
                v = torch.exp(v * sc)
                if mult is not None:
                    v = v * mult.view(1, P, 1)
            _view(d.out[k], B * P * co, np.float32).view(B, P, co).copy_(v)

    def otal_head_gather_bwd(self, desc, prior_of_col, stream):
        d = desc._obj
        B, S, P, sep, lvl, mult = self._headout(d)
        for k in range(d.n):
            cp, co = d.cpad[k], d.cout[k]
            g = _view(d.gout[k], B * P * co, np.float32).view(B, P, co).clone() if d.gout[k] else torch.zeros(B, P, co)
            if d.mode[k] == 1:
                raw = _view(d.raw[k], B * cp * S, np.float32).view(B, cp, S)[:, :co, sep].permute(0, 2, 1)
                if d.bias[k]:
                    raw = raw + _view(d.bias[k], co, np.float32)
                out = _view(d.out[k], B * P * co, np.float32).view(B, P, co)
                for l in range(int(lvl.max()) + 1):
                    m = lvl == l
                    _view(d.dscale[l], 1, np.float32).add_((g[:, m] * out[:, m] * raw[:, m]).sum())
                sc = torch.stack([_view(d.scale[int(l)], 1, np.float32)[0] for l in lvl]).view(1, P, 1)
                g = g * out * sc
            if d.dbias[k]:
                _view(d.dbias[k], co, np.float32).add_(g.sum(dim=(0, 1)))
            full = torch.zeros(B, S, cp)
            full[:, sep, :co] = g
            _store(full, d.d_hi[k], d.d_lo[k], (B, S, cp), 0)

    def otal_ncl_to_nlc_split_ex(self, x, x_bstride, hi, lo, B, C, T, cstride, coff, stream):
        base = x.value if isinstance(x, ctypes.c_void_p) else int(x)
        v = torch.stack([_view(base + 4 * b * x_bstride, C * T, np.float32).view(C, T) for b in range(B)]).permute(0, 2, 1).contiguous()
        _store(v, hi, lo, (B, T, cstride), coff)

    def otal_boundary_bce_fwd_ex(self, x, x_rstride, target, tstride, row_loss, coef, B, T, C, stream):
        xv = _view(x, (B * T - 1) * x_rstride + C, np.float32)
        rows = torch.stack([xv[r * x_rstride:r * x_rstride + C] for r in range(B * T)]).view(B, T, C)
        tg = torch.stack([_view(int(target) + 4 * b * int(tstride), T, np.float32) for b in range(B)])
        s = torch.tanh(rows).mean(-1)
        loss = -(tg * torch.log(s).clamp(min=-100) + (1 - tg) * torch.log(1 - s).clamp(min=-100))
        _view(row_loss, B * T, np.float32).copy_(loss.reshape(-1))
        _view(coef, B * T, np.float32).copy_(((s - tg) / (s * (1 - s)).clamp(min=1e-12) / (B * T * C)).reshape(-1))

    def otal_segment_mean(self, x, offsets, nseg, out, stream):
        offs = [int(offsets[i]) for i in range(nseg + 1)]
        xv = _view(x, offs[-1], np.float32)
        ov = _view(out, nseg, np.float32)
        for i in range(nseg):
            ov[i] = xv[offs[i]:offs[i + 1]].mean() if offs[i + 1] > offs[i] else 0.0

    def otal_boundary_bce_bwd_ex(self, x, x_rstride, coef, g, gx, B, T, C, stream):
        xv = _view(x, (B * T - 1) * x_rstride + C, np.float32)
        rows = torch.stack([xv[r * x_rstride:r * x_rstride + C] for r in range(B * T)])
        th = torch.tanh(rows)
        _view(gx, B * T * C, np.float32).view(B * T, C).copy_(_view(g, 1, np.float32) * _view(coef, B * T, np.float32).view(-1, 1) * (1 - th * th))

    def otal_make_segments(self, loc, prior, level_len, level_off, seg_level, seg_concat, frame_seg, B, P, frame_num, stream):
        lc = _view(loc, B * P * 2, np.float32).view(B, P, 2)
        pri = _view(prior, P, np.float32).view(1, P, 1)
        t = _view(level_len, P, np.int32).view(1, P, 1).float()
        off = _view(level_off, P, np.int32).view(1, P, 1).float()
        seg = lc / frame_num * t                                                                  # BDNet.py:355-384, all levels at once
        centre = torch.round(pri * t - 0.5)
        plen = seg[:, :, :1] + seg[:, :, 1:]
        inl, outl = torch.clamp(plen / 4.0, min=1.0), torch.clamp(plen / 10.0, min=1.0)
        ls, rs = centre - seg[:, :, :1], centre + seg[:, :, 1:]
        segments = torch.cat([torch.round(ls - outl), torch.round(ls + inl), torch.round(rs - inl), torch.round(rs + outl)], -1)
        if seg_level:
            _view(seg_level, B * P * 4, np.float32).view(B, P, 4).copy_(segments)
        clamped = torch.minimum(segments.trunc().clamp(min=0), t - 1) + off                       # kernel.cu:33-38 + level offset
        _view(seg_concat, B * P * 4, np.float32).view(B, P, 4).copy_(clamped)
        dl, dr = pri * frame_num - lc[:, :, :1], pri * frame_num + lc[:, :, 1:]
        plen = dr - dl + 1.0
        inl, outl = torch.clamp(plen / 4.0, min=1.0), torch.clamp(plen / 10.0, min=1.0)
        fs = torch.cat([torch.round(dl - outl), torch.round(dl + inl), torch.round(dr - inl), torch.round(dr + outl)], -1)
        _view(frame_seg, B * P * 4, np.float32).view(B, P, 4).copy_(fs)

    def otal_make_segments_ex(self, loc, prior, level_len, level_off, out_row, S, seg_concat, frame_seg, B, P, frame_num, stream):
        sc, fs = torch.zeros(B, P, 4), torch.zeros(B, P, 4)
        self.otal_make_segments(loc, prior, level_len, level_off, None, sc.data_ptr(), fs.data_ptr(), B, P, frame_num, stream)
        rows = _view(out_row, P, np.int32).long()
        _view(seg_concat, B * S * 4, np.float32).view(B, S, 4)[:, rows] = sc
        _view(frame_seg, B * S * 4, np.float32).view(B, S, 4)[:, rows] = fs

    def otal_dirichlet_uncertainty(self, logit, unct, M, K, stream):
        x = _view(logit, M * K, np.float32).view(M, K)
        _view(unct, M, np.float32).copy_(K / (torch.exp(torch.clamp(x, -10, 10)) + 1).sum(-1))

    def otal_bmp_forward_f32(self, inp, seg, out, B, C, T, K, stream):
        import opental_oracle as O
        x = _view(inp, B * C * T, np.float32).view(B, C, T)
        sg = _view(seg, B * K * 4, np.float32).view(B, K, 4)
        _view(out, B * C * K, np.float32).view(B, C, K).copy_(O.boundary_max_pooling(x, sg, False))

    def otal_bmp_backward_f32(self, gout, inp, seg, gin, B, C, T, K, compat, stream):
        import opental_oracle as O
        sg = _view(seg, B * K * 4, np.float32).view(B, K, 4)
        with torch.enable_grad():
            x = _view(inp, B * C * T, np.float32).view(B, C, T).clone().requires_grad_(True)
            (g,) = torch.autograd.grad(O.boundary_max_pooling(x, sg, bool(compat)), x, _view(gout, B * C * K, np.float32).view(B, C, K))
        _view(gin, B * C * T, np.float32).view(B, C, T).copy_(g)

    def otal_boundary_bce_fwd(self, x, target, tstride, row_loss, coef, B, T, C, stream):
        xv = _view(x, B * T * C, np.float32).view(B, T, C)
        tg = torch.stack([_view(int(target) + 4 * b * int(tstride), T, np.float32) for b in range(B)])      # rows of the score maps
        s = torch.tanh(xv).mean(-1)
        loss = -(tg * torch.log(s).clamp(min=-100) + (1 - tg) * torch.log(1 - s).clamp(min=-100))
        _view(row_loss, B * T, np.float32).copy_(loss.reshape(-1))
        _view(coef, B * T, np.float32).copy_(((s - tg) / (s * (1 - s)).clamp(min=1e-12) / (B * T * C)).reshape(-1))

    def otal_boundary_bce_bwd(self, x, coef, g, gx, B, T, C, stream):
        xv = _view(x, B * T * C, np.float32).view(B * T, C)
        th = torch.tanh(xv)
        _view(gx, B * T * C, np.float32).view(B * T, C).copy_(_view(g, 1, np.float32) * _view(coef, B * T, np.float32).view(-1, 1) * (1 - th * th))

    def otal_msl_forward(self, desc, stream):
        """The 7 losses by the ORACLE's restatement of the reference loss (independent of the product's torch formulation) — the
        THUMOS14 EDL, ActivityNet EDL or closed-set focal flavour; the unit gradients of each loss w.r.t. each head output are kept
        on the side, keyed by the workspace pointer."""
        import opental_oracle as O
        d = desc._obj
        B, P, K, G = d.B, d.P, d.K, d.G
        names = ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act")
        shapes = ((B, P, 2), (B, P, K), (B, P, 2), (B, P, K), (B, P, 1), (B, P, 1), (B, P, 1))
        ptrs = (d.loc, d.conf, d.prop_loc, d.prop_conf, d.center, d.act, d.prop_act)
        focal, anet = d.flavour == 2, d.flavour == 1
        assert d.reweight in (0, 1) and not d.cls_all, "abi_emu: the ablation branches of the fused loss are GPU-tested only"
        assert focal or (d.act and d.prop_act), "abi_emu: the EDL flavours of the fused loss are emulated for the os_head configuration"
        with torch.enable_grad():
            out = {n: (_view(p_, int(np.prod(sh)), np.float32).view(*sh).clone().requires_grad_(True) if p_ else None)
                   for n, sh, p_ in zip(names, shapes, ptrs)}
            ncol = 2 if anet else 1
            pri = torch.stack([_view(d.priors, (P - 1) * d.prior_stride + ncol, np.float32)[c::d.prior_stride][:P] for c in range(ncol)], 1)
            tg = _view(d.targets, B * G * 3, np.float32).view(B, G, 3)
            va = _view(d.valid, B * G, np.uint8).view(B, G).bool()
            targets = [tg[b][va[b]] for b in range(B)]
            cfg = O.OracleConfig(num_classes=K, clip_length=int(d.clip_length), piou=float(d.overlap_thresh), with_ibm=bool(d.use_ibm),
                                 ibm_start=10, momentum=float(d.momentum), num_bins=max(int(d.num_bins), 1), iou_aware=bool(d.iou_aware),
                                 act_weight=float(d.act_weight), act_margin=float(d.act_margin))
            state = O.LossState(epoch=11 if d.use_ibm else 1)
            wa = _view(d.weight_accum, d.num_bins, np.float32) if d.weight_accum and d.num_bins and not anet else None
            if wa is not None:
                state.weight_accum = wa.clone()
            if focal:
                losses = O.multisegment_loss_closed(dict(out, priors=pri), targets, cfg)
            elif anet:
                cfg.ibm_coeff = float(d.ibm_coeff)
                assert [float(v) for v in d.level_bounds[:12]] == [float(v) for pair in O.ANET_BOUNDS for v in pair]
                losses = O.multisegment_loss_anet(dict(out, priors=pri), targets, state, cfg)
            else:
                losses = O.multisegment_loss(dict(out, priors=pri), targets, state, cfg)
            live = [out[n] for n in names if out[n] is not None]
            unit = []
            for l in losses:
                gs = iter(torch.autograd.grad(l, live, retain_graph=True, allow_unused=True) if torch.is_tensor(l) and l.requires_grad
                          else [None] * len(live))
                unit.append([next(gs) if out[n] is not None else None for n in names])
            unit += [[None] * 7] * (7 - len(unit))
        if wa is not None:
            wa.copy_(state.weight_accum)
        lv = _view(d.losses, 16, np.float32)
        lv.zero_()
        for i, l in enumerate(losses):
            lv[i] = float(l)
        if not focal and not anet:
            _, conf_t, _, prop_conf_t, iou_pred = O.match_priors(out["loc"].detach(), pri, targets, cfg)
            pos, ppos = (conf_t > 0).view(-1), (prop_conf_t > 0).view(-1)
            lv[7], lv[8] = float(pos.sum()), float(ppos.sum())
            lv[9] = float(O.actionness_loss(out["act"].detach().view(-1, 1), pos.float(), cfg)[1])
            lv[10] = float(O.actionness_loss(out["prop_act"].detach().view(-1, 1), ppos.float(), cfg)[1])
            lv[11] = float(O.iou_calibration(out["prop_conf"].detach().view(-1, K), iou_pred.reshape(-1), cfg)) if d.iou_aware else 0.0
        if not hasattr(self, "_msl"):
            self._msl = {}
        key = d.workspace.value if isinstance(d.workspace, ctypes.c_void_p) else int(d.workspace)
        self._msl[key] = [[g if g is not None else torch.zeros(sh) for g, sh in zip(gs, shapes)] for gs in unit]

    def otal_msl_backward(self, B, P, K, ws, grad_losses, g_loc, g_conf, g_ploc, g_pconf, g_center, g_act, g_pact, stream):
        unit = self._msl[int(ws)]
        gl = _view(grad_losses, 7, np.float32)
        outs = (g_loc, g_conf, g_ploc, g_pconf, g_center, g_act, g_pact)
        for j, ptr in enumerate(outs):
            if not ptr:
                continue
            tot = sum(float(gl[i]) * unit[i][j] for i in range(7))
            _view(ptr, tot.numel(), np.float32).copy_(tot.reshape(-1))

    # ---------------------------------------------------------------------------------------------- inference post-processing
    def otal_decode_scores(self, loc, ploc, conf, pconf, center, act, pact, prior, offset, seg, scores, unct, actn, B, P, K, clip_length,
                           sample_fps, stream):
        import opental_oracle as O
        assert act and pact, "abi_emu: decode is emulated for the open-set head"
        v = lambda p_, *sh: _view(p_, int(np.prod(sh)), np.float32).view(*sh)           # noqa: E731
        out = dict(loc=v(loc, B, P, 2), prop_loc=v(ploc, B, P, 2), conf=v(conf, B, P, K), prop_conf=v(pconf, B, P, K),
                   center=v(center, B, P, 1), act=v(act, B, P, 1), prop_act=v(pact, B, P, 1), priors=v(prior, P, 1))
        offs = v(offset, B) if offset else torch.zeros(B)
        cfg = O.OracleConfig(num_classes=K, clip_length=int(clip_length))
        for b in range(B):
            s_, sc_, u_, a_ = O.decode_predictions(out, b, float(offs[b]), sample_fps, cfg)
            v(seg, B, P, 2)[b], v(scores, B, K, P)[b], v(unct, B, P)[b], v(actn, B, P)[b] = s_, sc_, u_, a_

    def otal_softnms(self, segments, stride, scores, keep, count, C, M, sigma, top_k, thr, stream):
        import opental_oracle as O
        sc = _view(scores, C * M, np.float32).view(C, M)
        kp = _view(keep, C * M, np.uint8).view(C, M)
        ct = _view(count, C, np.int32)
        for c in range(C):
            sg = _view(int(segments) + 4 * c * int(stride), M * 2, np.float32).view(M, 2)
            cand = torch.cat([sg, sc[c].clone().view(-1, 1)], -1)
            # softnms_v2 decays a working copy in place; re-run it on a copy that exposes the decayed scores
            work = cand.clone()
            _, n, mask = self._softnms(work, sigma, top_k, thr)
            sc[c], kp[c], ct[c] = work[:, 2], mask.to(torch.uint8), n

    @staticmethod
    def _softnms(seg, sigma, top_k, thr):
        """O.softnms_v2 on `seg` itself (no clone), so that the caller sees the decayed scores."""
        ts, te, sc = seg[:, 0], seg[:, 1], seg[:, 2]
        done = torch.zeros_like(sc, dtype=torch.bool)
        undone = sc >= thr
        while int(undone.sum()) > 1 and int(done.sum()) < top_k:
            cand = undone.nonzero().view(-1)
            idx = int(cand[sc[undone].argmax()])
            undone[idx] = False
            done[idx] = True
            tt1, tt2 = ts[undone].clamp(min=float(ts[idx])), te[undone].clamp(max=float(te[idx]))
            inter = (tt2 - tt1).clamp(min=0)
            iou = inter / (torch.clamp(te[idx] - ts[idx], min=1e-5) + (te[undone] - ts[undone]) - inter)
            sc[undone] = sc[undone] * torch.exp(-iou ** 2 / sigma)
            undone[sc < thr] = False
        return seg[done], int(done.sum()), done

    def otal_adam_step_dev(self, p, g, m, v, n, lr, b1, b2, eps, wd, grad_scale, step_dev, stream):
        self.otal_adam_step(p, g, m, v, n, lr, b1, b2, eps, wd, grad_scale, int(_view(step_dev, 1, np.int32)[0]), stream)

    def otal_adam_step(self, p, g, m, v, n, lr, b1, b2, eps, wd, grad_scale, step, stream):
        pv, gv, mv, vv = (_view(t, n, np.float32) for t in (p, g, m, v))
        gr = gv * grad_scale + wd * pv
        mv.mul_(b1).add_(gr, alpha=1 - b1)
        vv.mul_(b2).addcmul_(gr, gr, value=1 - b2)
        pv.sub_(lr / (1 - b1 ** step) * mv / ((vv / (1 - b2 ** step)).sqrt() + eps))


def install(monkeypatch) -> Emulator:
    """Route every `_lib.call` to the emulator and let CPU tensors through the wrappers' CUDA guards (tests only)."""
    from opental_b200 import _lib, ops
    emu = Emulator()
    monkeypatch.setattr(_lib, "call", emu)
    monkeypatch.setattr(ops, "_require_cuda", lambda *t: None)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "OVERLAP_WGRAD", False)
    monkeypatch.setattr(ops, "join", lambda: None)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)
    import contextlib
    monkeypatch.setattr(torch.cuda, "device", lambda *_a, **_k: contextlib.nullcontext())
    return emu
