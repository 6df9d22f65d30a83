"""Explicit forward / backward schedule of CoarsePyramid (AFSD/thumos14/BDNet.py:295-432; AFSD/anet/BDNet.py:281-391).

`bdnet.CoarsePyramid.forward` used to glue the native kernels with torch autograd: every conv was its own autograd node, and
F.interpolate / + / torch.cat / index_select / permute, the gradient sums of every fan-out, the bias gradients (`sum`) and the
AccumulateGrad of ~150 small parameters ran as ~300 ATen launches per step around ~230 native ones (round-1 launch list).  Here
the whole head is ONE autograd node with a hand-written schedule, like the backbone:

  * every activation between two convolutions lives in the LEVEL-SEPARATED ("sep") layout [B,C,S]: the 6 pyramid levels side by
    side along T with one zero column between them (S = 136 for 256-frame clips).  A k=3 "same" conv sees its per-level zero
    padding, a 1x1 conv does not care, GroupNorm normalises the level ranges independently and writes zeros in between.  The
    proposal branches run in the same layout (the reference's per-level loop, BDNet.py:386-397, batched over the levels), so no
    index_select is needed between the towers and the branches.
  * GroupNorm+ReLU writes its result directly as the channels-last bf16 hi/lo planes the next tensor-core conv reads (or into
    its slice of the [roi | boundary | centre] concat buffer, BDNet.py:111), and its backward writes the conv's output-gradient
    planes and accumulates d gamma, d beta and the conv's bias gradient in place (ops.groupnorm_relu_*_ex).
  * a fan-out's gradient sum is the `accumulate` mode of the data-gradient convs into one fp32 buffer.
  * upsample + top-down add + the sep layout itself and their transposes are ops.rows_combine with small index tables;
    ops.head_gather_* turns the head convs' raw outputs into the reference's [B,P,C] tensors (bias, ScaleExp, FPN stride).

Weight gradients go to the side stream (ops.fork / join), joined once at the end of the backward.
"""
from __future__ import annotations

import torch

from . import ops
from .headconv import _same_pad_front
from .ops import Planes
from .prop_pooling import BoundaryMaxPoolingFunction


# ------------------------------------------------------------------------------------------------------------------
# tables
# ------------------------------------------------------------------------------------------------------------------
def _tables(cp, device) -> dict:
    """Index tables of the schedule on `device` (built once): the sep layout, the frame-level upsampling and their transposes."""
    key = ("sched", device)
    if key in cp._tables:
        return cp._tables[key]
    L = cp.layer_num
    thumos = cp.variant == "thumos"
    t = cp.level_t
    S, P = cp.sep_len, cp.num_priors
    sep_off = [o for o, _ in cp.sep_segments]
    r0 = cp.frame_num // t[0]                          # frame positions per level-0 position (4; ActivityNet 8)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=device).contiguous()   # noqa: E731
    # forward: x_sep column j <- level i column c (+ the top-down term of level 0: BDNet.py:317-319)
    np_sep = 2 if thumos else 1
    sep = [[[-1, 0]] * np_sep for _ in range(S)]
    prior_of_col = [-1] * S
    p = 0
    for i in range(L):
        for c in range(t[i]):
            e = [[i, c]]
            if thumos:
                e.append([1, c // 2] if i == 0 else [-1, 0])
            sep[sep_off[i] + c] = e
            prior_of_col[sep_off[i] + c] = p
            p += 1
    # forward: frame-level input, F.interpolate(feats[0], frame_num) (BDNet.py:324): nearest, source column j // r0
    frame = []
    for j in range(cp.frame_num):
        e = [[0, j // r0]]
        if thumos:
            e.append([1, j // (2 * r0)])
        frame.append(e)
    # backward: sources [d_x_sep, d_frame_in, d_p0]
    d_lvl = []
    for i in range(L):
        rows = []
        for c in range(t[i]):
            e = [[0, sep_off[i] + c]]
            if i == 0:
                e += [[1, r0 * c + r] for r in range(r0)]
            if i == 1 and thumos:
                e += [[2, 2 * c], [2, 2 * c + 1]]
            rows.append(e)
        d_lvl.append(i32(rows))
    tb = cp._tables_on(device)
    out = dict(sep=i32(sep), frame=i32(frame), d_lvl=d_lvl, prior_of_col=i32(prior_of_col),
               sep_idx=tb["sep_idx"].to(torch.int32).contiguous(), level_id=tb["level_id"].to(torch.int32).contiguous(),
               level_off_sep=i32([sep_off[i] for i in range(L) for _ in range(t[i])]), seg_buf={})
    cp._tables[key] = out
    return out


def _dilated(st: dict, g: torch.Tensor, T_in: int, with_lo: bool) -> Planes:
    """The zero-upsampled output gradient of a stride-2 conv as channels-last planes [B,T_in,1,1,C] (its data gradient is the stride-1
    data gradient of that): written into a cached buffer whose odd rows were zeroed once and are never touched again."""
    from . import _lib
    B, C, To = g.shape
    key = ("dil", B, C, T_in)
    if key not in st["seg_buf"]:
        hi = torch.zeros((B, T_in, 1, 1, C), dtype=torch.bfloat16, device=g.device)
        st["seg_buf"][key] = Planes(hi, torch.zeros_like(hi) if with_lo else None)
    pl = st["seg_buf"][key]
    _lib.call("otal_ncl_to_nlc_split", g.data_ptr(), pl.hi.data_ptr(), ops._ptr(pl.lo), B, C, To, C, T_in, 2, 0, ops._stream())
    return pl


def _grad(p: torch.Tensor) -> torch.Tensor:
    """The parameter's gradient buffer (the Trainer binds it to a flat buffer and zeroes it every step; allocated here otherwise)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


# ------------------------------------------------------------------------------------------------------------------
# layer helpers
# ------------------------------------------------------------------------------------------------------------------
class _Ctx:
    """Per-call state: the conv store, precision, and what the backward needs."""

    def __init__(self, cp, dev, need_grad):
        self.cp, self.dev, self.need_grad = cp, dev, need_grad
        self.store = cp.conv_store
        self.with_lo = self.store.with_lo
        self.sv: dict = {}
        self.keep: list = []        # tensors a side stream reads: referenced until the stream has been joined
        self.forked = False


def _conv(c: _Ctx, xp: Planes, unit, stride: int = 1, conv3d: bool = False) -> torch.Tensor:
    """Unit1D / head-side Unit3D on channels-last planes -> fp32 [B,cpad,To] (bias added only when the output is not padded)."""
    store, rec = unit._native
    mod = unit.conv3d if conv3d else unit.conv1d
    B, T = xp.hi.shape[0], xp.hi.shape[1]
    k = rec.taps
    pf = _same_pad_front(T, k, stride)
    To = -(-T // stride)
    y = torch.empty((B, rec.cpad, To), dtype=torch.float32, device=c.dev)
    fused_bias = mod.bias is not None and rec.cpad == rec.cout
    ops.conv_igemm(xp, store.w(rec), kernel=(k, 1, 1), pad_front=(pf, 0, 0), stride=(stride, 1, 1),
                   shift=mod.bias.detach() if fused_bias else None, out_f32=y, want_planes=False, f32_ncdhw=True)
    return y


def _conv_bwd(c: _Ctx, xp: Planes, dp: Planes, unit, *, stride: int = 1, gx: torch.Tensor | None = None, accumulate: bool = False,
              want_dgrad: bool = True, want_wgrad: bool = True, dp_dgrad: Planes | None = None,
              ndhwc_out: torch.Tensor | None = None) -> torch.Tensor | None:
    """Weight gradient (side stream) + data gradient of one head conv.  dp: the output gradient as channels-last planes.
    gx: fp32 [B,Cin,T] destination (allocated if None); accumulate = add to it.  dp_dgrad: the zero-upsampled gradient planes of a
    strided conv.  ndhwc_out: destination of a head-side Unit3D's data gradient ([B,T,1,1,Cin])."""
    store, rec = unit._native
    T = xp.hi.shape[1]
    k = rec.taps
    pf = _same_pad_front(T, k, stride)
    if rec.weight.requires_grad and want_wgrad:
        dw = store.block(store.flat_g, rec)
        if ops.OVERLAP_WGRAD:
            with torch.cuda.stream(ops.fork()):
                ops.conv_wgrad(xp, dp, dw, kernel=(k, 1, 1), pad_front=(pf, 0, 0), stride=(stride, 1, 1))
            c.forked = True
            c.keep.append((xp, dp))      # the caching allocator would hand a released block to the next main-stream allocation
        else:
            ops.conv_wgrad(xp, dp, dw, kernel=(k, 1, 1), pad_front=(pf, 0, 0), stride=(stride, 1, 1))
    if not want_dgrad:
        return None
    d = dp_dgrad if dp_dgrad is not None else dp
    if ndhwc_out is not None:
        ops.conv_igemm(d, store.w(rec), kernel=(1, 1, 1), pad_front=(0, 0, 0), out_f32=ndhwc_out, want_planes=False, dgrad=True)
        return ndhwc_out
    if gx is None:
        gx = torch.empty((xp.hi.shape[0], rec.cin, T), dtype=torch.float32, device=c.dev)
        accumulate = False
    ops.conv_igemm(d, store.w(rec), kernel=(k, 1, 1), pad_front=(k - 1 - pf, 0, 0), out_f32=gx, want_planes=False, dgrad=True,
                   f32_ncdhw=True, accumulate=accumulate)
    return gx


def _gn(c: _Ctx, x: torch.Tensor, gn, name: str, *, segments=None, want_y=False, want_planes=True, planes=None, coff=0, yt_range=None):
    """GroupNorm+ReLU of a conv output; saves (x, stats) under `name` for the backward.  Returns (y, planes, yt)."""
    y, pl, yt, stats = ops.groupnorm_relu_fwd_ex(x, gn.weight.detach(), gn.bias.detach(), groups=gn.num_groups, eps=gn.eps, relu=True,
                                                 segments=segments, want_y=want_y, planes=planes, planes_coff=coff,
                                                 want_planes=want_planes and planes is None, with_lo=c.with_lo, yt_range=yt_range)
    if c.need_grad:
        c.sv["gn:" + name] = (x, stats, segments)
    return y, pl, yt


def _gn_bwd(c: _Ctx, name: str, gn, gy, conv_unit, *, gy_coff=0, gy2=None, gy2_off=0, want_gx=False, conv3d=False):
    """Backward of `_gn`: returns (conv output-gradient planes, fp32 copy | None); accumulates d gamma / d beta and the bias
    gradient of the conv in front of it."""
    x, stats, segments = c.sv.pop("gn:" + name)
    mod = conv_unit.conv3d if conv3d else conv_unit.conv1d
    dbias = _grad(mod.bias) if (mod.bias is not None and mod.bias.requires_grad) else None
    return ops.groupnorm_relu_bwd_ex(gy, x, gn.weight.detach(), gn.bias.detach(), stats, dgamma=_grad(gn.weight), dbeta=_grad(gn.bias),
                                     dbias=dbias, groups=gn.num_groups, relu=True, segments=segments, gy_coff=gy_coff,
                                     with_lo=c.with_lo, want_gx=want_gx, gy2=gy2, gy2_off=gy2_off)


# ------------------------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------------------------
def forward(cp, x1, x2, forced_segments, need_grad: bool):
    """x1 / x2: the backbone's Mixed_4f / Mixed_5c feature maps as NCDHW views of channels-last fp32 storage (x1 None for the
    ActivityNet flavour).  Returns (outputs dict, saved state | None)."""
    dev = x2.device
    c = _Ctx(cp, dev, need_grad)
    store = c.store
    store.prepare(dev)
    if need_grad:
        store.bind_grads()
    tb = cp._tables_on(dev)
    st = _tables(cp, dev)
    thumos = cp.variant == "thumos"
    B = x2.size(0)
    L, S, P, t = cp.layer_num, cp.sep_len, cp.num_priors, cp.level_t
    segs = cp.sep_segments
    sv = c.sv

    def feat_planes(x):
        xc = x.permute(0, 2, 3, 4, 1).contiguous()          # the backbone hands out channels-last storage: no copy on that path
        Bx, T = xc.shape[0], xc.shape[1]
        return ops.split_bf16(xc.reshape(Bx, T, 1, 1, -1), c.with_lo)

    # ---- pyramid (BDNet.py:311-322; anet/BDNet.py:281-289)
    pf32, ppl, pin = [], [], []           # per level: fp32 [B,512,t_i], planes (or None), the input planes of its conv
    for i, blk in enumerate(cp.pyramids):
        unit, gn = blk[0], blk[1]
        if i == 0 or (i == 1 and thumos):
            xin = feat_planes(x1 if (i == 0 and thumos) else x2)
            raw = _conv(c, xin, unit, conv3d=True)
        else:
            xin = ppl[i - 1]
            raw = _conv(c, xin, unit, stride=2)
        feeds_conv = i + 1 < L and not (i == 0 and thumos)                  # THUMOS14: level 1 comes from Mixed_5c, not from level 0
        y, pl, _ = _gn(c, raw, gn, f"pyr{i}", want_y=True, want_planes=feeds_conv)
        pf32.append(y); ppl.append(pl); pin.append(xin)
    _, x_sep = ops.rows_combine(pf32, st["sep"], want_planes=True, with_lo=c.with_lo)
    _, frame_in = ops.rows_combine(pf32[:2] if thumos else pf32[:1], st["frame"], want_planes=True, with_lo=c.with_lo)

    # Two independent halves run side by side (their kernels are far smaller than the GPU): the frame-level feature on the branch
    # stream next to the towers / coarse heads / windows, then the conf proposal branch next to the loc one.
    par = bool(getattr(cp, "two_streams", True)) and ops.OVERLAP_WGRAD

    # ---- frame-level feature (BDNet.py:324-331)
    dc = cp.deconv

    def frame_path():
        _, d1_, _ = _gn(c, _conv(c, frame_in, dc[0]), dc[1], "dc1")
        _, d2_, _ = _gn(c, _conv(c, d1_, dc[3]), dc[4], "dc2")
        fr, _, fr_t = _gn(c, _conv(c, d2_, dc[6]), dc[7], "dc3", want_y=True, want_planes=False, yt_range=(0, cp.frame_num))
        return d1_, d2_, fr, fr_t

    if par:
        with ops.parallel_branch():
            d1, d2, frame, frame_t = frame_path()
    else:
        d1, d2, frame, frame_t = frame_path()
    half = frame.shape[1] // 2
    start, end = frame_t[:, :, :half], frame_t[:, :, half:]

    # ---- towers and coarse heads (BDNet.py:333-353), all levels at once
    def tower(tw, tag):
        _, a, _ = _gn(c, _conv(c, x_sep, tw[0][0]), tw[0][1], tag + "1", segments=segs)
        _, b, _ = _gn(c, _conv(c, a, tw[1][0]), tw[1][1], tag + "2", segments=segs)
        return a, b

    lt1, loc_feat = tower(cp.loc_tower, "lt")
    ct1, conf_feat = tower(cp.conf_tower, "ct")
    heads = [(cp.loc_head, loc_feat, 2, 1), (cp.conf_head, conf_feat, cp.num_classes, 0)]
    if cp.os_head:
        heads.append((cp.actionness_head, conf_feat, 1, 0))
    raws = [_conv(c, f, u) for u, f, _, _ in heads]
    scales = [h.scale.detach() for h in cp.loc_heads]
    mult = tb["stride"] if cp.variant == "anet" else None
    outs = ops.head_gather_fwd(raws, [h[2] for h in heads], [h[3] for h in heads], [h[0].conv1d.bias.detach() for h in heads],
                               st["sep_idx"], st["level_id"], mult, scales)
    loc, conf = outs[0], outs[1]
    act = outs[2] if cp.os_head else None

    # ---- proposal windows (BDNet.py:355-384) in the sep layout; separator rows keep the window (0,0,0,0)
    if forced_segments is not None:
        seg_sep = torch.zeros(B, S, 4, device=dev)
        fseg_sep = torch.zeros(B, S, 4, device=dev)
        for (seg, fseg), (off, tl) in zip(forced_segments, segs):
            seg_sep[:, off:off + tl] = torch.trunc(seg).clamp(0, tl - 1) + off
            fseg_sep[:, off:off + tl] = fseg
    else:
        if B not in st["seg_buf"]:
            st["seg_buf"][B] = (torch.zeros(B, S, 4, device=dev), torch.zeros(B, S, 4, device=dev))
        seg_sep, fseg_sep = st["seg_buf"][B]
        from . import _lib
        _lib.call("otal_make_segments_ex", loc.data_ptr(), tb["centre"].data_ptr(), tb["level_len"].data_ptr(),
                  st["level_off_sep"].data_ptr(), st["sep_idx"].data_ptr(), S, seg_sep.data_ptr(), fseg_sep.data_ptr(), B, P,
                  float(cp.frame_num), ops._stream())
    if par:
        ops.join_branch()                                                            # the frame-level feature is complete
    pooled = ops.bmp_forward(frame, fseg_sep)                                        # [B,512,S], shared by both branches (F5)
    hi = torch.empty((B, S, 1, 1, pooled.shape[1]), dtype=torch.bfloat16, device=dev)
    pooled_p = Planes(hi, torch.empty_like(hi) if c.with_lo else None)
    ops.ncl_to_nlc_into(pooled, pooled_p, 0)

    # ---- the two proposal branches (BDNet.py:64-113, :386-397)
    def branch(br, feat, tag):
        pc = br.proposal_conv[0].conv1d.in_channels
        q = pc // 4
        hi = torch.empty((B, S, 1, 1, pc), dtype=torch.bfloat16, device=dev)
        cbuf = Planes(hi, torch.empty_like(hi) if c.with_lo else None)
        _gn(c, _conv(c, feat, br.cur_point_conv[0]), br.cur_point_conv[1], tag + "cp", segments=segs, planes=cbuf, coff=3 * q)
        lr, _, lr_t = _gn(c, _conv(c, feat, br.lr_conv[0]), br.lr_conv[1], tag + "lr", segments=segs, want_y=True, want_planes=False,
                          yt_range=(segs[0][0], t[0]))
        prop = ops.bmp_forward(lr, seg_sep)                                          # [B,2q,S]
        ops.ncl_to_nlc_into(prop, cbuf, q)
        _gn(c, _conv(c, pooled_p, br.roi_conv[0]), br.roi_conv[1], tag + "roi", segments=segs, planes=cbuf, coff=0)
        _, out, _ = _gn(c, _conv(c, cbuf, br.proposal_conv[0]), br.proposal_conv[1], tag + "pp", segments=segs)
        if need_grad:
            sv[tag + "cbuf"], sv[tag + "lr_y"] = cbuf, lr
        return out, lr_t

    if par:
        with ops.parallel_branch():
            conf_prop, conf_lr_t = branch(cp.conf_proposal_branch, conf_feat, "cb")
        loc_prop, loc_lr_t = branch(cp.loc_proposal_branch, loc_feat, "lb")
        ops.join_branch()
    else:
        loc_prop, loc_lr_t = branch(cp.loc_proposal_branch, loc_feat, "lb")
        conf_prop, conf_lr_t = branch(cp.conf_proposal_branch, conf_feat, "cb")
    nd = loc_lr_t.shape[2] // 2

    # ---- refined heads (BDNet.py:399-412)
    pheads = [(cp.prop_loc_head, loc_prop, 2), (cp.prop_conf_head, conf_prop, cp.num_classes)]
    if cp.os_head:
        pheads.append((cp.prop_actionness_head, conf_prop, 1))
    pheads.append((cp.center_head, loc_prop, 1))
    praws = [_conv(c, f, u) for u, f, _ in pheads]
    pouts = ops.head_gather_fwd(praws, [h[2] for h in pheads], [0] * len(pheads), [h[0].conv1d.bias.detach() for h in pheads],
                                st["sep_idx"])
    out = dict(loc=loc, conf=conf, priors=tb["prior"], prop_loc=pouts[0], prop_conf=pouts[1], center=pouts[-1], start=start, end=end,
               start_loc_prop=loc_lr_t[:, :, :nd], end_loc_prop=loc_lr_t[:, :, nd:], start_conf_prop=conf_lr_t[:, :, :nd],
               end_conf_prop=conf_lr_t[:, :, nd:], act=act, prop_act=pouts[2] if cp.os_head else None)
    if not need_grad:
        return out, None
    sv.update(pin=pin, ppl=ppl, x_sep=x_sep, frame_in=frame_in, d1=d1, d2=d2, frame=frame, lt1=lt1, ct1=ct1, loc_feat=loc_feat,
              conf_feat=conf_feat, heads=heads, raws=raws, outs=outs, pheads=pheads, praws=praws, seg_sep=seg_sep, fseg_sep=fseg_sep,
              pooled_p=pooled_p, loc_prop=loc_prop, conf_prop=conf_prop, B=B, x1_shape=None if x1 is None else tuple(x1.shape),
              x2_shape=tuple(x2.shape), forced=forced_segments is not None)
    if forced_segments is None:
        # the cached window buffers are overwritten by the next forward: a backward that follows a later forward (the SSL pass
        # runs on another path, evaluation under no_grad does not save) must see its own windows
        sv["seg_sep"], sv["fseg_sep"] = seg_sep.clone(), fseg_sep.clone()
    return out, c


# ------------------------------------------------------------------------------------------------------------------
# backward
# ------------------------------------------------------------------------------------------------------------------
def backward(c: _Ctx, grads: dict):
    """grads: output name -> gradient tensor or None.  Returns (g_x1 | None, g_x2) in the NCDHW-view shapes of the inputs;
    every parameter gradient is accumulated in place."""
    cp, sv, dev = c.cp, c.sv, c.dev
    st = _tables(cp, dev)
    tb = cp._tables_on(dev)
    thumos = cp.variant == "thumos"
    L, S, t = cp.layer_num, cp.sep_len, cp.level_t
    segs = cp.sep_segments
    B = sv["B"]
    c.forked = False
    c.store.bind_grads()
    compat = BoundaryMaxPoolingFunction.compat_tscale_bug

    def g(name):
        v = grads.get(name)
        return v.contiguous() if v is not None else None

    # ---- refined heads
    pheads, praws = sv["pheads"], sv["praws"]
    pg = [g("prop_loc"), g("prop_conf")] + ([g("prop_act")] if cp.os_head else []) + [g("center")]
    pd = ops.head_gather_bwd(praws, [h[2] for h in pheads], [0] * len(pheads), [h[0].conv1d.bias.detach() for h in pheads],
                             [_grad(h[0].conv1d.bias) for h in pheads], pg, [None] * len(pheads), st["sep_idx"], st["prior_of_col"],
                             with_lo=c.with_lo)
    d_prop = {}
    for (unit, feat, _), dpl in zip(pheads, pd):
        key = id(feat)
        d_prop[key] = _conv_bwd(c, feat, dpl, unit, gx=d_prop.get(key), accumulate=key in d_prop)
    d_loc_prop, d_conf_prop = d_prop[id(sv["loc_prop"])], d_prop[id(sv["conf_prop"])]

    # ---- proposal branches and coarse heads: the conf half on the branch stream next to the loc half
    par = bool(getattr(cp, "two_streams", True)) and ops.OVERLAP_WGRAD
    d_feat = {}
    d_pooled = None
    heads, raws, outs = sv["heads"], sv["raws"], sv["outs"]
    hg = [g("loc"), g("conf")] + ([g("act")] if cp.os_head else [])
    scales = [h.scale.detach() for h in cp.loc_heads]
    hd = ops.head_gather_bwd(raws, [h[2] for h in heads], [h[3] for h in heads], [h[0].conv1d.bias.detach() for h in heads],
                             [_grad(h[0].conv1d.bias) for h in heads], hg, outs, st["sep_idx"], st["prior_of_col"], st["level_id"],
                             tb["stride"] if cp.variant == "anet" else None, scales, [_grad(h.scale) for h in cp.loc_heads],
                             with_lo=c.with_lo)

    def branch_bwd(br, tag, feat, d_out, g_start, g_end, defer_roi):
        """One proposal branch + the coarse heads that read the same tower feature.  Returns the roi conv's output-gradient planes
        when its data gradient (into the buffer both branches share) is left to the caller."""
        nonlocal d_pooled
        cbuf = sv.pop(tag + "cbuf")
        q = cbuf.hi.shape[-1] // 4
        dpl, _ = _gn_bwd(c, tag + "pp", br.proposal_conv[1], d_out, br.proposal_conv[0])
        d_cbuf = _conv_bwd(c, cbuf, dpl, br.proposal_conv[0])                               # [B,4q,S]
        # roi part -> the shared pooled frame feature
        dpl_roi, _ = _gn_bwd(c, tag + "roi", br.roi_conv[1], d_cbuf, br.roi_conv[0], gy_coff=0)
        if defer_roi:
            _conv_bwd(c, sv["pooled_p"], dpl_roi, br.roi_conv[0], want_dgrad=False)
        else:
            d_pooled = _conv_bwd(c, sv["pooled_p"], dpl_roi, br.roi_conv[0], gx=d_pooled, accumulate=d_pooled is not None)
        # boundary part -> BoundaryMaxPooling backward -> lr_conv
        lr = sv.pop(tag + "lr_y")
        d_lr = ops.bmp_backward(d_cbuf[:, q:3 * q].contiguous(), lr, sv["seg_sep"], compat)
        dpl, _ = _gn_bwd(c, tag + "lr", br.lr_conv[1], d_lr, br.lr_conv[0], gy2=(g_start, g_end) if (g_start is not None or g_end is not None) else None,
                         gy2_off=segs[0][0])
        key = id(feat)
        d_feat[key] = _conv_bwd(c, feat, dpl, br.lr_conv[0])
        # centre part
        dpl, _ = _gn_bwd(c, tag + "cp", br.cur_point_conv[1], d_cbuf, br.cur_point_conv[0], gy_coff=3 * q)
        d_feat[key] = _conv_bwd(c, feat, dpl, br.cur_point_conv[0], gx=d_feat[key], accumulate=True)
        # the coarse heads on this tower feature
        for (unit, f, _, _), dp_head in zip(heads, hd):
            if f is feat:
                d_feat[key] = _conv_bwd(c, feat, dp_head, unit, gx=d_feat[key], accumulate=True)
        c.keep.append((d_cbuf, d_lr, d_out))
        return dpl_roi

    if par:
        with ops.parallel_branch():
            roi_c = branch_bwd(cp.conf_proposal_branch, "cb", sv["conf_feat"], d_conf_prop, g("start_conf_prop"), g("end_conf_prop"), True)
        branch_bwd(cp.loc_proposal_branch, "lb", sv["loc_feat"], d_loc_prop, g("start_loc_prop"), g("end_loc_prop"), False)
        ops.join_branch()
        d_pooled = _conv_bwd(c, sv["pooled_p"], roi_c, cp.conf_proposal_branch.roi_conv[0], gx=d_pooled, accumulate=True, want_wgrad=False)
    else:
        branch_bwd(cp.loc_proposal_branch, "lb", sv["loc_feat"], d_loc_prop, g("start_loc_prop"), g("end_loc_prop"), False)
        branch_bwd(cp.conf_proposal_branch, "cb", sv["conf_feat"], d_conf_prop, g("start_conf_prop"), g("end_conf_prop"), False)

    # ---- frame-level feature on the branch stream (pooling backward, deconv chain), towers on the main stream
    dc = cp.deconv
    gs, ge = g("start"), g("end")

    def frame_bwd():
        frame, fseg = sv["frame"], sv["fseg_sep"]
        if not compat:
            d_frame = ops.bmp_backward(d_pooled, frame, fseg, False)
        else:       # the reference backward's tscale quirk depends on the per-level call shape (K = t of the level)
            d_frame = None
            for off, tl in segs:
                gi = ops.bmp_backward(d_pooled[:, :, off:off + tl].contiguous(), frame, fseg[:, off:off + tl].contiguous(), True)
                d_frame = gi if d_frame is None else d_frame + gi
        dpl, _ = _gn_bwd(c, "dc3", dc[7], d_frame, dc[6], gy2=(gs, ge) if (gs is not None or ge is not None) else None, gy2_off=0)
        d_d2 = _conv_bwd(c, sv["d2"], dpl, dc[6])
        dpl, _ = _gn_bwd(c, "dc2", dc[4], d_d2, dc[3])
        d_d1 = _conv_bwd(c, sv["d1"], dpl, dc[3])
        dpl, _ = _gn_bwd(c, "dc1", dc[1], d_d1, dc[0])
        c.keep.append((d_frame, d_d2, d_d1))
        return _conv_bwd(c, sv["frame_in"], dpl, dc[0])

    if par:
        with ops.parallel_branch():
            d_frame_in = frame_bwd()
    d_x_sep = None
    for tw, tag, first, feat in ((cp.loc_tower, "lt", sv["lt1"], sv["loc_feat"]), (cp.conf_tower, "ct", sv["ct1"], sv["conf_feat"])):
        dpl, _ = _gn_bwd(c, tag + "2", tw[1][1], d_feat[id(feat)], tw[1][0])
        d_first = _conv_bwd(c, first, dpl, tw[1][0])
        dpl, _ = _gn_bwd(c, tag + "1", tw[0][1], d_first, tw[0][0])
        d_x_sep = _conv_bwd(c, sv["x_sep"], dpl, tw[0][0], gx=d_x_sep, accumulate=d_x_sep is not None)
        c.keep.append(d_first)
    if par:
        ops.join_branch()
    else:
        d_frame_in = frame_bwd()

    # ---- pyramid: transposes of the sep layout / upsampling / top-down add, then the conv chain from the top level down
    pin = sv["pin"]
    d_p = [None] * L
    d_p[0], _ = ops.rows_combine([d_x_sep, d_frame_in], st["d_lvl"][0], want_f32=True)
    for i in range(1, L):
        srcs = [d_x_sep, d_frame_in, d_p[0]] if (i == 1 and thumos) else [d_x_sep]
        d_p[i], _ = ops.rows_combine(srcs, st["d_lvl"][i], want_f32=True)
    g_x1 = g_x2 = None
    for i in range(L - 1, -1, -1):
        unit, gn = cp.pyramids[i][0], cp.pyramids[i][1]
        src3d = i == 0 or (i == 1 and thumos)
        if src3d:
            dpl, _ = _gn_bwd(c, f"pyr{i}", gn, d_p[i], unit, conv3d=True)
            xin = pin[i]
            gxf = torch.empty((B, xin.hi.shape[1], 1, 1, xin.hi.shape[-1]), dtype=torch.float32, device=dev)
            _conv_bwd(c, xin, dpl, unit, ndhwc_out=gxf)
            if i == 0 and thumos:
                g_x1 = gxf
            else:
                g_x2 = gxf
        else:
            dpl, gxf32 = _gn_bwd(c, f"pyr{i}", gn, d_p[i], unit, want_gx=True)
            T_in = pin[i].hi.shape[1]
            dil = _dilated(st, gxf32, T_in, c.with_lo)
            _conv_bwd(c, pin[i], dpl, unit, stride=2, gx=d_p[i - 1], accumulate=True, dp_dgrad=dil)
    if c.forked:
        ops.join()
    c.keep.clear()
    sv.clear()

    def as_view(gx, shape):
        if gx is None or shape is None:
            return None
        Bx, C, T, H, W = shape
        return gx.view(Bx, T, H, W, C).permute(0, 4, 1, 2, 3)

    return as_view(g_x1, sv_shape(c, "x1")), as_view(g_x2, sv_shape(c, "x2"))


def sv_shape(c: _Ctx, which: str):
    return c.shapes.get(which)


OUT_KEYS = ("loc", "conf", "prop_loc", "prop_conf", "center", "start", "end", "start_loc_prop", "end_loc_prop", "start_conf_prop",
            "end_conf_prop", "act", "prop_act")


class _HeadFn(torch.autograd.Function):
    """Autograd boundary of the explicit head schedule: inputs are the two backbone feature maps, outputs the 13 tensors of the
    reference's output dict (None entries are returned as empty tensors).  Parameter gradients are accumulated in place by the
    kernels, like the backbone's."""

    @staticmethod
    def forward(ctx, cp, x1, x2, forced_segments, anchor):
        need = any(ctx.needs_input_grad)
        out, state = forward(cp, x1, x2, forced_segments, need)
        ctx.state = state
        if state is not None:
            state.shapes = dict(x1=None if x1 is None else tuple(x1.shape), x2=tuple(x2.shape))
        ctx.present = [out[k] is not None for k in OUT_KEYS]
        ctx.priors = out["priors"]
        res = tuple(out[k] if out[k] is not None else x2.new_zeros(0) for k in OUT_KEYS)
        return res

    @staticmethod
    def backward(ctx, *gs):
        state = ctx.state
        if state is None:
            raise RuntimeError("head activations were not saved")
        ctx.state = None
        grads = {k: (gv if ok else None) for k, gv, ok in zip(OUT_KEYS, gs, ctx.present)}
        g1, g2 = backward(state, grads)
        return None, g1, g2, None, None


def run(cp, feat_dict, forced_segments=None) -> dict:
    """CoarsePyramid.forward(feat_dict) through the explicit schedule."""
    x1, x2 = feat_dict.get("Mixed_4f") if cp.variant == "thumos" else None, feat_dict["Mixed_5c"]
    tb = cp._tables_on(x2.device)
    if torch.is_grad_enabled() and any(p.requires_grad for p in cp.parameters()):
        anchor = cp._sched_anchor(x2.device)
    else:
        anchor = x2.new_zeros(0)
    res = _HeadFn.apply(cp, x1, x2, forced_segments, anchor)
    out = {k: (v if v.numel() or k in ("loc", "conf") else None) for k, v in zip(OUT_KEYS, res)}
    out["priors"] = tb["prior"]
    for k in ("act", "prop_act"):
        if not cp.os_head:
            out[k] = None
    return out
