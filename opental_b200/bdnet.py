"""BDNet (THUMOS14 flavour) with the reference's module API and state_dict, on the sm_100a kernels.

Call-compatible with `AFSD/thumos14/BDNet.py:435-535`:
    BDNet(in_channels=3, backbone_model=None, training=True, use_edl=False, use_rpl=False)
    .forward(x, proposals=None, ssl=False, get_feat=False) -> dict(loc, conf, priors, prop_loc, prop_conf, center,
        start, end, start_loc_prop, end_loc_prop, start_conf_prop, end_conf_prop, act, prop_act[, unct, prop_unct])
The reference reads num_classes / os_head / evidence / ... from a global config evaluated at import time
(BDNet.py:12-18); here they are keyword arguments, and `BDNet.from_config(cfg_dict, ...)` maps the same yaml keys.

Backbone: `I3DBackbone` (opental_b200/backbone.py) — every Unit3D, max-pool, their gradients and the concat are
hand-written sm_100a kernels.  Head (CoarsePyramid, BDNet.py:117-432): every Unit1D / head-side Unit3D runs on the same
tensor-core implicit-GEMM kernels (headconv.py), GroupNorm+ReLU, window generation, BoundaryMaxPooling and the Dirichlet
uncertainty are native kernels; torch only glues them (index_select / cat / exp / autograd bookkeeping).  Two structural
changes against the reference's per-level Python loop, both value-preserving:
  * the towers, heads and proposal branches share their weights across the 6 pyramid levels, so they run ONCE on all levels
    laid side by side (123 -> 28 conv calls per forward);
  * the duplicated frame-level pooling of the two proposal branches (BDNet.py:109 called from :386 and :388 with identical
    arguments) is computed once.
`variant='anet'` selects the ActivityNet flavour (AFSD/anet/BDNet.py).

`use_rpl`, `get_feat`, the TransformerHead and dropout > 0 are baseline / ablation variants that are off in every
OpenTAL config: NotImplementedError.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .backbone import I3DBackbone
from .headconv import HeadConvStore, head_conv
from .prop_pooling import BoundaryMaxPooling, BoundaryMaxPoolingFunction

LAYER_NUM = 6        # BDNet.py:20
CONV_CHANNELS = 512  # BDNet.py:21


class _no_tf32:
    """Scope for the (non-native) library convolutions of the head: they must not silently drop to TF32 (torch's default for
    cuDNN convs) — single-pass TF32 alone costs ~1e-3 relative at the outputs, the whole error budget of the path.
    Scoped, so that building a BDNet does not change the precision of anything else in the process."""

    def __enter__(self):
        self._saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self._saved
        return False


class Unit1D(nn.Module):
    """conv1d with TF-style "same" padding and bias, no activation (AFSD/common/layers.py:178-214)."""

    def __init__(self, in_channels, output_channels, kernel_shape=1, stride=1):
        super().__init__()
        self.conv1d = nn.Conv1d(in_channels, output_channels, kernel_shape, stride, padding=0, bias=True)
        self._k, self._s = kernel_shape, stride
        self._native = None      # (HeadConvStore, record) once CoarsePyramid has registered this conv

    def forward(self, x):
        if self._native is not None:
            store, rec = self._native
            return head_conv(x, self.conv1d.weight, self.conv1d.bias, store, rec, self._s)
        t = x.shape[2]
        total = max(self._k - self._s, 0) if t % self._s == 0 else max(self._k - t % self._s, 0)
        if total:
            x = F.pad(x, [total // 2, total - total // 2])
        with _no_tf32():
            return self.conv1d(x)


class Unit3DValid(nn.Module):
    """Head-side Unit3D, padding='spatial_valid' with a full-extent spatial kernel (layers.py:143-175): a GEMM that
    collapses H x W."""

    def __init__(self, in_channels, output_channels, kernel_shape):
        super().__init__()
        assert kernel_shape[0] == 1
        self.conv3d = nn.Conv3d(in_channels, output_channels, kernel_shape, bias=True)
        self._native = None

    def forward(self, x):
        """x: NCDHW view of the backbone's channels-last feature map; returns [B,Cout,T] (H, W collapsed)."""
        if self._native is not None:
            store, rec = self._native
            return head_conv(x.permute(0, 2, 3, 4, 1), self.conv3d.weight, self.conv3d.bias, store, rec, 1)
        with _no_tf32():
            return self.conv3d(x).squeeze(-1).squeeze(-1)


class GroupNormReLU(nn.GroupNorm):
    """nn.GroupNorm(32, C) fused with the ReLU that follows it everywhere in CoarsePyramid (BDNet.py:72-73 etc.): one
    native launch (opental_b200/csrc/gn.cu).  Same parameters / state_dict keys as nn.GroupNorm."""

    def forward(self, x, segments=None):
        return ops.groupnorm_relu(x, self.weight, self.bias, self.num_groups, self.eps, relu=True, segments=segments)


def _conv_gn(seq, x, segments=None):
    """Apply a `_unit_gn` block; `segments` = per-level column ranges when x holds several pyramid levels side by side."""
    return seq[1](seq[0](x), segments)


def _unit_gn(unit, channels):
    # index 1 keeps the reference's state_dict keys ('<seq>.1.weight/bias'); index 2 was the in-place ReLU
    return nn.Sequential(unit, GroupNormReLU(32, channels), nn.Identity())


class ScaleExp(nn.Module):
    """exp(x * scale) (BDNet.py:55-61)."""

    def __init__(self, init_value=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.FloatTensor([init_value]))

    def forward(self, x):
        return torch.exp(x * self.scale)


class ProposalBranch(nn.Module):
    """BDNet.py:64-113, level-batched: `feature` [B,C,P] holds all pyramid levels side by side (P = 126 priors), every
    conv here is 1x1, GroupNorm statistics stay per level (`level_segments`), `segments` are windows into that
    concatenated axis (already clamped to their level).  Takes the pooled frame-level feature, which both branches
    share (the reference pools it twice with identical arguments, BDNet.py:109 called from :386 and :388)."""

    def __init__(self, in_channels, proposal_channels):
        super().__init__()
        self.cur_point_conv = _unit_gn(Unit1D(in_channels, proposal_channels, 1), proposal_channels)
        self.lr_conv = _unit_gn(Unit1D(in_channels, proposal_channels * 2, 1), proposal_channels * 2)
        self.boundary_max_pooling = BoundaryMaxPooling()
        self.roi_conv = _unit_gn(Unit1D(proposal_channels, proposal_channels, 1), proposal_channels)
        self.proposal_conv = _unit_gn(Unit1D(proposal_channels * 4, in_channels, 1), in_channels)

    def forward(self, feature, pooled_frame_feature, segments, level_segments=None):
        fm_short = _conv_gn(self.cur_point_conv, feature, level_segments)
        feature = _conv_gn(self.lr_conv, feature, level_segments)
        prop_feature = self.boundary_max_pooling(feature, segments)
        prop_roi_feature = _conv_gn(self.roi_conv, pooled_frame_feature, level_segments)
        prop_feature = torch.cat([prop_roi_feature, prop_feature, fm_short], dim=1)
        return _conv_gn(self.proposal_conv, prop_feature, level_segments), feature


class _FramePoolFn(torch.autograd.Function):
    """BoundaryMaxPooling of the frame-level feature for the windows of ALL levels in one launch.  In the parity mode
    that reproduces the reference backward's tscale quirk (boundary_max_pooling_kernel.cu:121) the quirk depends on the
    per-level call shape (K = t of the level), so that mode falls back to one backward launch per level."""

    @staticmethod
    def forward(ctx, frame, frame_segments, level_segments):
        ctx.save_for_backward(frame, frame_segments)
        ctx.level_segments = level_segments
        return ops.bmp_forward(frame, frame_segments)

    @staticmethod
    def backward(ctx, g):
        frame, fs = ctx.saved_tensors
        g = g.contiguous()
        if not BoundaryMaxPoolingFunction.compat_tscale_bug:
            return ops.bmp_backward(g, frame, fs, False), None, None
        total = None
        for off, t in ctx.level_segments:
            gi = ops.bmp_backward(g[:, :, off:off + t].contiguous(), frame, fs[:, off:off + t].contiguous(), True)
            total = gi if total is None else total + gi
        return total, None, None


ANET_FPN_STRIDES = (4, 8, 16, 32, 64, 128)      # AFSD/anet/BDNet.py:20


class CoarsePyramid(nn.Module):
    """BDNet.py:117-432 (OpenTAL configuration: no RPL head, no transformer, dropout 0).

    variant='anet' is the ActivityNet flavour (AFSD/anet/BDNet.py:120-391): the pyramid starts from Mixed_5c alone
    (:130-155, :281-289), feat_t = frame_num // 8 (:18-21), `loc` is multiplied by the level's FPN stride (:307-311) and
    the priors carry the level index as a second column (:262-269)."""

    def __init__(self, feat_channels, num_classes, frame_num=256, os_head=False, precision="bf16x3", native_convs=True,
                 variant="thumos"):
        super().__init__()
        assert variant in ("thumos", "anet")
        out_channels = CONV_CHANNELS
        self.variant = variant
        self.num_classes = num_classes
        self.frame_num = frame_num
        self.os_head = os_head
        self.layer_num = LAYER_NUM
        self.pyramids = nn.ModuleList()
        if variant == "anet":
            feat_t = frame_num // 8                  # temporal extent of Mixed_5c
            self.pyramids.append(_unit_gn(Unit3DValid(feat_channels[1], out_channels, [1, 3, 3]), out_channels))
            n_src = 1
        else:
            feat_t = frame_num // 4                  # temporal extent of Mixed_4f; Mixed_5c has half of it
            assert (frame_num // 8) * 2 == feat_t
            self.pyramids.append(_unit_gn(Unit3DValid(feat_channels[0], out_channels, [1, 6, 6]), out_channels))
            self.pyramids.append(_unit_gn(Unit3DValid(feat_channels[1], out_channels, [1, 3, 3]), out_channels))
            n_src = 2
        for _ in range(n_src, LAYER_NUM):
            self.pyramids.append(_unit_gn(Unit1D(out_channels, out_channels, 3, stride=2), out_channels))
        self.loc_heads = nn.ModuleList([ScaleExp() for _ in range(LAYER_NUM)])

        def tower():
            return nn.Sequential(*[_unit_gn(Unit1D(out_channels, out_channels, 3), out_channels) for _ in range(2)])

        self.loc_tower, self.conf_tower = tower(), tower()
        self.loc_head = Unit1D(out_channels, 2, 3)
        self.conf_head = Unit1D(out_channels, num_classes, 3)
        if os_head:
            self.actionness_head = Unit1D(out_channels, 1, 3)
        self.loc_proposal_branch = ProposalBranch(out_channels, 512)
        self.conf_proposal_branch = ProposalBranch(out_channels, 512)
        self.prop_loc_head = Unit1D(out_channels, 2, 1)
        self.prop_conf_head = Unit1D(out_channels, num_classes, 1)
        if os_head:
            self.prop_actionness_head = Unit1D(out_channels, 1, 1)
        self.center_head = Unit1D(out_channels, 1, 3)
        self.deconv = nn.Sequential(
            Unit1D(out_channels, out_channels, 3), GroupNormReLU(32, out_channels), nn.Identity(),
            Unit1D(out_channels, out_channels, 3), GroupNormReLU(32, out_channels), nn.Identity(),
            Unit1D(out_channels, out_channels, 1), GroupNormReLU(32, out_channels), nn.Identity())
        self.boundary_max_pooling = BoundaryMaxPooling()
        self.priors = []
        t = feat_t
        for i in range(LAYER_NUM):
            if variant == "anet":
                self.priors.append(torch.tensor([[(c + 0.5) / t, i] for c in range(t)], dtype=torch.float32).view(-1, 2))
            else:
                self.priors.append(torch.tensor([[(c + 0.5) / t] for c in range(t)], dtype=torch.float32).view(-1, 1))
            t = t // 2
        # Level-batched layouts.  The towers / heads / proposal branches share their weights across the 6 levels
        # (BDNet.py:333-412), so they run ONCE on all levels laid side by side along T:
        #   "sep" layout  [B,C,S]: [0 | L0 | 0 | L1 | 0 | ... | L5 | 0 ...] — one zero column between levels gives every
        #                 k=3 "same" convolution its per-level zero padding; S is padded to a multiple of 8
        #   "cat" layout  [B,C,P]: the 126 priors back to back (1x1 convs, pooling windows, outputs)
        self.level_t = [feat_t >> i for i in range(LAYER_NUM)]
        self.num_priors = sum(self.level_t)
        sep_off, pos = [], 1
        for t in self.level_t:
            sep_off.append(pos)
            pos += t + 1
        self.sep_len = (pos + 7) // 8 * 8
        self.sep_segments = tuple(zip(sep_off, self.level_t))
        cat_off = [sum(self.level_t[:i]) for i in range(LAYER_NUM)]
        self.cat_segments = tuple(zip(cat_off, self.level_t))
        self._tables = {}
        # the explicit forward / backward schedule (head_schedule.py: one autograd node for the whole head); OTAL_HEAD_SCHEDULE=0
        # selects the per-module autograd formulation below (the SSL / triplet pass always takes it)
        import os
        self.native_schedule = os.environ.get("OTAL_HEAD_SCHEDULE", "1") != "0"
        self._anchors = {}
        self.two_streams = os.environ.get("OTAL_HEAD_ONE_STREAM") is None      # loc / conf halves of the schedule side by side
        # every head conv on the tensor-core kernels, weights re-homed into one packed flat buffer (headconv.py);
        # native_convs=False keeps torch's library convs (debugging aid, never selected implicitly)
        self.conv_store = None
        if native_convs:
            self.conv_store = HeadConvStore(precision)
            for m in self.modules():
                if isinstance(m, Unit1D):
                    m._native = (self.conv_store, self.conv_store.register(m.conv1d.weight, m.conv1d.bias, "conv1d"))
                elif isinstance(m, Unit3DValid):
                    m._native = (self.conv_store, self.conv_store.register(m.conv3d.weight, m.conv3d.bias, "valid3d"))

    def _tables_on(self, device):
        """Per-prior lookup tables on `device`: prior position, level length, level offset (cat layout), level id, and
        the column of every prior in the sep layout."""
        if device not in self._tables:
            lens = torch.tensor([t for t in self.level_t for _ in range(t)], dtype=torch.int32)
            offs = torch.tensor([o for (o, t) in self.cat_segments for _ in range(t)], dtype=torch.int32)
            lid = torch.tensor([i for i, t in enumerate(self.level_t) for _ in range(t)], dtype=torch.long)
            sep = torch.cat([torch.arange(o, o + t) for o, t in self.sep_segments])
            prior = torch.cat(self.priors, 0)
            stride = torch.tensor([ANET_FPN_STRIDES[i] for i in lid.tolist()], dtype=torch.float32)
            self._tables[device] = dict(priors=[p.to(device) for p in self.priors], prior=prior.to(device),
                                        centre=prior[:, 0].contiguous().to(device), level_len=lens.to(device),
                                        level_off=offs.to(device), level_id=lid.to(device), sep_idx=sep.to(device),
                                        stride=stride.to(device))
        return self._tables[device]

    def _sched_anchor(self, device):
        """A differentiable dummy input that makes autograd visit the head's single node even when the features need no gradient."""
        if device not in self._anchors:
            self._anchors[device] = torch.zeros(1, device=device, requires_grad=True)
        return self._anchors[device]

    def _segments(self, loc, tb):
        """Window generation (no_grad, BDNet.py:355-384) for all levels: one native launch."""
        _, seg_cat, frame_seg = ops.make_segments(loc, tb["centre"], tb["level_len"], tb["level_off"], self.frame_num)
        return seg_cat, frame_seg

    def _forced(self, forced_segments):
        """Per-level (segments, frame_segments) as the reference produces them -> the level-concatenated form."""
        segs, fsegs = [], []
        for (seg, fseg), (off, t) in zip(forced_segments, self.cat_segments):
            segs.append(torch.trunc(seg).clamp(0, t - 1) + off)
            fsegs.append(fseg)
        return torch.cat(segs, 1).contiguous(), torch.cat(fsegs, 1).contiguous()

    def forward(self, feat_dict, ssl=False, get_feat=False, forced_segments=None):
        if get_feat:
            raise NotImplementedError("get_feat is an analysis-only path (SURVEY D10)")
        if self.native_schedule and self.conv_store is not None and not ssl:
            from . import head_schedule
            return head_schedule.run(self, feat_dict, forced_segments)
        x1, x2 = feat_dict.get("Mixed_4f"), feat_dict["Mixed_5c"]
        if self.conv_store is not None:
            self.conv_store.prepare(x2.device)
        B = x2.size(0)
        K = self.num_classes
        tb = self._tables_on(x2.device)
        # ---- pyramid (BDNet.py:311-322; anet/BDNet.py:281-289) and frame-level feature (:324-331)
        feats = []
        if self.variant == "anet":
            x = None
            for i, conv in enumerate(self.pyramids):
                x = conv(x2) if i == 0 else conv(x)
                feats.append(x)
        else:
            for i, conv in enumerate(self.pyramids):
                if i == 0:
                    x = conv(x1)
                elif i == 1:
                    x = conv(x2)
                    feats[-1] = feats[-1] + F.interpolate(x, feats[-1].shape[2:], mode="nearest")
                else:
                    x = conv(x)
                feats.append(x)
        frame = F.interpolate(feats[0].unsqueeze(-1), [self.frame_num, 1]).squeeze(-1)
        frame = self.deconv(frame).contiguous()
        if ssl:
            # BDNet.py:397-400: the SSL / triplet pass only needs the frame-level feature and the boundary (lr_conv)
            # features of level 0 — run the towers on that level alone
            loc0, conf0 = feats[0], feats[0]
            for blk in self.loc_tower:
                loc0 = _conv_gn(blk, loc0)
            for blk in self.conf_tower:
                conf0 = _conv_gn(blk, conf0)
            return [frame.clone(), _conv_gn(self.loc_proposal_branch.lr_conv, loc0),
                    _conv_gn(self.conf_proposal_branch.lr_conv, conf0)]
        start = frame[:, :256].permute(0, 2, 1).contiguous()
        end = frame[:, 256:].permute(0, 2, 1).contiguous()

        # ---- towers and coarse heads on all levels at once (sep layout), BDNet.py:333-353
        z = feats[0].new_zeros(B, feats[0].shape[1], 1)
        parts = [z]
        for f in feats:
            parts += [f, z]
        tail = self.sep_len - (self.num_priors + LAYER_NUM + 1)
        if tail:
            parts.append(feats[0].new_zeros(B, feats[0].shape[1], tail))
        x_sep = torch.cat(parts, dim=2)
        loc_feat, conf_feat = x_sep, x_sep
        for blk in self.loc_tower:
            loc_feat = _conv_gn(blk, loc_feat, self.sep_segments)
        for blk in self.conf_tower:
            conf_feat = _conv_gn(blk, conf_feat, self.sep_segments)
        sep_idx = tb["sep_idx"]

        def to_out(y):                                  # [B,c,P] -> the reference's [B,P,c]
            return y.permute(0, 2, 1).contiguous()

        scale = torch.cat([h.scale for h in self.loc_heads])[tb["level_id"]]            # ScaleExp of the prior's level
        loc = torch.exp(self.loc_head(loc_feat).index_select(2, sep_idx) * scale)
        if self.variant == "anet":
            loc = loc * tb["stride"]                                                        # anet/BDNet.py:307-311
        loc = to_out(loc)
        conf = to_out(self.conf_head(conf_feat).index_select(2, sep_idx))
        act = to_out(self.actionness_head(conf_feat).index_select(2, sep_idx)) if self.os_head else None

        # ---- proposal windows and the two proposal branches (cat layout), BDNet.py:355-397
        if forced_segments is not None:
            segments, frame_segments = self._forced(forced_segments)
        else:
            segments, frame_segments = self._segments(loc, tb)
        pooled_frame = _FramePoolFn.apply(frame, frame_segments, self.cat_segments)     # shared by both branches (F5)
        loc_cat = loc_feat.index_select(2, sep_idx)
        conf_cat = conf_feat.index_select(2, sep_idx)
        loc_prop, loc_lr = self.loc_proposal_branch(loc_cat, pooled_frame, segments, self.cat_segments)
        conf_prop, conf_lr = self.conf_proposal_branch(conf_cat, pooled_frame, segments, self.cat_segments)
        t0 = self.level_t[0]
        nd = loc_lr.size(1) // 2
        extra = dict(start_loc_prop=loc_lr[:, :nd, :t0].permute(0, 2, 1).contiguous(),
                     end_loc_prop=loc_lr[:, nd:, :t0].permute(0, 2, 1).contiguous(),
                     start_conf_prop=conf_lr[:, :nd, :t0].permute(0, 2, 1).contiguous(),
                     end_conf_prop=conf_lr[:, nd:, :t0].permute(0, 2, 1).contiguous())
        # ---- refined heads (BDNet.py:399-412): 1x1 heads on the cat layout; the k=3 centre head needs the separators
        prop_loc = to_out(self.prop_loc_head(loc_prop))
        prop_conf = to_out(self.prop_conf_head(conf_prop))
        prop_act = to_out(self.prop_actionness_head(conf_prop)) if self.os_head else None
        lp_sep = loc_prop.new_zeros(B, loc_prop.shape[1], self.sep_len).index_copy(2, sep_idx, loc_prop)
        center = to_out(self.center_head(lp_sep).index_select(2, sep_idx))
        return dict(loc=loc, conf=conf, priors=tb["prior"], prop_loc=prop_loc, prop_conf=prop_conf, center=center,
                    start=start, end=end, **extra, act=act, prop_act=prop_act)


class DirichletLayer(nn.Module):
    """BDNet.py:538-561."""

    def __init__(self, evidence="exp", dim=-1):
        super().__init__()
        assert evidence in ("relu", "exp", "softplus")
        self.evidence, self.dim = evidence, dim

    def evidence_func(self, logit):
        if self.evidence == "relu":
            return F.relu(logit)
        if self.evidence == "exp":
            return torch.exp(torch.clamp(logit, -10, 10))
        return F.softplus(logit)

    def compute_uncertainty(self, logit):
        num_classes = logit.size(-1)
        alpha = self.evidence_func(logit) + 1
        return num_classes / alpha.sum(dim=self.dim)

    def forward(self, logit):
        alpha = self.evidence_func(logit) + 1
        return alpha / alpha.sum(dim=self.dim, keepdim=True)


class BDNet(nn.Module):
    def __init__(self, in_channels=3, backbone_model=None, training=True, use_edl=False, use_rpl=False, *,
                 num_classes=21, os_head=False, frame_num=256, evidence="exp", dropout=0.0, precision="bf16x3",
                 freeze_bn=True, freeze_bn_affine=True, native_head_convs=True, variant="thumos"):
        super().__init__()
        if use_rpl:
            raise NotImplementedError("the RPL head is a competing baseline, off in every OpenTAL config (SURVEY D10)")
        if dropout:
            raise NotImplementedError("dropout is 0 in every OpenTAL config (SURVEY D6)")
        self.os_head = os_head
        self.num_classes = num_classes - 1 if os_head else num_classes       # BDNet.py:440
        self.variant = variant
        self.coarse_pyramid_detection = CoarsePyramid([832, 1024], self.num_classes, frame_num, os_head, precision,
                                                      native_head_convs, variant)
        self.reset_params()
        self.backbone = I3DBackbone(in_channels, precision=precision, freeze_bn=freeze_bn,
                                    freeze_bn_affine=freeze_bn_affine)
        self.boundary_max_pooling = BoundaryMaxPooling()
        self._training = training
        if training and backbone_model is not None:
            self.load_pretrained_weight(backbone_model)
        self.scales = [1, 4, 4]
        self.use_edl = use_edl
        self.evidence = evidence
        if use_edl:
            self.out_layer = DirichletLayer(evidence, dim=-1)
        self.use_rpl = False

    @classmethod
    def from_config(cls, config: dict, **kw):
        """Map the reference's yaml keys (AFSD/common/config.py, BDNet.py:12-18) to constructor arguments; explicit
        keyword arguments (the reference's own call `BDNet(in_channels=..., backbone_model=..., use_edl=...)`,
        train.py:314) win over the config."""
        m = config.get("model", {})
        args = dict(in_channels=m.get("in_channels", 3), backbone_model=m.get("backbone_model"),
                    num_classes=config["dataset"]["num_classes"], os_head=m.get("os_head", False),
                    evidence=m.get("evidence", "exp"), dropout=m.get("dropout", 0.0),
                    frame_num=config["dataset"]["training"]["clip_length"],
                    freeze_bn=m.get("freeze_bn", True), freeze_bn_affine=m.get("freeze_bn_affine", True))
        args.update(kw)
        return cls(**args)

    def load_pretrained_weight(self, model_path):
        """I3D_BackBone.load_pretrained_weight (BDNet.py:35-37): Kinetics I3D state_dict, strict=False."""
        sd = torch.load(model_path, map_location="cpu")
        self.backbone._model.load_state_dict(sd, strict=False)

    @staticmethod
    def weight_init(m):
        """Glorot-uniform conv weights, zero bias (BDNet.py:460-473)."""
        if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Conv3d)):
            fan_in, fan_out = nn.init._calculate_fan_in_and_fan_out(m.weight)
            limit = math.sqrt(3.0 / max(1.0, (fan_in + fan_out) / 2.0))
            nn.init.uniform_(m.weight, -limit, limit)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def reset_params(self):
        for m in self.modules():
            self.weight_init(m)
        if getattr(self, "variant", "thumos") == "anet":
            # the ActivityNet flavour re-initialises the shared head convs with N(0, 0.01) (anet/BDNet.py:435-451)
            cp = self.coarse_pyramid_detection
            for mod in (cp.loc_tower, cp.conf_tower, cp.loc_head, cp.conf_head, cp.loc_proposal_branch, cp.conf_proposal_branch,
                        cp.prop_loc_head, cp.prop_conf_head, cp.center_head):
                for layer in mod.modules():
                    if isinstance(layer, nn.Conv1d):
                        nn.init.normal_(layer.weight, mean=0, std=0.01)
                        nn.init.constant_(layer.bias, 0)

    def forward(self, x, proposals=None, ssl=False, get_feat=False, forced_segments=None):
        feat_dict = self.backbone(x)
        if ssl:
            top_feat = self.coarse_pyramid_detection(feat_dict, ssl=True)
            # BDNet.py:482-503.  The reference indexes segments of sample 0 for every sample (only defined for batch 1,
            # SURVEY D11); here the proposals of sample 0 are broadcast explicitly.
            dec = proposals[0].unsqueeze(0)
            plen = dec[:, :, 1:] - dec[:, :, :1] + 1.0
            inl, outl = torch.clamp(plen / 4.0, min=1.0), torch.clamp(plen / 10.0, min=1.0)
            fs = torch.cat([torch.round(dec[:, :, :1] - outl), torch.round(dec[:, :, :1] + inl),
                            torch.round(dec[:, :, 1:] - inl), torch.round(dec[:, :, 1:] + outl)], dim=-1)
            anchor, positive, negative = [], [], []
            for i in range(3):
                seg = (fs / self.scales[i]).expand(top_feat[i].size(0), -1, -1).contiguous()
                bound = self.boundary_max_pooling(top_feat[i].contiguous(), seg)
                nd = bound.size(1) // 2
                anchor.append(bound[:, nd:, 0])
                positive.append(bound[:, :nd, 1])
                negative.append(bound[:, :nd, 2])
            return anchor, positive, negative
        out = self.coarse_pyramid_detection(feat_dict, get_feat=get_feat, forced_segments=forced_segments)
        if self.use_edl:
            if self.evidence == "exp":        # native kernel (no autograd graph: these are inference-side scores)
                out["unct"] = ops.dirichlet_uncertainty(out["conf"])
                out["prop_unct"] = ops.dirichlet_uncertainty(out["prop_conf"])
            else:
                out["unct"] = self.out_layer.compute_uncertainty(out["conf"])
                out["prop_unct"] = self.out_layer.compute_uncertainty(out["prop_conf"])
        return out
