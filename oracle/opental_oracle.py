"""CPU oracle for the OpenTAL hot path  —  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch, functional restatement (plain torch fp32/fp64 CPU ops driven by a state_dict) of the reference
algorithm on the path BDNet.forward -> MultiSegmentLoss (THUMOS14 flavour).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py` (cpu_baseline / --impl reference) may import this module, and only as
the checker / the timed CPU baseline — never as something the product path calls.

Pinning.  The reference ships no tests or golden vectors for this path, so the oracle is
pinned against the reference's own Python code imported from /root/reference in the build container:
`oracle/make_golden.py` runs both on identical keyed-synthetic weights/inputs, asserts agreement and writes the
fixtures under tests/golden/ that the CPU test-suite re-checks this file against.  The arithmetic below the
torch API (conv/group_norm/… on CPU, torch 2.11) is third-party and is the oracle-of-record.

Every function cites the reference lines it restates (paths relative to the OpenTAL repository).
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------------------
# architecture tables (AFSD/common/i3d_backbone.py:193-296, AFSD/thumos14/BDNet.py:20-22,117-293)
# ----------------------------------------------------------------------------------------------------------
I3D_ENDPOINTS = [
    # (name, kind, args)
    ("Conv3d_1a_7x7", "conv", dict(cin=3, cout=64, k=(7, 7, 7), s=(2, 2, 2))),
    ("MaxPool3d_2a_3x3", "pool", dict(k=(1, 3, 3), s=(1, 2, 2))),
    ("Conv3d_2b_1x1", "conv", dict(cin=64, cout=64, k=(1, 1, 1), s=(1, 1, 1))),
    ("Conv3d_2c_3x3", "conv", dict(cin=64, cout=192, k=(3, 3, 3), s=(1, 1, 1))),
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
BN_EPS = 1e-3      # i3d_backbone.py:43
GN_GROUPS = 32     # BDNet.py:72 etc.
GN_EPS = 1e-5
HEAD_CH = 512      # BDNet.py:21
NUM_LEVELS = 6     # BDNet.py:20


@dataclass
class OracleConfig:
    """The handful of config.py / yaml values the hot path reads (AFSD/thumos14/BDNet.py:12-18)."""

    num_classes: int = 15        # model classes: yaml num_classes - 1 when os_head (BDNet.py:440)
    os_head: bool = True
    use_edl: bool = True
    frame_num: int = 256         # clip length (BDNet.py:117; config dataset.training.clip_length)
    feat_t: int = 64             # BDNet.py:22
    clip_length: int = 256       # multisegment_loss.py:110
    piou: float = 0.5            # overlap_thresh (train.py:327)
    # edl_config (configs/thumos14_opental_final.yaml:38-49)
    with_ibm: bool = True
    ibm_start: int = 10
    momentum: float = 0.99
    num_bins: int = 50
    iou_aware: bool = True
    # act_config (:50-52)
    act_weight: float = 0.0
    act_margin: float = 1.0
    # 'thumos' (AFSD/thumos14) or 'anet' (AFSD/anet: single-source pyramid, loc x fpn stride, level-gated matching,
    # per-sample loss normalisation, exp-form IBM weight)
    variant: str = "thumos"
    ibm_coeff: float = 10.0      # anet/cls_loss.py:99


ANET_FPN_STRIDES = [4, 8, 16, 32, 64, 128]                                       # anet/BDNet.py:20
ANET_BOUNDS = [[0, 30], [15, 60], [30, 120], [60, 240], [96, 768], [256, 768]]  # anet/multisegment_loss.py:69


def anet_config() -> "OracleConfig":
    """configs/anet_opental.yaml --open_set: 151 - 1 classes (os_head), 768-frame clips, feat_t = 768 // 8
    (anet/BDNet.py:18-21), ActionnessLoss(weight=0.1) (anet/multisegment_loss.py:102)."""
    return OracleConfig(num_classes=150, frame_num=768, feat_t=96, clip_length=768, act_weight=0.1, variant="anet")


# ----------------------------------------------------------------------------------------------------------
# keyed synthetic weights (SURVEY §8d "Synthetic-data conventions")
# ----------------------------------------------------------------------------------------------------------
def model_spec(cfg: OracleConfig) -> list[tuple[str, tuple[int, ...], str]]:
    """(state_dict key, shape, kind) for every tensor of BDNet, in the reference's registration order.

    kind in {conv_w, conv_b, bn_w, bn_b, bn_mean, bn_var, bn_count, gn_w, gn_b, scale}."""
    out: list[tuple[str, tuple[int, ...], str]] = []
    cp = "coarse_pyramid_detection."

    def conv1d(prefix, cin, cout, k):
        out.append((prefix + "conv1d.weight", (cout, cin, k), "conv_w"))
        out.append((prefix + "conv1d.bias", (cout,), "conv_b"))

    def gn(prefix, c):
        out.append((prefix + "weight", (c,), "gn_w"))
        out.append((prefix + "bias", (c,), "gn_b"))

    # pyramids (BDNet.py:129-168; anet/BDNet.py:130-155: one source, Mixed_5c)
    sources = [(1024, (1, 3, 3))] if cfg.variant == "anet" else [(832, (1, 6, 6)), (1024, (1, 3, 3))]
    for i, (cin, k) in enumerate(sources):
        out.append((f"{cp}pyramids.{i}.0.conv3d.weight", (HEAD_CH, cin, *k), "conv_w"))
        out.append((f"{cp}pyramids.{i}.0.conv3d.bias", (HEAD_CH,), "conv_b"))
        gn(f"{cp}pyramids.{i}.1.", HEAD_CH)
    for i in range(len(sources), NUM_LEVELS):
        conv1d(f"{cp}pyramids.{i}.0.", HEAD_CH, HEAD_CH, 3)
        gn(f"{cp}pyramids.{i}.1.", HEAD_CH)
    for i in range(NUM_LEVELS):
        out.append((f"{cp}loc_heads.{i}.scale", (1,), "scale"))
    for tower in ("loc_tower", "conf_tower"):
        for i in range(2):
            conv1d(f"{cp}{tower}.{i}.0.", HEAD_CH, HEAD_CH, 3)
            gn(f"{cp}{tower}.{i}.1.", HEAD_CH)
    conv1d(f"{cp}loc_head.", HEAD_CH, 2, 3)
    conv1d(f"{cp}conf_head.", HEAD_CH, cfg.num_classes, 3)
    if cfg.os_head:
        conv1d(f"{cp}actionness_head.", HEAD_CH, 1, 3)
    for br in ("loc_proposal_branch", "conf_proposal_branch"):
        for name, cin, cout in (("cur_point_conv", HEAD_CH, 512), ("lr_conv", HEAD_CH, 1024),
                                ("roi_conv", 512, 512), ("proposal_conv", 2048, HEAD_CH)):
            conv1d(f"{cp}{br}.{name}.0.", cin, cout, 1)
            gn(f"{cp}{br}.{name}.1.", cout)
    conv1d(f"{cp}prop_loc_head.", HEAD_CH, 2, 1)
    conv1d(f"{cp}prop_conf_head.", HEAD_CH, cfg.num_classes, 1)
    if cfg.os_head:
        conv1d(f"{cp}prop_actionness_head.", HEAD_CH, 1, 1)
    conv1d(f"{cp}center_head.", HEAD_CH, 1, 3)
    for j, k in ((0, 3), (3, 3), (6, 1)):
        conv1d(f"{cp}deconv.{j}.", HEAD_CH, HEAD_CH, k)
        gn(f"{cp}deconv.{j + 1}.", HEAD_CH)

    # backbone (i3d_backbone.py:193-296; Unit3D = conv3d(no bias) + BatchNorm3d)
    def unit3d(prefix, cin, cout, k):
        out.append((prefix + "conv3d.weight", (cout, cin, *k), "conv_w"))
        out.append((prefix + "bn.weight", (cout,), "bn_w"))
        out.append((prefix + "bn.bias", (cout,), "bn_b"))
        out.append((prefix + "bn.running_mean", (cout,), "bn_mean"))
        out.append((prefix + "bn.running_var", (cout,), "bn_var"))
        out.append((prefix + "bn.num_batches_tracked", (), "bn_count"))

    bp = "backbone._model."
    for name, kind, a in I3D_ENDPOINTS:
        if kind == "conv":
            unit3d(f"{bp}{name}.", a["cin"], a["cout"], a["k"])
        elif kind == "mixed":
            w = a["widths"]
            unit3d(f"{bp}{name}.b0.", a["cin"], w[0], (1, 1, 1))
            unit3d(f"{bp}{name}.b1a.", a["cin"], w[1], (1, 1, 1))
            unit3d(f"{bp}{name}.b1b.", w[1], w[2], (3, 3, 3))
            unit3d(f"{bp}{name}.b2a.", a["cin"], w[3], (1, 1, 1))
            unit3d(f"{bp}{name}.b2b.", w[3], w[4], (3, 3, 3))
            unit3d(f"{bp}{name}.b3b.", a["cin"], w[5], (1, 1, 1))
    return out


def synthetic_state_dict(cfg: OracleConfig, loc_bias_shift: float = 0.0, dtype=torch.float32) -> dict[str, torch.Tensor]:
    """Deterministic weights keyed by crc32(state_dict key): identical for the reference, the oracle and the
    CUDA model regardless of construction order.  `loc_bias_shift=log(32)` makes predicted extents overlap the
    ground truth so the refined (prop_*) losses see positives."""
    sd: dict[str, torch.Tensor] = {}
    for key, shape, kind in model_spec(cfg):
        g = torch.Generator().manual_seed(zlib.crc32(key.encode()))
        if kind == "conv_w":
            fan_in = math.prod(shape[1:])
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif kind == "conv_b":
            t = 0.01 * torch.randn(shape, generator=g)
        elif kind in ("bn_w", "gn_w"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind in ("bn_b", "gn_b", "bn_mean"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_var":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g).abs()
        elif kind == "bn_count":
            t = torch.zeros((), dtype=torch.long)
        elif kind == "scale":
            t = torch.ones(shape)
        else:  # pragma: no cover
            raise AssertionError(kind)
        if t.is_floating_point():
            t = t.to(dtype)
        sd[key] = t
    if loc_bias_shift:
        sd["coarse_pyramid_detection.loc_head.conv1d.bias"] = sd["coarse_pyramid_detection.loc_head.conv1d.bias"] + loc_bias_shift
    return sd


def synthetic_clip(index: int, rank: int = 0, frames: int = 256, crop: int = 96) -> torch.Tensor:
    """uint8 [T,112,112,3] uniform pixels -> centre crop -> (x/255)*2-1 -> fp32 [3,T,crop,crop]
    (video2npy.py:61-74 format; thumos_dataset.py:261-263 normalisation)."""
    g = torch.Generator().manual_seed(1000 * rank + index)
    px = torch.randint(0, 256, (frames, 112, 112, 3), generator=g, dtype=torch.uint8)
    o = (112 - crop) // 2
    px = px[:, o:o + crop, o:o + crop, :]
    return (px.permute(3, 0, 1, 2).float() / 255.0) * 2.0 - 1.0


def synthetic_targets(index: int, rank: int = 0, num_classes: int = 15) -> torch.Tensor:
    """[2,3] rows (start, end, label) normalised to the clip, labels in 1..K (thumos_dataset.py:58-66)."""
    g = torch.Generator().manual_seed(7_000_000 + 1000 * rank + index)
    rows = []
    for j in range(2):
        s = 0.10 + 0.45 * j + (torch.rand((), generator=g).item() * 0.06 - 0.03)
        e = s + 0.25 + (torch.rand((), generator=g).item() * 0.06 - 0.03)
        lab = int(torch.randint(1, num_classes + 1, (), generator=g).item())
        rows.append([s, e, float(lab)])
    return torch.tensor(rows, dtype=torch.float32)


def synthetic_scores(targets: torch.Tensor, frames: int = 256) -> torch.Tensor:
    """start/end score maps [2,T]: ones inside a band of width d = max(len/10, 2) frames centred on each start / end
    boundary, python rounding, clipped to the clip — the loader's rule at thumos_dataset.py:109-120."""
    scores = torch.zeros(2, frames)
    for s, e, _ in targets.tolist():
        s_f, e_f = s * frames, e * frames
        d = max((e_f - s_f) / 10.0, 2.0)
        for row, centre in ((0, s_f), (1, e_f)):
            lo = min(max(int(round(centre - d / 2.0)), 0), frames - 1)
            hi = min(max(int(round(centre + d / 2.0)), 0), frames - 1)
            scores[row, lo:hi + 1] = 1.0
    return scores


# ----------------------------------------------------------------------------------------------------------
# op layer
# ----------------------------------------------------------------------------------------------------------
def same_pad(size: int, k: int, s: int) -> tuple[int, int]:
    """TF-style 'same' padding split (front, back): i3d_backbone.py:45-69 / layers.py:11-29,198-210."""
    total = max(k - s, 0) if size % s == 0 else max(k - size % s, 0)
    return total // 2, total - total // 2


def _pad3d(x, k, s):
    t, h, w = x.shape[2:]
    pt, ph, pw = same_pad(t, k[0], s[0]), same_pad(h, k[1], s[1]), same_pad(w, k[2], s[2])
    return F.pad(x, [pw[0], pw[1], ph[0], ph[1], pt[0], pt[1]])


def unit3d_bn_relu(x, sd, prefix, k, s=(1, 1, 1)):
    """Backbone Unit3D: same-pad -> conv3d (no bias) -> BatchNorm3d in eval mode (frozen) -> ReLU
    (i3d_backbone.py:51-87; BN frozen even in train mode: thumos14/BDNet.py:39-49)."""
    y = F.conv3d(_pad3d(x, k, s), sd[prefix + "conv3d.weight"], None, stride=s)
    inv = torch.rsqrt(sd[prefix + "bn.running_var"] + BN_EPS)
    scale = (sd[prefix + "bn.weight"] * inv).view(1, -1, 1, 1, 1)
    shift = (sd[prefix + "bn.bias"] - sd[prefix + "bn.running_mean"] * sd[prefix + "bn.weight"] * inv).view(1, -1, 1, 1, 1)
    return F.relu(y * scale + shift)


def maxpool3d_same(x, k, s):
    """Zero-pad (constant 0, not -inf) then max-pool: layers.py:9-35."""
    return F.max_pool3d(_pad3d(x, k, s), k, s)


def inception(x, sd, prefix, widths):
    """Four branches concatenated on channels in the order [b0|b1b|b2b|b3b]: i3d_backbone.py:116-121."""
    b0 = unit3d_bn_relu(x, sd, prefix + "b0.", (1, 1, 1))
    b1 = unit3d_bn_relu(unit3d_bn_relu(x, sd, prefix + "b1a.", (1, 1, 1)), sd, prefix + "b1b.", (3, 3, 3))
    b2 = unit3d_bn_relu(unit3d_bn_relu(x, sd, prefix + "b2a.", (1, 1, 1)), sd, prefix + "b2b.", (3, 3, 3))
    b3 = unit3d_bn_relu(maxpool3d_same(x, (3, 3, 3), (1, 1, 1)), sd, prefix + "b3b.", (1, 1, 1))
    return torch.cat([b0, b1, b2, b3], dim=1)


def i3d_features(x, sd, prefix="backbone._model.", keep=("Mixed_4f", "Mixed_5c")):
    """InceptionI3d.extract_features (i3d_backbone.py:335-342), keeping only the consumed endpoints."""
    feats = {}
    for name, kind, a in I3D_ENDPOINTS:
        if kind == "conv":
            x = unit3d_bn_relu(x, sd, f"{prefix}{name}.", a["k"], a["s"])
        elif kind == "pool":
            x = maxpool3d_same(x, a["k"], a["s"])
        else:
            x = inception(x, sd, f"{prefix}{name}.", a["widths"])
        if keep is None or name in keep:
            feats[name] = x
    return feats


def unit1d(x, sd, prefix, stride=1):
    """Unit1D with 'same' padding, bias, no activation: layers.py:204-214."""
    w = sd[prefix + "conv1d.weight"]
    pf, pb = same_pad(x.shape[2], w.shape[2], stride)
    return F.conv1d(F.pad(x, [pf, pb]), w, sd[prefix + "conv1d.bias"], stride=stride)


def gn_relu(x, sd, prefix):
    """nn.GroupNorm(32, C) + ReLU (e.g. BDNet.py:72-73)."""
    return F.relu(F.group_norm(x, GN_GROUPS, sd[prefix + "weight"], sd[prefix + "bias"], GN_EPS))


def head_unit3d_valid(x, sd, prefix):
    """Head-side Unit3D, padding='spatial_valid': temporal same-pad (k_t = 1 -> none), full-extent spatial
    kernel, bias: layers.py:143-175."""
    return F.conv3d(x, sd[prefix + "conv3d.weight"], sd[prefix + "conv3d.bias"])


# ----------------------------------------------------------------------------------------------------------
# BoundaryMaxPooling (boundary_max_pooling_kernel.cu:18-82, :114-145; boundary_pooling_op.py:7-24)
# ----------------------------------------------------------------------------------------------------------
def _bmp_windows(segments: torch.Tensor, tlen: int) -> torch.Tensor:
    """[B,K,4] float -> int64 (l_start, r_start, l_end, r_end), C truncation then clamp (.cu:33-36)."""
    return segments.detach().trunc().to(torch.int64).clamp_(0, tlen - 1)


def _bmp_max_argmax(rows: torch.Tensor, win: torch.Tensor):
    """rows [B,C,L], win [B,K,4] -> (max [B,C,K], argmax [B,C,K]); strict '>' so the first maximum wins and a
    window with r < l degenerates to the single element l (.cu:37-44, :67-78)."""
    B, C, L = rows.shape
    K = win.shape[1]
    half = C // 2
    t = torch.arange(L).view(1, 1, L)
    outs, args = [], []
    for st in range(2):
        l = win[:, :, 2 * st].unsqueeze(-1)
        r = torch.maximum(win[:, :, 2 * st + 1].unsqueeze(-1), l)
        mask = (t >= l) & (t <= r)                                   # [B,K,L]
        part = rows[:, st * half:(st + 1) * half]                   # [B,half,L]
        neg = torch.finfo(rows.dtype).min
        vals = torch.where(mask.unsqueeze(1), part.unsqueeze(2), torch.full((), neg, dtype=rows.dtype))
        m = vals.max(dim=-1).values                                  # [B,half,K]
        first = ((vals == m.unsqueeze(-1)) & mask.unsqueeze(1)).to(torch.uint8).argmax(dim=-1)
        outs.append(m)
        args.append(first)
    return torch.cat(outs, 1), torch.cat(args, 1)


class _BoundaryMaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, segments, compat_tscale_bug):
        out, arg = _bmp_max_argmax(inp, _bmp_windows(segments, inp.shape[2]))
        ctx.save_for_backward(inp, segments, arg)
        ctx.compat = bool(compat_tscale_bug)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        inp, segments, arg = ctx.saved_tensors
        B, C, T = inp.shape
        K = segments.shape[1]
        grad_out = grad_out.contiguous()
        if not ctx.compat or K == T:
            gin = torch.zeros_like(inp).scatter_add_(2, arg, grad_out)
            return gin, None, None
        # reference backward: tscale = grad_output.size(2) = K for clamping AND addressing (.cu:121): the flat
        # input / grad buffers are re-read as rows of K floats.
        if K > T:
            raise RuntimeError("compat backward with K > T reads out of bounds in the reference")
        flat_rows = inp.contiguous().view(-1)[: B * C * K].view(B, C, K)
        _, arg2 = _bmp_max_argmax(flat_rows, _bmp_windows(segments, K))
        gflat = torch.zeros(B * C * T, dtype=inp.dtype)
        gflat[: B * C * K] = torch.zeros(B, C, K, dtype=inp.dtype).scatter_add_(2, arg2, grad_out).view(-1)
        return gflat.view(B, C, T), None, None


def boundary_max_pooling(inp, segments, compat_tscale_bug=True):
    return _BoundaryMaxPool.apply(inp, segments, compat_tscale_bug)


# ----------------------------------------------------------------------------------------------------------
# model layer (AFSD/thumos14/BDNet.py)
# ----------------------------------------------------------------------------------------------------------
def proposal_branch(feat, frame_feat, segments, frame_segments, sd, prefix, compat):
    """ProposalBranch.forward: BDNet.py:105-113."""
    fm_short = gn_relu(unit1d(feat, sd, prefix + "cur_point_conv.0."), sd, prefix + "cur_point_conv.1.")
    lr = gn_relu(unit1d(feat, sd, prefix + "lr_conv.0."), sd, prefix + "lr_conv.1.")
    prop = boundary_max_pooling(lr, segments, compat)
    roi = boundary_max_pooling(frame_feat, frame_segments, compat)
    roi = gn_relu(unit1d(roi, sd, prefix + "roi_conv.0."), sd, prefix + "roi_conv.1.")
    cat = torch.cat([roi, prop, fm_short], dim=1)
    return gn_relu(unit1d(cat, sd, prefix + "proposal_conv.0."), sd, prefix + "proposal_conv.1."), lr


def make_segments(loc, prior, t, frame_num):
    """Segment generation, no grad: BDNet.py:355-384.  loc [B,t,2] (frames), prior [t,1] in (0,1)."""
    with torch.no_grad():
        B = loc.shape[0]
        seg = loc / frame_num * t
        pri = prior.view(1, t, 1).expand(B, t, 1)
        centre = torch.round(pri * t - 0.5)
        plen = seg[:, :, :1] + seg[:, :, 1:]
        inl = torch.clamp(plen / 4.0, min=1.0)
        outl = torch.clamp(plen / 10.0, min=1.0)
        ls = centre - seg[:, :, :1]
        rs = centre + seg[:, :, 1:]
        segments = torch.cat([torch.round(ls - outl), torch.round(ls + inl),
                              torch.round(rs - inl), torch.round(rs + outl)], dim=-1)
        dl = pri * frame_num - loc[:, :, :1]
        dr = pri * frame_num + loc[:, :, 1:]
        plen = dr - dl + 1.0
        inl = torch.clamp(plen / 4.0, min=1.0)
        outl = torch.clamp(plen / 10.0, min=1.0)
        frame_segments = torch.cat([torch.round(dl - outl), torch.round(dl + inl),
                                    torch.round(dr - inl), torch.round(dr + outl)], dim=-1)
    return segments, frame_segments


def level_priors(cfg: OracleConfig) -> list[torch.Tensor]:
    """(c + 0.5)/t per level: BDNet.py:286-293; ANet appends the level index: anet/BDNet.py:262-269."""
    out, t = [], cfg.feat_t
    for i in range(NUM_LEVELS):
        if cfg.variant == "anet":
            out.append(torch.tensor([[(c + 0.5) / t, i] for c in range(t)], dtype=torch.float32))
        else:
            out.append(torch.tensor([[(c + 0.5) / t] for c in range(t)], dtype=torch.float32))
        t //= 2
    return out


def coarse_pyramid(feats, sd, cfg: OracleConfig, compat=True, forced_segments=None, return_segments=False, ssl=False):
    """CoarsePyramid.forward: BDNet.py:295-432 (non-ssl path).  `forced_segments` (list of per-level
    (segments, frame_segments)) lets layer-wise parity tests bypass the discrete rounding hazard."""
    cp = "coarse_pyramid_detection."
    B = feats["Mixed_5c"].shape[0]
    K = cfg.num_classes
    anet = cfg.variant == "anet"
    if anet:       # anet/BDNet.py:281-289: Mixed_5c only
        B = feats["Mixed_5c"].shape[0]
        x = gn_relu(head_unit3d_valid(feats["Mixed_5c"], sd, cp + "pyramids.0.0."), sd, cp + "pyramids.0.1.")
        x = x.squeeze(-1).squeeze(-1)
        levels, first = [x], 1
    else:
        x0 = gn_relu(head_unit3d_valid(feats["Mixed_4f"], sd, cp + "pyramids.0.0."), sd, cp + "pyramids.0.1.")
        x0 = x0.squeeze(-1).squeeze(-1)
        x1 = gn_relu(head_unit3d_valid(feats["Mixed_5c"], sd, cp + "pyramids.1.0."), sd, cp + "pyramids.1.1.")
        x1 = x1.squeeze(-1).squeeze(-1)
        levels, first = [x0 + F.interpolate(x1, x0.shape[2:], mode="nearest"), x1], 2
        x = x1
    for i in range(first, NUM_LEVELS):
        x = gn_relu(unit1d(x, sd, f"{cp}pyramids.{i}.0.", stride=2), sd, f"{cp}pyramids.{i}.1.")
        levels.append(x)

    frame = F.interpolate(levels[0].unsqueeze(-1), [cfg.frame_num, 1]).squeeze(-1)
    for j in (0, 3, 6):
        frame = gn_relu(unit1d(frame, sd, f"{cp}deconv.{j}."), sd, f"{cp}deconv.{j + 1}.")
    start = frame[:, :256].permute(0, 2, 1).contiguous()
    end = frame[:, 256:].permute(0, 2, 1).contiguous()

    priors = level_priors(cfg)
    locs, confs, acts, centers, plocs, pconfs, pacts, segs = [], [], [], [], [], [], [], []
    extra = {}
    for i, feat in enumerate(levels):
        lf, cf = feat, feat
        for j in range(2):
            lf = gn_relu(unit1d(lf, sd, f"{cp}loc_tower.{j}.0."), sd, f"{cp}loc_tower.{j}.1.")
            cf = gn_relu(unit1d(cf, sd, f"{cp}conf_tower.{j}.0."), sd, f"{cp}conf_tower.{j}.1.")
        t = feat.shape[2]
        loc = torch.exp(unit1d(lf, sd, cp + "loc_head.") * sd[f"{cp}loc_heads.{i}.scale"])     # ScaleExp BDNet.py:55-61
        loc = loc.view(B, 2, -1).permute(0, 2, 1).contiguous()
        if anet:
            loc = loc * ANET_FPN_STRIDES[i]                                                    # anet/BDNet.py:307-311
        locs.append(loc)
        confs.append(unit1d(cf, sd, cp + "conf_head.").view(B, K, -1).permute(0, 2, 1).contiguous())
        if cfg.os_head:
            acts.append(unit1d(cf, sd, cp + "actionness_head.").view(B, 1, -1).permute(0, 2, 1).contiguous())
        if forced_segments is not None:
            segments, frame_segments = forced_segments[i]
        else:
            segments, frame_segments = make_segments(loc, priors[i][:, :1], t, cfg.frame_num)
        segs.append((segments, frame_segments))
        lp, lp_ = proposal_branch(lf, frame, segments, frame_segments, sd, cp + "loc_proposal_branch.", compat)
        cpf, cp_ = proposal_branch(cf, frame, segments, frame_segments, sd, cp + "conf_proposal_branch.", compat)
        if i == 0 and ssl:      # BDNet.py:397-400: the frame-level feature and the two level-0 boundary features
            return [frame.clone(), lp_.clone(), cp_.clone()]
        if i == 0:
            nd = lp_.shape[1] // 2
            extra = dict(start_loc_prop=lp_[:, :nd].permute(0, 2, 1).contiguous(),
                         end_loc_prop=lp_[:, nd:].permute(0, 2, 1).contiguous(),
                         start_conf_prop=cp_[:, :nd].permute(0, 2, 1).contiguous(),
                         end_conf_prop=cp_[:, nd:].permute(0, 2, 1).contiguous())
        plocs.append(unit1d(lp, sd, cp + "prop_loc_head.").view(B, 2, -1).permute(0, 2, 1).contiguous())
        pconfs.append(unit1d(cpf, sd, cp + "prop_conf_head.").view(B, K, -1).permute(0, 2, 1).contiguous())
        if cfg.os_head:
            pacts.append(unit1d(cpf, sd, cp + "prop_actionness_head.").view(B, 1, -1).permute(0, 2, 1).contiguous())
        centers.append(unit1d(lp, sd, cp + "center_head.").view(B, 1, -1).permute(0, 2, 1).contiguous())

    out = dict(loc=torch.cat(locs, 1), conf=torch.cat(confs, 1), priors=torch.cat(priors, 0),
               prop_loc=torch.cat(plocs, 1), prop_conf=torch.cat(pconfs, 1), center=torch.cat(centers, 1),
               start=start, end=end, **extra,
               act=torch.cat(acts, 1) if cfg.os_head else None,
               prop_act=torch.cat(pacts, 1) if cfg.os_head else None)
    if return_segments:
        return out, segs
    return out


def dirichlet_uncertainty(logit):
    """DirichletLayer.compute_uncertainty, 'exp' evidence: BDNet.py:544-556."""
    alpha = torch.exp(torch.clamp(logit, -10, 10)) + 1
    return logit.shape[-1] / alpha.sum(-1)


def bdnet_forward(x, sd, cfg: OracleConfig, compat=True, forced_segments=None, return_segments=False):
    """BDNet.forward, non-ssl: BDNet.py:479-535."""
    feats = i3d_features(x, sd)
    res = coarse_pyramid(feats, sd, cfg, compat, forced_segments, return_segments)
    out, segs = res if return_segments else (res, None)
    if cfg.use_edl:
        out["unct"] = dirichlet_uncertainty(out["conf"])
        out["prop_unct"] = dirichlet_uncertainty(out["prop_conf"])
    return (out, segs) if return_segments else out


def bdnet_forward_ssl(x, sd, cfg: OracleConfig, proposals, compat=True):
    """BDNet.forward(ssl=True): BDNet.py:482-503.  proposals: list with ONE tensor [3,2] of (start, end) frames — anchor,
    positive and negative segment of the cut-paste augmentation (thumos_dataset.py:187-226).  Returns the three lists
    (anchor, positive, negative), one entry per feature: frame level, loc branch, conf branch."""
    feats = i3d_features(x, sd)
    top = coarse_pyramid(feats, sd, cfg, compat, ssl=True)
    dec = proposals[0].unsqueeze(0)
    plen = dec[:, :, 1:] - dec[:, :, :1] + 1.0
    inl = torch.clamp(plen / 4.0, min=1.0)
    outl = torch.clamp(plen / 10.0, min=1.0)
    fs = torch.cat([torch.round(dec[:, :, :1] - outl), torch.round(dec[:, :, :1] + inl),
                    torch.round(dec[:, :, 1:] - inl), torch.round(dec[:, :, 1:] + outl)], dim=-1)
    anchor, positive, negative = [], [], []
    for i, scale in enumerate((1, 4, 4)):
        seg = (fs / scale).expand(top[i].shape[0], -1, -1).contiguous()      # the reference indexes sample 0 for all (D11)
        bound = boundary_max_pooling(top[i], seg, compat)
        nd = bound.shape[1] // 2
        anchor.append(bound[:, nd:, 0])
        positive.append(bound[:, :nd, 1])
        negative.append(bound[:, :nd, 2])
    return anchor, positive, negative


def triplet_loss(anchor, positive, negative):
    """Sum of nn.TripletMarginLoss() (margin 1, p 2, eps 1e-6, mean) with weights (1, 0.1, 0.1): train.py:174-184."""
    total = 0.0
    for a, p, n, w in zip(anchor, positive, negative, (1.0, 0.1, 0.1)):
        d_ap = (a - p + 1e-6).norm(dim=1)
        d_an = (a - n + 1e-6).norm(dim=1)
        total = total + w * torch.clamp(d_ap - d_an + 1.0, min=0).mean()
    return total


# ----------------------------------------------------------------------------------------------------------
# loss layer (AFSD/thumos14/multisegment_loss.py, cls_loss.py)
# ----------------------------------------------------------------------------------------------------------
_EPS32 = torch.finfo(torch.float32).eps


def seg_iou(pred, target):
    """1-D IoU on (left, right) offsets: multisegment_loss.py:20-36."""
    inter = torch.min(pred[:, 0], target[:, 0]) + torch.min(pred[:, 1], target[:, 1])
    union = (target[:, 0] + target[:, 1]) + (pred[:, 0] + pred[:, 1]) - inter
    return inter / union.clamp(min=_EPS32), inter, union


def seg_giou_loss(pred, target):
    """1 - GIoU: multisegment_loss.py:40-43."""
    iou, _, union = seg_iou(pred, target)
    hull = torch.max(pred[:, 0], target[:, 0]) + torch.max(pred[:, 1], target[:, 1])
    return 1.0 - (iou - (hull - union) / hull.clamp(min=_EPS32))


@dataclass
class LossState:
    """Mutable state of EvidenceLoss: epoch (train.py:360-362) and the IBM EMA buffer (cls_loss.py:110-115)."""

    epoch: int = 0
    weight_accum: torch.Tensor = field(default_factory=lambda: torch.ones(50))


def match_priors(loc, priors, targets, cfg: OracleConfig):
    """Prior <-> ground-truth matching, no grad: multisegment_loss.py:120-153."""
    B, P = loc.shape[:2]
    clip = float(cfg.clip_length)
    loc_t = torch.zeros(B, P, 2, dtype=loc.dtype)
    conf_t = torch.zeros(B, P, dtype=torch.long)
    prop_loc_t = torch.zeros(B, P, 2, dtype=loc.dtype)
    prop_conf_t = torch.zeros(B, P, dtype=torch.long)
    iou_pred = torch.zeros(P, B, dtype=loc.dtype)
    with torch.no_grad():
        c = priors[:, 0]
        for b in range(B):
            tr, lab = targets[b][:, :2].to(loc.dtype), targets[b][:, 2].long()
            left = (c[:, None] - tr[None, :, 0]) * clip
            right = (tr[None, :, 1] - c[:, None]) * clip
            area = left + right
            big = clip * 2
            area = torch.where((left < 0) | (right < 0), torch.full_like(area, big), area)
            best, idx = area.min(1)
            loc_t[b, :, 0] = (c - tr[idx, 0]) * clip
            loc_t[b, :, 1] = (tr[idx, 1] - c) * clip
            conf = torch.where(best >= big, torch.zeros_like(lab[idx]), lab[idx])
            conf_t[b] = conf
            iou = seg_iou(loc[b], loc_t[b])[0]
            iou_pred[:, b] = iou
            prop_conf_t[b] = torch.where(iou < cfg.piou, torch.zeros_like(conf), conf)
            w = loc[b, :, 0] + loc[b, :, 1]
            prop_loc_t[b] = (loc_t[b] - loc[b]) / (0.5 * w)[:, None]
    return loc_t, conf_t, prop_loc_t, prop_conf_t, iou_pred


def edl_loss(logit, target, state: LossState, cfg: OracleConfig):
    """EvidenceLoss.forward + edl_loss, loss_type 'log', 'exp' evidence, optional IBM re-weighting:
    cls_loss.py:132-168, 212-278 (IBM :257-270).  Mutates state.weight_accum like the reference."""
    K = cfg.num_classes
    y = torch.eye(K, dtype=logit.dtype)[target]
    alpha = torch.exp(torch.clamp(logit, -10, 10)) + 1
    S = alpha.sum(1, keepdim=True)
    per = (y * (torch.log(S) - torch.log(alpha))).sum(1)
    if cfg.with_ibm and state.epoch >= cfg.ibm_start:
        with torch.no_grad():
            feat_norm = logit.abs().sum(1)
            a = alpha.detach()
            u = K / a.sum(-1, keepdim=True)
            gnorm = ((1 / a - u).abs() * y).sum(1)
            ghat = gnorm * feat_norm
            bins = torch.ceil(gnorm * cfg.num_bins).long()
            for i in range(cfg.num_bins):
                sel = bins == i + 1
                if sel.any():
                    state.weight_accum[i] = cfg.momentum * state.weight_accum[i] + (1 - cfg.momentum) * ghat[sel].mean()
            w = state.weight_accum[bins - 1]   # bin 0 -> index -1 (python wrap), cls_loss.py:268
        per = w * per
    return per.sum()


class EdlVariantState:
    """Mutable state of the ablation branches of EvidenceLoss (cls_loss.py:101-116): epoch, GHM `acc_sum` (python
    floats), IBM `weight_accum`."""

    def __init__(self, num_bins: int, epoch: int = 0):
        self.epoch = epoch
        self.acc_sum = [0.0] * num_bins
        self.weight_accum = torch.ones(num_bins)


def evidence_loss_variant(logit, target, state: EdlVariantState, num_cls: int, cfg: dict, size_average=False):
    """EvidenceLoss.forward with EVERY branch of the reference (cls_loss.py:132-285): loss_type log / digamma / mse,
    evidence exp / relu / softplus, soft labels, and the re-weighting branches in the reference's precedence order
    focal (:221-227) > GHM (:228-249) > IB (:250-256) > IBM (:257-270) > plain (:272).  The OpenTAL configuration is the
    IBM branch (`edl_loss` above); the others are the paper's ablations."""
    K = num_cls
    y = torch.eye(K, dtype=logit.dtype)[target.view(-1)]
    soft = cfg.get("soft_label", 0.0)
    ones = y == 1
    y = torch.where(ones, torch.full_like(y, 1 - soft), torch.full_like(y, soft / (K - 1)))   # :149-150
    ev = cfg["evidence"]
    e = F.relu(logit) if ev == "relu" else torch.exp(torch.clamp(logit, -10, 10)) if ev == "exp" else F.softplus(logit)
    alpha = e + 1
    S = alpha.sum(1, keepdim=True)
    red = (lambda v: v.mean()) if size_average else (lambda v: v.sum())
    if cfg["loss_type"] == "mse":                                       # :193-209, 280-284: err + var, both 'loss' keys
        err = ((y - alpha / S) ** 2).sum(1, keepdim=True)
        var = (alpha * (S - alpha) / (S * S * (S + 1))).sum(1, keepdim=True)
        return red(err) + red(var)
    func = torch.log if cfg["loss_type"] == "log" else torch.digamma
    base = y * (func(S) - func(alpha))
    num_bins = cfg.get("num_bins", 50)
    momentum = cfg.get("momentum", 0.99 if cfg.get("with_ibm") else 0.0)
    if cfg.get("with_focal", False):
        a_cls = torch.ones(K) * (1 - cfg["alpha"])
        a_cls[0] = cfg["alpha"]
        pred = (alpha / S).max(1).values                                # NOT detached in the reference (:224-226)
        w = a_cls[target.view(-1)] * (1.0 - pred) ** cfg["gamma"]
        per = (base * w.unsqueeze(-1)).sum(1)
    elif cfg.get("with_ghm", False) and state.epoch >= cfg.get("ghm_start", 0):
        a = alpha.detach()
        u = K / a.sum(-1, keepdim=True)
        g = (1 / a - u).abs() * y
        edges = [float(x) / num_bins for x in range(num_bins + 1)]
        edges[-1] += 1e-6
        weights = torch.zeros_like(alpha)
        n = 0
        for i in range(num_bins):
            inds = (g >= edges[i]) & (g < edges[i + 1])
            cnt = int(inds.sum())
            if cnt > 0:
                if momentum > 0:
                    state.acc_sum[i] = momentum * state.acc_sum[i] + (1 - momentum) * cnt
                    weights[inds] = 1.0 / state.acc_sum[i]
                else:
                    weights[inds] = 1.0 / cnt
                n += 1
        if n > 0:
            weights = weights / n
        per = (base * weights).sum(1)
    elif cfg.get("with_ibloss", False) and state.epoch >= cfg.get("ib_start", 10):
        a = alpha.detach()
        u = K / a.sum(-1, keepdim=True)
        g = ((1 / a - u).abs() * y).sum(1)
        per = base.sum(1) / (g * logit.detach().abs().sum(1))
    elif cfg.get("with_ibm", False) and state.epoch >= cfg.get("ibm_start", 0):
        a = alpha.detach()
        u = K / a.sum(-1, keepdim=True)
        g = ((1 / a - u).abs() * y).sum(1)
        ghat = g * logit.detach().abs().sum(1)
        bins = torch.ceil(g * num_bins).long()
        for i in range(num_bins):
            sel = bins == i + 1
            if sel.any():
                state.weight_accum[i] = momentum * state.weight_accum[i] + (1 - momentum) * ghat[sel].mean()
        per = state.weight_accum[bins - 1] * base.sum(1)
    else:
        per = base.sum(1)
    return red(per)


EDL_VARIANTS = {
    # name: (edl_config, epochs the two calls run at)            -- branches of cls_loss.py:212-278
    "plain_log_exp": (dict(loss_type="log", evidence="exp"), (1, 2)),
    "digamma_softplus": (dict(loss_type="digamma", evidence="softplus"), (1, 2)),
    "mse_relu": (dict(loss_type="mse", evidence="relu"), (1, 2)),
    "soft_label": (dict(loss_type="log", evidence="exp", soft_label=0.1), (1, 2)),
    "focal": (dict(loss_type="log", evidence="exp", with_focal=True, alpha=0.25, gamma=2.0), (1, 2)),
    "ghm_momentum": (dict(loss_type="log", evidence="exp", with_ghm=True, num_bins=30, momentum=0.75, ghm_start=2), (1, 2, 3)),
    "ghm_plain": (dict(loss_type="log", evidence="exp", with_ghm=True, num_bins=10, momentum=0.0), (1, 2)),
    "ibloss": (dict(loss_type="log", evidence="exp", with_ibloss=True, ib_start=2), (1, 2)),
    "ibm": (dict(loss_type="log", evidence="exp", with_ibm=True, ibm_start=0, momentum=0.9, num_bins=50), (1, 2)),
    "ibm_digamma_mean": (dict(loss_type="digamma", evidence="exp", with_ibm=True, ibm_start=0), (1, 2)),
}


def edl_inputs(name: str, call: int, K: int = 15, M: int = 96):
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) + call)
    return 2.5 * torch.randn(M, K, generator=g), torch.randint(0, K, (M,), generator=g)


def fake_head_outputs(B: int, seed: int, K: int = 15, P: int = 126, loc_scale: float = 30.0) -> dict:
    """Seeded stand-in for the head outputs (loss-only fixtures)."""
    g = torch.Generator().manual_seed(seed)
    return dict(loc=(torch.rand(B, P, 2, generator=g) * loc_scale + 1), conf=2 * torch.randn(B, P, K, generator=g),
                prop_loc=0.3 * torch.randn(B, P, 2, generator=g), prop_conf=2 * torch.randn(B, P, K, generator=g),
                center=torch.randn(B, P, 1, generator=g), act=torch.randn(B, P, 1, generator=g),
                prop_act=torch.randn(B, P, 1, generator=g))


def iou_calibration(logit, ious, cfg: OracleConfig):
    """EvidenceLoss.iou_calib (mean): cls_loss.py:120-129."""
    ious = torch.where(ious < 0, torch.full_like(ious, 1e-3), ious)
    u = cfg.num_classes / (torch.exp(torch.clamp(logit, -10, 10)) + 1).sum(-1)
    return (-ious * torch.log(1 - u) - (1 - ious) * torch.log(u)).mean()


def actionness_loss(logit, label, cfg: OracleConfig):
    """ActionnessLoss.forward: cls_loss.py:299-339.  Returns (loss, #pos + #kept_neg)."""
    pred, label = logit.view(-1), label.view(-1)
    pos, neg = pred[label > 0], pred[label == 0]
    npos, nneg = pos.numel(), neg.numel()
    top_m = min(npos, nneg) - 1
    if top_m > 0:
        kept = neg[neg.sort().indices[:top_m]]
        use = torch.cat([pos, kept])
        tgt = torch.cat([torch.ones_like(pos), torch.zeros_like(kept)])
        nneg = top_m
    else:
        use, tgt = pred, label
    loss = F.binary_cross_entropy_with_logits(use, tgt, reduction="sum")
    if top_m > 0:
        rank = torch.clamp(cfg.act_margin - neg.max() + pos.max().detach(), min=0.0)
        loss = loss + cfg.act_weight * rank
    return loss, npos + nneg


def multisegment_loss(out, targets, state: LossState, cfg: OracleConfig):
    """MultiSegmentLoss.forward (cls_loss_type='edl', os_head): multisegment_loss.py:92-259.
    Returns (loss_l, loss_c, loss_prop_l, loss_prop_c, loss_ct, loss_act, loss_prop_act)."""
    loc, conf, ploc, pconf, center, priors = (out[k] for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "priors"))
    K = cfg.num_classes
    loc_t, conf_t, prop_loc_t, prop_conf_t, iou_pred = match_priors(loc, priors, targets, cfg)
    pos, ppos = conf_t > 0, prop_conf_t > 0
    zero = loc.sum() * 0

    loss_l = seg_giou_loss(loc[pos], loc_t[pos]).sum() if pos.any() else zero
    loss_prop_l = (ploc[ppos] - prop_loc_t[ppos]).abs().sum() if ppos.any() else zero
    if pos.any():
        pre = loc[pos]
        cur = 0.5 * (pre[:, 0] + pre[:, 1]).unsqueeze(-1) * ploc[pos] + pre
        q = seg_iou(cur, loc_t[pos])[0].clamp(min=0)
        loss_ct = F.binary_cross_entropy_with_logits(center[pos].view(-1), q, reduction="sum")
    else:
        loss_ct = zero

    flat_pos = pos.view(-1)
    loss_c = edl_loss(conf.view(-1, K)[flat_pos], conf_t.view(-1)[flat_pos] - 1, state, cfg) if pos.any() else torch.tensor(0.0)
    loss_act, AN = actionness_loss(out["act"].view(-1, 1), flat_pos.float(), cfg)
    flat_ppos = ppos.view(-1)
    loss_prop_c = edl_loss(pconf.view(-1, K)[flat_ppos], prop_conf_t.view(-1)[flat_ppos] - 1, state, cfg) if ppos.any() else torch.tensor(0.0)
    # NB the reference pairs logits flattened [B*P] with iou_pred flattened from [P,B] (multisegment_loss.py:116,
    # 146, 236): identical for B == 1, a (reproduced) mis-pairing for B > 1.
    loss_iouc = iou_calibration(pconf.view(-1, K), iou_pred.reshape(-1), cfg) if cfg.iou_aware else 0.0
    loss_prop_act, PAN = actionness_loss(out["prop_act"].view(-1, 1), flat_ppos.float(), cfg)

    N = max(int(pos.sum()), 1)
    PN = max(int(ppos.sum()), 1)
    return (loss_l / N, loss_c / N, loss_prop_l / PN, loss_prop_c / PN + loss_iouc, loss_ct / N,
            loss_act / AN, loss_prop_act / PAN)


def focal_loss_ori(prob, target, num_class: int, alpha: float = 0.25, gamma: float = 2.0, balance_index: int = 0):
    """FocalLoss_Ori.forward, size_average=False (cls_loss.py:6-78): `prob` are softmax scores [N,K+1]; class
    `balance_index` (background) gets weight alpha, every other class 1 - alpha; eps 1e-6 added to p_t."""
    a = torch.ones(num_class, dtype=prob.dtype) * (1 - alpha)
    a[balance_index] = alpha
    pt = prob.gather(1, target.view(-1, 1)).view(-1) + 1e-6
    return (-1 * torch.pow(1.0 - pt, gamma) * (a[target.view(-1)] * pt.log())).sum()


def multisegment_loss_closed(out, targets, cfg: OracleConfig):
    """MultiSegmentLoss.forward for the closed-set baseline (configs/thumos14.yaml: cls_loss_type='focal', no os_head, K+1
    classes incl. background): multisegment_loss.py:92-259 with the focal branches (:193-195, :217-218) — every prior
    contributes to the classification terms, background priors with label 0.  Returns the 5 losses (act terms are None)."""
    loc, conf, ploc, pconf, center, priors = (out[k] for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "priors"))
    K1 = cfg.num_classes
    loc_t, conf_t, prop_loc_t, prop_conf_t, _ = match_priors(loc, priors, targets, cfg)
    pos, ppos = conf_t > 0, prop_conf_t > 0
    zero = loc.sum() * 0
    loss_l = seg_giou_loss(loc[pos], loc_t[pos]).sum() if pos.any() else zero
    loss_prop_l = (ploc[ppos] - prop_loc_t[ppos]).abs().sum() if ppos.any() else zero
    if pos.any():
        pre = loc[pos]
        cur = 0.5 * (pre[:, 0] + pre[:, 1]).unsqueeze(-1) * ploc[pos] + pre
        q = seg_iou(cur, loc_t[pos])[0].clamp(min=0)
        loss_ct = F.binary_cross_entropy_with_logits(center[pos].view(-1), q, reduction="sum")
    else:
        loss_ct = zero
    loss_c = focal_loss_ori(F.softmax(conf.view(-1, K1), dim=1), conf_t.view(-1), K1)
    loss_prop_c = focal_loss_ori(F.softmax(pconf.view(-1, K1), dim=1), prop_conf_t.view(-1), K1)
    N = max(int(pos.sum()), 1)
    PN = max(int(ppos.sum()), 1)
    return loss_l / N, loss_c / N, loss_prop_l / PN, loss_prop_c / PN, loss_ct / N


def edl_loss_anet(logit, target, state: LossState, cfg: OracleConfig):
    """EvidenceLoss of the ActivityNet flavour: same log loss, stateless exp-form IBM weight
    1 / (||logit||_1 * exp(coeff * grad_norm) + 1e-10): anet/cls_loss.py:116-152, :225-232."""
    K = cfg.num_classes
    y = torch.eye(K, dtype=logit.dtype)[target]
    alpha = torch.exp(torch.clamp(logit, -10, 10)) + 1
    S = alpha.sum(1, keepdim=True)
    per = (y * (torch.log(S) - torch.log(alpha))).sum(1)
    if cfg.with_ibm and state.epoch >= cfg.ibm_start:
        feat_norm = logit.abs().sum(1)            # NOT detached in the ANet flavour (anet/cls_loss.py:136, :229):
        with torch.no_grad():                     # the weight back-propagates through ||logit||_1
            a = alpha.detach()
            u = K / a.sum(-1, keepdim=True)
            gnorm = ((1 / a - u).abs() * y).sum(1)
        w = 1.0 / (feat_norm * torch.exp(cfg.ibm_coeff * gnorm) + 1e-10)
        per = w * per
    return per.sum()


def multisegment_loss_anet(out, targets, state: LossState, cfg: OracleConfig):
    """ActivityNet MultiSegmentLoss.forward (cls_loss_type='edl', os_head): anet/multisegment_loss.py:106-301.
    Per-sample loop; level-range gated matching (:156-166, bounds :69-83); refined positives need
    IoU >= min(piou, best IoU among the positives) (:178-184); smooth-L1 refinement loss (:206); every term is
    normalised per sample and averaged over the batch (:268-297)."""
    loc, conf, ploc, pconf, center, priors = (out[k] for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "priors"))
    K, B, clip = cfg.num_classes, loc.shape[0], float(cfg.clip_length)
    lb = torch.tensor([ANET_BOUNDS[int(l)][0] for l in priors[:, 1]], dtype=loc.dtype)[:, None]
    rb = torch.tensor([ANET_BOUNDS[int(l)][1] for l in priors[:, 1]], dtype=loc.dtype)[:, None]
    sums = [0.0] * 7
    for b in range(B):
        with torch.no_grad():
            tr, lab = targets[b][:, :2].to(loc.dtype), targets[b][:, 2].long()
            c = priors[:, 0]
            left = (c[:, None] - tr[None, :, 0]) * clip
            right = (tr[None, :, 1] - c[:, None]) * clip
            max_dis = torch.max(left, right)
            area = left + right
            big = clip * 2
            bad = (left < 0) | (right < 0) | (max_dis <= lb) | (max_dis > rb)
            area = torch.where(bad, torch.full_like(area, big), area)
            best, idx = area.min(1)
            loc_t = torch.stack([(c - tr[idx, 0]) * clip, (tr[idx, 1] - c) * clip], 1)
            conf_t = torch.where(best >= big, torch.zeros_like(lab[idx]), lab[idx])
            iou = seg_iou(loc[b], loc_t)[0]
            max_iou = iou[conf_t > 0].max() if (conf_t > 0).any() else torch.tensor(2.0)
            thr = torch.minimum(torch.tensor(cfg.piou, dtype=iou.dtype), max_iou.to(iou.dtype))
            prop_conf_t = torch.where(iou < thr, torch.zeros_like(conf_t), conf_t)
            w = loc[b, :, 0] + loc[b, :, 1]
            prop_loc_t = (loc_t - loc[b]) / (0.5 * w)[:, None]
        pos, ppos = conf_t > 0, prop_conf_t > 0
        zero = loc[b].sum() * 0
        loss_l = seg_giou_loss(loc[b][pos], loc_t[pos]).sum() if pos.any() else zero
        loss_prop_l = F.smooth_l1_loss(ploc[b][ppos], prop_loc_t[ppos], reduction="sum") if ppos.any() else zero
        if pos.any():
            pre = loc[b][pos]
            cur = 0.5 * (pre[:, 0] + pre[:, 1]).unsqueeze(-1) * ploc[b][pos] + pre
            q = seg_iou(cur, loc_t[pos])[0].clamp(min=0)
            loss_ct = F.binary_cross_entropy_with_logits(center[b][pos].view(-1), q, reduction="sum")
        else:
            loss_ct = zero
        loss_c = edl_loss_anet(conf[b][pos], conf_t[pos] - 1, state, cfg) if pos.any() else torch.tensor(0.0)
        loss_act, AN = actionness_loss(out["act"][b].view(-1, 1), pos.float(), cfg)
        loss_prop_c = edl_loss_anet(pconf[b][ppos], prop_conf_t[ppos] - 1, state, cfg) if ppos.any() else torch.tensor(0.0)
        loss_iouc = iou_calibration(pconf[b], iou, cfg) if cfg.iou_aware else 0.0
        loss_prop_act, PAN = actionness_loss(out["prop_act"][b].view(-1, 1), ppos.float(), cfg)
        N, PN = max(int(pos.sum()), 1), max(int(ppos.sum()), 1)
        terms = (loss_l / N, loss_c / N, loss_prop_l / PN, loss_prop_c / PN + loss_iouc, loss_ct / N, loss_act / AN,
                 loss_prop_act / PAN)
        sums = [s + t for s, t in zip(sums, terms)]
    return tuple(s / B for s in sums)


def boundary_bce(start, end, scores):
    """calc_bce_loss: tanh -> mean over channels -> BCE(mean): thumos14/train.py:152-161."""
    s = torch.tanh(start).mean(-1)
    e = torch.tanh(end).mean(-1)
    return (F.binary_cross_entropy(s.view(-1), scores[:, 0].contiguous().view(-1)),
            F.binary_cross_entropy(e.view(-1), scores[:, 1].contiguous().view(-1)))


def training_cost(out, targets, scores, state: LossState, cfg: OracleConfig, lw=1.0, cw=10.0, ctw=1.0, actw=1.0):
    """Total cost of one non-ssl training step: thumos14/train.py:186-200, 226-235."""
    l, c, pl, pc, ct, act, pact = multisegment_loss(out, targets, state, cfg)
    ls, le = boundary_bce(out["start"], out["end"], scores)
    sc4 = F.interpolate(scores, scale_factor=1.0 / 4)
    a, b = boundary_bce(out["start_loc_prop"], out["end_loc_prop"], sc4)
    c2, d = boundary_bce(out["start_conf_prop"], out["end_conf_prop"], sc4)
    ls = ls + 0.1 * (a + c2)
    le = le + 0.1 * (b + d)
    cost = lw * l + cw * c + lw * pl + cw * pc + ctw * ct + ls + le + actw * act + actw * pact
    return cost, dict(loss_l=l, loss_c=c, loss_prop_l=pl, loss_prop_c=pc, loss_ct=ct, loss_start=ls, loss_end=le,
                      loss_act=act, loss_prop_act=pact)


# ----------------------------------------------------------------------------------------------------------
# inference post-processing (AFSD/thumos14/test.py:112-200, AFSD/common/segment_utils.py:128-162)
# ----------------------------------------------------------------------------------------------------------
def decode_predictions(out, b, offset, sample_fps, cfg: OracleConfig):
    """decode_predictions for clip `b` of an output dict (test.py:112-140) with the Dirichlet score function
    (DirichletLayer.forward, BDNet.py:558-561: alpha / sum alpha) and the open-set head.
    Returns (segments [P,2] seconds, scores [K,P], uncertainty [P], actionness [P])."""
    loc, ploc, priors = out["loc"][b], out["prop_loc"][b], out["priors"]
    w = loc[:, :1] + loc[:, 1:]
    loc = 0.5 * w * ploc + loc
    seg = torch.cat([priors[:, :1] * cfg.clip_length - loc[:, :1], priors[:, :1] * cfg.clip_length + loc[:, 1:]], dim=-1)
    seg = seg.clamp(min=0, max=cfg.clip_length)
    seg = (seg + offset) / sample_fps
    unct = (dirichlet_uncertainty(out["conf"][b]) + dirichlet_uncertainty(out["prop_conf"][b])) / 2.0
    actionness = (out["act"][b].squeeze(-1).sigmoid() + out["prop_act"][b].squeeze(-1).sigmoid()) / 2.0

    def prob(logit):
        alpha = torch.exp(torch.clamp(logit, -10, 10)) + 1
        return alpha / alpha.sum(-1, keepdim=True)

    conf = (prob(out["conf"][b]) + prob(out["prop_conf"][b])) / 2.0
    conf = conf * out["center"][b].sigmoid() * actionness.unsqueeze(-1)
    return seg, conf.view(-1, cfg.num_classes).transpose(1, 0).contiguous(), unct, actionness


def filter_candidates(seg, score_cls, unct, actionness, conf_thresh):
    """filtering (test.py:143-162), open-set head: rows (start, end, score, uncertainty, actionness) or None."""
    m = (score_cls > conf_thresh) & (actionness > 0.5)
    if not m.any():
        return None
    return torch.cat([seg[m], score_cls[m, None], unct[m, None], actionness[m, None]], -1)


def softnms_v2(segments, sigma=0.5, top_k=1000, score_threshold=0.001):
    """softnms_v2 (segment_utils.py:128-162).  segments [N, >=3] (start, end, score, ...).  Gaussian soft-NMS; NB the loop
    stops when ONE candidate is left undone, so that last candidate is never kept.  Returns (kept rows, count, mask)."""
    seg = segments.clone()
    ts, te, sc = seg[:, 0], seg[:, 1], seg[:, 2]
    done = torch.zeros_like(sc, dtype=torch.bool)
    undone = sc >= score_threshold
    while int(undone.sum()) > 1 and int(done.sum()) < top_k:
        cand = undone.nonzero().view(-1)
        idx = int(cand[sc[undone].argmax()])
        undone[idx] = False
        done[idx] = True
        tt1 = ts[undone].clamp(min=float(ts[idx]))
        tt2 = te[undone].clamp(max=float(te[idx]))
        inter = (tt2 - tt1).clamp(min=0)
        dur = te[undone] - ts[undone]
        width = torch.clamp(te[idx] - ts[idx], min=1e-5)
        iou = inter / (width + dur - inter)
        sc[undone] = sc[undone] * torch.exp(-iou ** 2 / sigma)
        undone[sc < score_threshold] = False
    return seg[done], int(done.sum()), done
