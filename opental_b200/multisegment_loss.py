"""MultiSegmentLoss / EvidenceLoss / ActionnessLoss / FocalLoss with the reference's API, restructured for the GPU.

Reference: AFSD/thumos14/multisegment_loss.py:70-259 and AFSD/thumos14/cls_loss.py (FocalLoss_Ori :6-78, EvidenceLoss
:81-285, ActionnessLoss :288-339).  Same constructor arguments, same 7-tuple, same mutable `cls_loss.epoch` /
`cls_loss.total_epoch` / `cls_loss.weight_accum` state — but a different execution model:

  * the reference loops over the batch in Python, gathers positives with boolean masks (dynamic shapes -> a
    device->host sync per gather), and walks the 50 IBM bins with `.item()` (about 100 syncs per step);
  * here every term is a fixed-shape masked reduction over all B x 126 priors: prior<->GT matching is one batched
    op over zero-padded targets, the IBM per-bin EMA is a one-hot segmented mean, the actionness top-M selection is
    a rank comparison against a device scalar.  Nothing reads a value back to the host, so the whole loss can be
    enqueued behind the head (and captured in a CUDA graph).
The arithmetic per element follows the reference line by line (cited inline); summation order differs, which is
fp32 rounding noise (tests/ check 1e-5 relative against the oracle and the reference-generated golden values).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

_EPS = torch.finfo(torch.float32).eps


def _iou(pred, target):
    """1-D IoU of (left, right) offset pairs, last dim 2 (multisegment_loss.py:20-36)."""
    inter = torch.min(pred[..., 0], target[..., 0]) + torch.min(pred[..., 1], target[..., 1])
    union = (target[..., 0] + target[..., 1]) + (pred[..., 0] + pred[..., 1]) - inter
    return inter / union.clamp(min=_EPS), union


def _giou_loss(pred, target):
    """1 - GIoU (multisegment_loss.py:40-43)."""
    iou, union = _iou(pred, target)
    hull = torch.max(pred[..., 0], target[..., 0]) + torch.max(pred[..., 1], target[..., 1])
    return 1.0 - (iou - (hull - union) / hull.clamp(min=_EPS))


def pad_targets(targets, device=None, slots=None):
    """list of B tensors [N_i,3] (start, end, label; normalised) -> ([B,G,3] zero padded, [B,G] valid mask).
    G = the largest N_i of the batch, or `slots` when given (a fixed geometry for a captured step graph; ValueError when a
    sample has more segments than slots).  An already padded (tensor, mask) pair is returned as is."""
    if isinstance(targets, (tuple, list)) and len(targets) == 2 and torch.is_tensor(targets[0]) and targets[0].dim() == 3:
        return targets
    B = len(targets)
    G = max(1, max(int(t.shape[0]) for t in targets))
    if slots is not None:
        if G > slots:
            raise ValueError(f"a sample has {G} ground-truth segments, the padded geometry has {slots} slots")
        G = int(slots)
    device = device or targets[0].device
    padded = torch.zeros(B, G, 3, dtype=torch.float32)
    valid = torch.zeros(B, G, dtype=torch.bool)
    for b, t in enumerate(targets):
        n = int(t.shape[0])
        if n:
            padded[b, :n] = t.detach().to("cpu", torch.float32)
            valid[b, :n] = True
    return padded.to(device, non_blocking=True), valid.to(device, non_blocking=True)


class FocalLoss_Ori(nn.Module):
    """Focal loss on softmax probabilities with per-class alpha (cls_loss.py:6-78).  `weight` masks samples."""

    def __init__(self, num_class, alpha=None, gamma=2, balance_index=-1, size_average=True):
        super().__init__()
        self.num_class, self.gamma, self.size_average, self.eps = num_class, gamma, size_average, 1e-6
        if alpha is None:
            alpha = [0.25, 0.75]
        if isinstance(alpha, (list, tuple)):
            assert len(alpha) == num_class
            a = torch.tensor(list(alpha), dtype=torch.float32)
        elif isinstance(alpha, (float, int)):
            assert 0 < alpha < 1.0 and balance_index > -1
            a = torch.ones(num_class) * (1 - alpha)
            a[balance_index] = alpha
        else:
            a = alpha
        self.register_buffer("alpha", a, persistent=False)
        # host copy of the (background, other classes) weights when alpha has that two-value form (the fused kernel's form)
        self.alpha0 = float(alpha) if isinstance(alpha, (float, int)) and balance_index == 0 else None
        self.epoch, self.total_epoch = 0, 25

    def forward(self, prob, target, weight=None):
        target = target.view(-1, 1)
        pt = prob.gather(1, target).view(-1) + self.eps
        alpha = self.alpha.to(prob.device).gather(0, target.view(-1))
        loss = -1 * torch.pow(1.0 - pt, self.gamma) * (alpha * pt.log())
        if weight is not None:
            loss = torch.where(weight, loss, torch.zeros_like(loss))
        return loss.mean() if self.size_average else loss.sum()


class EvidenceLoss(nn.Module):
    """EDL classification loss with every branch of the reference (cls_loss.py:81-285): loss_type log / digamma / mse,
    evidence exp / relu / softplus, soft labels, IoU-aware calibration, and the re-weighting branches in the
    reference's precedence order focal > GHM > IB > IBM > plain (cls_loss.py:221-272).  All branches are written in
    the masked, synchronisation-free form (`weight` marks the real rows; the reference gathers them with a boolean
    mask and loops over bins with `.item()`), so they can be captured into the training graph.  The OpenTAL
    configuration (log / exp / IBM) additionally has the fused CUDA kernel (MultiSegmentLoss._fused_ok)."""

    def __init__(self, num_cls, cfg, size_average=False):
        super().__init__()
        self.num_cls = num_cls
        self.loss_type = cfg["loss_type"]
        self.evidence = cfg["evidence"]
        if self.loss_type not in ("log", "digamma", "mse"):
            raise NotImplementedError(f"loss_type {self.loss_type}")          # cls_loss.py:177-178
        self.with_focal = cfg.get("with_focal", False)
        self.soft_label = cfg.get("soft_label", 0.0)
        self.iou_aware = cfg.get("iou_aware", False)
        self.with_ghm = cfg.get("with_ghm", False)
        self.with_ibloss = cfg.get("with_ibloss", False)
        self.with_ibm = cfg.get("with_ibm", False)
        if self.with_focal:
            alpha = torch.ones(num_cls) * (1 - cfg["alpha"])
            alpha[0] = cfg["alpha"]
            self.register_buffer("alpha", alpha, persistent=False)
            self.alpha0 = float(cfg["alpha"])          # host copy for the fused kernel
            self.gamma = cfg["gamma"]
        if self.with_ghm:
            self.num_bins = cfg["num_bins"]
            self.momentum = cfg["momentum"]
            self.ghm_start = cfg.get("ghm_start", 0)
            edges = [float(x) / self.num_bins for x in range(self.num_bins + 1)]
            edges[-1] += 1e-6
            self.register_buffer("edges", torch.tensor(edges, dtype=torch.float32), persistent=False)
            # the reference keeps python floats (double precision): fp64 buffer
            self.register_buffer("acc_sum", torch.zeros(self.num_bins, dtype=torch.float64))
        if self.with_ibloss:
            self.ib_start = cfg.get("ib_start", 10)
        if self.with_ibm:
            self.ibm_start = cfg.get("ibm_start", 0)
            self.num_bins = cfg.get("num_bins", 50)
            self.momentum = cfg.get("momentum", 0.99)
            # explicit buffer (the reference keeps it outside the state_dict); persistent so it is saved
            self.register_buffer("weight_accum", torch.ones(self.num_bins))
        self.epoch, self.total_epoch = 0, 25
        self.size_average = size_average

    def evidence_func(self, logit):
        if self.evidence == "relu":
            return F.relu(logit)
        if self.evidence == "exp":
            return torch.exp(torch.clamp(logit, -10, 10))
        return F.softplus(logit)

    def iou_calib(self, logits, ious, mean=False):
        """cls_loss.py:120-129 (the reference patches `ious` in place; a functional where() is equivalent)."""
        ious = torch.where(ious < 0, torch.full_like(ious, 1e-3), ious)
        unc = self.num_cls / (self.evidence_func(logits) + 1).sum(dim=-1)
        reg = -ious * torch.log(1 - unc) - (1 - ious) * torch.log(unc)
        return reg.mean() if mean else reg.sum()

    def _reduce(self, per, weight):
        per = torch.where(weight, per, torch.zeros_like(per))
        if self.size_average:
            return per.sum() / weight.sum().clamp(min=1)
        return per.sum()

    def forward(self, logit, target, weight=None):
        """logit [M,K]; target [M] in 0..K-1; weight [M] bool = which rows are real samples (masked formulation of the
        reference's boolean gather).  Returns the summed (or mean) loss."""
        target = target.view(-1)
        M, K = logit.shape[0], self.num_cls
        if weight is None:
            weight = torch.ones(M, dtype=torch.bool, device=logit.device)
        wf = weight.to(logit.dtype)
        y = (target.unsqueeze(-1) == torch.arange(K, device=logit.device)).to(logit.dtype)   # sync-free one-hot
        if self.soft_label:
            y = torch.where(y == 1, torch.full_like(y, 1 - self.soft_label), torch.full_like(y, self.soft_label / (K - 1)))
        alpha = self.evidence_func(logit) + 1
        S = alpha.sum(dim=1, keepdim=True)
        if self.loss_type == "mse":                      # cls_loss.py:193-209: both terms carry 'loss' in their key
            err = ((y - alpha / S) ** 2).sum(dim=1)
            var = (alpha * (S - alpha) / (S * S * (S + 1))).sum(dim=1)
            return self._reduce(err, weight) + self._reduce(var, weight)
        func = torch.log if self.loss_type == "log" else torch.digamma
        base = y * (func(S) - func(alpha))
        if self.with_focal:                              # cls_loss.py:221-227 (the modulating factor is NOT detached)
            pred = (alpha / S).max(dim=1).values
            w = self.alpha.to(logit.device)[target.clamp(0, K - 1)] * torch.pow(1.0 - pred, self.gamma)
            return self._reduce((base * w.unsqueeze(-1)).sum(dim=1), weight)
        if self.with_ghm and self.epoch >= self.ghm_start:          # cls_loss.py:228-249
            with torch.no_grad():
                a = alpha.detach()
                unc = K / a.sum(dim=-1, keepdim=True)
                g = (1 / a - unc).abs() * y                                        # [M,K] gradient length
                edges = self.edges.to(logit.device)
                inb = (g.unsqueeze(-1) >= edges[:-1]) & (g.unsqueeze(-1) < edges[1:]) & weight.view(M, 1, 1)   # [M,K,bins]
                cnt = inb.sum(dim=(0, 1)).to(torch.float64)
                if self.momentum > 0:
                    acc = torch.where(cnt > 0, self.momentum * self.acc_sum + (1 - self.momentum) * cnt, self.acc_sum)
                    self.acc_sum.copy_(acc)
                else:
                    acc = cnt
                per_bin = torch.where(cnt > 0, 1.0 / acc.clamp(min=1e-300), torch.zeros_like(acc)).to(logit.dtype)
                n = (cnt > 0).sum()
                w = (inb.to(logit.dtype) * per_bin).sum(dim=-1)                     # every element is in at most one bin
                w = torch.where(n > 0, w / n.clamp(min=1).to(logit.dtype), w)
            return self._reduce((base * w).sum(dim=1), weight)
        per = base.sum(dim=1)
        if self.with_ibloss and self.epoch >= self.ib_start:        # cls_loss.py:250-256
            with torch.no_grad():
                a = alpha.detach()
                unc = K / a.sum(dim=-1, keepdim=True)
                grad_norm = ((1 / a - unc).abs() * y).sum(dim=1)
                # padding rows carry target -1 -> grad_norm 0 -> 1/0: keep inf out of the graph (inf * 0 = NaN in backward)
                w = torch.where(weight, 1 / (grad_norm * logit.abs().sum(1)), torch.zeros_like(grad_norm))
            per = w * per
        elif self.with_ibm and self.epoch >= self.ibm_start:
            with torch.no_grad():       # cls_loss.py:257-270
                feat_norm = logit.abs().sum(1)
                a = alpha.detach()
                unc = K / a.sum(dim=-1, keepdim=True)
                grad_norm = ((1 / a - unc).abs() * y).sum(dim=1)
                grad_hat = grad_norm * feat_norm
                bins = torch.ceil(grad_norm * self.num_bins).long()                 # 1..num_bins (0 if grad_norm == 0)
                onehot = (bins.unsqueeze(-1) == torch.arange(1, self.num_bins + 1, device=logit.device)).to(logit.dtype)
                onehot = onehot * wf.unsqueeze(1)
                cnt = onehot.sum(0)
                mean = (onehot * grad_hat.unsqueeze(1)).sum(0) / cnt.clamp(min=1)
                if self.weight_accum.device != logit.device:
                    self.weight_accum = self.weight_accum.to(logit.device)
                acc = torch.where(cnt > 0, self.momentum * self.weight_accum + (1 - self.momentum) * mean, self.weight_accum)
                self.weight_accum.copy_(acc)     # in place: the buffer keeps its address (CUDA-graph replays update it)
                w = acc[(bins - 1) % self.num_bins]                                 # bin 0 -> index -1 (python wrap)
            per = w * per
        return self._reduce(per, weight)


class ActionnessLoss(nn.Module):
    """Positive-unlabeled actionness loss (cls_loss.py:288-339): BCE over the positives and the top-M lowest-scoring
    negatives, M = min(#pos, #neg) - 1, plus a rank term.  Returns (loss, #pos + #kept negatives) as tensors."""

    def __init__(self, size_average=False, cfg=None):
        super().__init__()
        self.size_average = size_average
        self.weight = cfg.get("weight", 0.1) if cfg is not None else 0.1
        self.margin = cfg.get("margin", 1.0) if cfg is not None else 1.0

    def forward(self, logit, target):
        pred = logit.reshape(-1)
        label = target.reshape(-1).to(pred.dtype)
        pos = label > 0
        neg = ~pos
        npos = pos.sum()
        nneg = neg.sum()
        top_m = torch.minimum(npos, nneg) - 1
        use_top = top_m > 0
        # rank of every negative among the negatives, ascending by score (positives pushed to the end)
        key = torch.where(neg, pred.detach(), torch.full_like(pred, float("inf")))
        order = key.argsort()
        rank = torch.empty_like(order)
        rank[order] = torch.arange(order.numel(), device=order.device)
        kept_neg = neg & (rank < top_m)
        sel = torch.where(use_top, pos | kept_neg, torch.ones_like(pos))
        bce = F.binary_cross_entropy_with_logits(pred, label, reduction="none")
        bce = torch.where(sel, bce, torch.zeros_like(bce))
        count = sel.sum()
        loss = bce.sum() / count.clamp(min=1) if self.size_average else bce.sum()
        if self.weight:
            neg_max = torch.where(neg, pred, torch.full_like(pred, -1e30)).max()
            pos_max = torch.where(pos, pred, torch.full_like(pred, -1e30)).max().detach()
            rank_loss = torch.clamp(self.margin - neg_max + pos_max, min=0.0)
            loss = loss + self.weight * torch.where(use_top, rank_loss, torch.zeros_like(rank_loss))
        return loss, count


class _FusedMSLFn(torch.autograd.Function):
    """Autograd boundary of the single-CTA loss kernel (opental_b200/csrc/msl.cu): returns the loss vector [7]."""

    @staticmethod
    def forward(ctx, loc, conf, ploc, pconf, center, act, pact, priors, tgt, valid, weight_accum, cfg):
        from . import ops
        losses, ws = ops.msl_forward(loc, conf, ploc, pconf, center, act, pact, priors, tgt, valid, weight_accum, **cfg)
        ctx.ws, ctx.dims, ctx.with_act = ws, tuple(conf.shape), act is not None
        ctx.mark_non_differentiable(losses)
        return losses[:7].clone(), losses

    @staticmethod
    def backward(ctx, g, _g_stats):
        from . import ops
        B, P, K = ctx.dims
        gl, gc, gpl, gpc, gct, ga, gpa = ops.msl_backward(ctx.ws, g, B, P, K, ctx.with_act)
        return gl, gc, gpl, gpc, gct, ga, gpa, None, None, None, None, None


class MultiSegmentLoss(nn.Module):
    def __init__(self, num_classes, overlap_thresh, negpos_ratio, use_gpu=True, cls_loss_type="focal", edl_config=None,
                 rpl_config=None, os_head=False, act_config=None, size_average=False, *, clip_length=256):
        super().__init__()
        self.num_classes = num_classes
        self.overlap_thresh = overlap_thresh
        self.negpos_ratio = negpos_ratio
        self.use_gpu = use_gpu
        self.cls_loss_type = cls_loss_type
        self.clip_length = clip_length          # the reference reads config['dataset']['training']['clip_length'] (:110)
        if cls_loss_type == "focal":
            self.cls_loss = FocalLoss_Ori(num_classes, balance_index=0, size_average=size_average, alpha=0.25)
        elif cls_loss_type == "edl":
            self.cls_loss = EvidenceLoss(num_classes, edl_config, size_average=size_average)
        else:
            raise NotImplementedError("RPLoss is a competing baseline outside the OpenTAL configs (SURVEY §2 row 6)")
        self.iou_aware = cls_loss_type == "edl" and self.cls_loss.iou_aware
        self.os_head = os_head
        if os_head:
            self.act_loss = ActionnessLoss(size_average=size_average, cfg=act_config)
        self.size_average = size_average
        self.fused = True       # use the single-CTA CUDA kernel when the configuration allows it (see _fused_ok)

    @torch.no_grad()
    def match(self, loc, priors, tgt, valid):
        """Prior <-> ground-truth matching for the whole batch (multisegment_loss.py:120-153)."""
        clip = float(self.clip_length)
        c = priors[:, 0].view(1, -1, 1)                                   # [1,P,1]
        left = (c - tgt[:, None, :, 0]) * clip                           # [B,P,G]
        right = (tgt[:, None, :, 1] - c) * clip
        big = clip * 2
        area = left + right
        area = torch.where((left < 0) | (right < 0), torch.full_like(area, big), area)
        area = torch.where(valid[:, None, :], area, torch.full_like(area, big * 2))   # padding never wins a tie
        best, idx = area.min(dim=2)                                      # first minimum, like the reference
        sel = idx.unsqueeze(-1)
        t_start = tgt[:, :, 0].gather(1, idx)
        t_end = tgt[:, :, 1].gather(1, idx)
        lab = tgt[:, :, 2].long().gather(1, idx)
        cc = priors[:, 0].view(1, -1)
        loc_t = torch.stack([(cc - t_start) * clip, (t_end - cc) * clip], dim=-1)      # [B,P,2]
        conf_t = torch.where(best >= big, torch.zeros_like(lab), lab)
        iou, _ = _iou(loc, loc_t)
        prop_conf_t = torch.where(iou < self.overlap_thresh, torch.zeros_like(conf_t), conf_t)
        w = (loc[..., 0] + loc[..., 1]).unsqueeze(-1)
        prop_loc_t = (loc_t - loc) / (0.5 * w)
        del sel
        return loc_t, conf_t, prop_loc_t, prop_conf_t, iou

    def _fused_ok(self, loc) -> bool:
        """The single-CTA CUDA kernel covers the OpenTAL configuration (edl: log / exp, plain or IBM, os_head) and the closed-set
        baseline (configs/thumos14.yaml: softmax focal loss, no os_head) and the re-weighting ablations of configs/ablations
        (focal-EDL, GHM, IB, no os_head); digamma / mse, relu / softplus evidence, soft labels and size_average run the masked
        torch formulation below."""
        c = self.cls_loss
        if not (self.fused and loc.is_cuda and loc.dtype == torch.float32 and not self.size_average
                and loc.shape[0] * loc.shape[1] <= 4096):
            return False
        if self.cls_loss_type == "focal":
            return not self.os_head and c.alpha0 is not None and c.num_class < 1024
        # edl: every branch a shipped config selects (configs/*.yaml, configs/ablations/*.yaml all use log / exp): plain, IBM, IB,
        # focal-EDL, GHM, with or without os_head.  digamma / mse, relu / softplus and soft labels stay torch formulations.
        return self.cls_loss_type == "edl" and c.loss_type == "log" and c.evidence == "exp" and not c.soft_label

    def forward(self, output_dict, targets, pre_locs=None):
        loc, conf, ploc, pconf, center, priors = (output_dict[k] for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "priors"))
        act, pact = output_dict.get("act"), output_dict.get("prop_act")
        B, P = loc.shape[:2]
        K = self.num_classes
        tgt, valid = pad_targets(targets, loc.device)
        self.last_vec = None
        if self._fused_ok(loc) and self.cls_loss_type == "focal":
            from . import ops
            cfg = dict(clip_length=float(self.clip_length), overlap_thresh=float(self.overlap_thresh), use_ibm=False, momentum=0.0,
                       iou_aware=False, act_weight=0.0, act_margin=0.0, flavour=ops.MSL_FOCAL,
                       focal_alpha=self.cls_loss.alpha0, focal_gamma=float(self.cls_loss.gamma))
            vec, stats = _FusedMSLFn.apply(loc, conf, ploc, pconf, center.reshape(B, P), None, None, priors, tgt, valid, None, cfg)
            self.last_stats, self.last_vec = stats, vec
            return tuple(vec[:5].unbind(0)) + (None, None)
        if self._fused_ok(loc):
            from . import ops
            c = self.cls_loss
            # the reference's precedence: focal > GHM > IB > IBM > plain (cls_loss.py:221-272), each from its start epoch on
            rw, extra = ops.MSL_RW_NONE, {}
            if c.with_focal:
                rw, extra = ops.MSL_RW_FOCAL, dict(edl_focal_alpha=float(c.alpha0), edl_focal_gamma=float(c.gamma))
            elif c.with_ghm and c.epoch >= c.ghm_start:
                if c.acc_sum.device != loc.device:
                    c.acc_sum = c.acc_sum.to(loc.device)
                rw, extra = ops.MSL_RW_GHM, dict(ghm_acc_sum=c.acc_sum, num_bins=int(c.num_bins))
            elif c.with_ibloss and c.epoch >= c.ib_start:
                rw = ops.MSL_RW_IB
            elif c.with_ibm and c.epoch >= c.ibm_start:
                rw = ops.MSL_RW_IBM
            if c.with_ibm and c.weight_accum.device != loc.device:
                c.weight_accum = c.weight_accum.to(loc.device)
            mom = float(c.momentum) if (c.with_ibm or c.with_ghm) else 0.0
            act_w, act_m = (float(self.act_loss.weight), float(self.act_loss.margin)) if self.os_head else (0.0, 0.0)
            cfg = dict(clip_length=float(self.clip_length), overlap_thresh=float(self.overlap_thresh), use_ibm=rw == ops.MSL_RW_IBM,
                       momentum=mom, iou_aware=bool(self.iou_aware), act_weight=act_w, act_margin=act_m, reweight=rw,
                       cls_all=not self.os_head, **extra)
            vec, stats = _FusedMSLFn.apply(loc, conf, ploc, pconf, center.reshape(B, P), act.reshape(B, P) if self.os_head else None,
                                           pact.reshape(B, P) if self.os_head else None, priors, tgt, valid,
                                           c.weight_accum if rw == ops.MSL_RW_IBM else None, cfg)
            self.last_stats = stats        # raw counts #pos, #refined pos, AN, PAN and loss_iouc at [7:12] (device; logging, engine._globalise)
            self.last_vec = vec            # the 7 losses as one tensor (training_cost takes it instead of re-stacking the tuple)
            if not self.os_head:
                return tuple(vec[:5].unbind(0)) + (None, None)
            return tuple(vec.unbind(0))
        loc_t, conf_t, prop_loc_t, prop_conf_t, iou = self.match(loc.detach(), priors, tgt, valid)
        pos, ppos = conf_t > 0, prop_conf_t > 0
        zero = loc.new_zeros(())

        # localisation: GIoU on positives, L1 on refined positives, IoU-quality BCE (:155-189)
        loss_l = torch.where(pos, _giou_loss(loc, loc_t), zero).sum()
        loss_prop_l = torch.where(ppos.unsqueeze(-1), (ploc - prop_loc_t).abs(), zero).sum()
        cur = 0.5 * (loc[..., 0] + loc[..., 1]).unsqueeze(-1) * ploc + loc
        q = _iou(cur, loc_t)[0].clamp(min=0)
        ct = F.binary_cross_entropy_with_logits(center.reshape(B, P), q, reduction="none")
        loss_ct = torch.where(pos, ct, zero).sum()
        if self.size_average:
            loss_l = loss_l / pos.sum().clamp(min=1)
            loss_prop_l = loss_prop_l / (2 * ppos.sum()).clamp(min=1)
            loss_ct = loss_ct / pos.sum().clamp(min=1)

        # classification, coarse and refined (:191-232)
        def cls(logits, labels, mask):
            logits = logits.reshape(-1, K)
            labels, mask = labels.reshape(-1), mask.reshape(-1)
            if self.cls_loss_type == "focal":
                logits = F.softmax(logits, dim=1)
            if self.os_head:
                return self.cls_loss(logits, (labels - 1).clamp(min=0), mask)
            return self.cls_loss(logits, labels, None)

        loss_c = cls(conf, conf_t, pos)
        loss_act = loss_prop_act = None
        if self.os_head:
            loss_act, AN = self.act_loss(act.reshape(-1, 1), pos.reshape(-1, 1).float())
        loss_prop_c = cls(pconf, prop_conf_t, ppos)
        if self.iou_aware:
            # The reference flattens its [P,B] iou buffer against [B*P] logits (:116,:146,:236): identical for B == 1,
            # a mis-pairing for B > 1 that is reproduced here for parity.
            loss_iouc = self.cls_loss.iou_calib(pconf.reshape(-1, K), iou.t().reshape(-1), mean=True)
        if self.os_head:
            loss_prop_act, PAN = self.act_loss(pact.reshape(-1, 1), ppos.reshape(-1, 1).float())

        N = pos.sum().clamp(min=1)
        PN = ppos.sum().clamp(min=1)
        if not self.size_average:
            loss_l, loss_c, loss_ct = loss_l / N, loss_c / N, loss_ct / N
            loss_prop_l, loss_prop_c = loss_prop_l / PN, loss_prop_c / PN
        if self.iou_aware:
            loss_prop_c = loss_prop_c + loss_iouc
        if self.os_head and not self.size_average:
            loss_act = loss_act / AN.clamp(min=1)
            loss_prop_act = loss_prop_act / PAN.clamp(min=1)
        with torch.no_grad():          # the same five numbers the fused kernel reports (engine.globalise_losses reads them)
            f = lambda v: torch.as_tensor(v, dtype=loc.dtype, device=loc.device).reshape(())   # noqa: E731
            self.last_stats = torch.stack([f(pos.sum()), f(ppos.sum()), f(AN) if self.os_head else f(0.0),
                                           f(PAN) if self.os_head else f(0.0), f(loss_iouc) if self.iou_aware else f(0.0)])
        return loss_l, loss_c, loss_prop_l, loss_prop_c, loss_ct, loss_act, loss_prop_act


ANET_BOUNDS = ((0, 30), (15, 60), (30, 120), (60, 240), (96, 768), (256, 768))   # anet/multisegment_loss.py:69


class MultiSegmentLossANet(nn.Module):
    """ActivityNet flavour of MultiSegmentLoss (AFSD/anet/multisegment_loss.py:86-301) with the reference's API:
    `forward([loc, conf, prop_loc, prop_conf, center, priors, act, prop_act], targets)` -> the same 7-tuple.

    Differences from the THUMOS14 loss, all reproduced: matching is gated by a per-level range of the larger of the two
    boundary distances (`bounds`, :69-83, :156-166); refined positives need IoU >= min(piou, best IoU among the sample's
    positives) (:178-184); the refinement loss is smooth-L1 (:206); every term is normalised PER SAMPLE and averaged
    over the batch (:268-297); the IBM weight is the stateless exp form 1 / (||logit||_1 exp(10 g) + 1e-10), which
    back-propagates through ||logit||_1 (anet/cls_loss.py:136, :229); ActionnessLoss(weight=0.1) (:102).
    The reference loops over the batch in Python with boolean gathers; here every term is a fixed-shape masked
    reduction along the prior axis — no host synchronisation."""

    def __init__(self, num_classes, overlap_thresh, negpos_ratio, use_gpu=True, cls_loss_type="focal", edl_config=None,
                 os_head=False, size_average=False, *, clip_length=768, act_weight=0.1, act_margin=1.0, ibm_coeff=10.0):
        super().__init__()
        if cls_loss_type != "edl" or not os_head or size_average:
            raise NotImplementedError("the ActivityNet loss is implemented for the OpenTAL configuration (edl, os_head)")
        self.num_classes, self.overlap_thresh, self.clip_length = num_classes, overlap_thresh, clip_length
        self.cls_loss = EvidenceLoss(num_classes, edl_config)           # holds epoch / ibm_start like the reference
        self.iou_aware = self.cls_loss.iou_aware
        self.act_weight, self.act_margin, self.ibm_coeff = act_weight, act_margin, ibm_coeff
        self.os_head = True
        self._bounds = {}
        self.fused = True       # the single-CTA CUDA kernel (csrc/msl.cu, flavour 1) when the inputs allow it

    def _edl(self, logit, label, mask):
        """Per-sample sum over the masked priors of the (IBM-weighted) log EDL loss.  logit [B,P,K]."""
        K = self.num_classes
        c = self.cls_loss
        y = (label.unsqueeze(-1) == torch.arange(K, device=logit.device)).to(logit.dtype)     # one-hot without the
        alpha = torch.exp(torch.clamp(logit, -10, 10)) + 1                                     # host sync of F.one_hot
        S = alpha.sum(-1, keepdim=True)
        per = (y * (torch.log(S) - torch.log(alpha))).sum(-1)
        if c.with_ibm and c.epoch >= c.ibm_start:
            feat_norm = logit.abs().sum(-1)                               # not detached (anet/cls_loss.py:136)
            with torch.no_grad():
                a = alpha.detach()
                gnorm = ((1 / a - K / a.sum(-1, keepdim=True)).abs() * y).sum(-1)
            per = per / (feat_norm * torch.exp(self.ibm_coeff * gnorm) + 1e-10)
        return torch.where(mask, per, torch.zeros_like(per)).sum(1)

    def _act(self, pred, pos):
        """ActionnessLoss per sample (anet/cls_loss.py:256-296).  pred [B,P], pos [B,P] bool -> normalised loss [B]."""
        neg = ~pos
        npos, nneg = pos.sum(1), neg.sum(1)
        top_m = torch.minimum(npos, nneg) - 1
        use_top = top_m > 0
        key = torch.where(neg, pred.detach(), torch.full_like(pred, float("inf")))
        rank = key.argsort(dim=1).argsort(dim=1)
        sel = torch.where(use_top[:, None], pos | (neg & (rank < top_m[:, None])), torch.ones_like(pos))
        bce = F.binary_cross_entropy_with_logits(pred, pos.to(pred.dtype), reduction="none")
        loss = torch.where(sel, bce, torch.zeros_like(bce)).sum(1)
        neg_max = torch.where(neg, pred, torch.full_like(pred, -1e30)).max(1).values
        pos_max = torch.where(pos, pred, torch.full_like(pred, -1e30)).max(1).values.detach()
        rank_loss = torch.clamp(self.act_margin - neg_max + pos_max, min=0.0)
        loss = loss + self.act_weight * torch.where(use_top, rank_loss, torch.zeros_like(rank_loss))
        return loss / sel.sum(1).clamp(min=1)

    def forward(self, predictions, targets, pre_locs=None):
        loc, conf, ploc, pconf, center, priors, act, pact = predictions
        B, P = loc.shape[:2]
        K = self.num_classes
        tgt, valid = pad_targets(targets, loc.device)
        clip = float(self.clip_length)
        self.last_vec = None
        if self.fused and loc.is_cuda and loc.dtype == torch.float32 and B * P <= 4096 and B <= 64 and priors.dim() == 2:
            from . import ops
            c = self.cls_loss
            cfg = dict(clip_length=clip, overlap_thresh=float(self.overlap_thresh), use_ibm=bool(c.with_ibm and c.epoch >= c.ibm_start),
                       momentum=0.0, iou_aware=bool(self.iou_aware), act_weight=float(self.act_weight), act_margin=float(self.act_margin),
                       flavour=ops.MSL_ANET, ibm_coeff=float(self.ibm_coeff), level_bounds=ANET_BOUNDS)
            vec, stats = _FusedMSLFn.apply(loc, conf, ploc, pconf, center.reshape(B, P), act.reshape(B, P), pact.reshape(B, P),
                                           priors, tgt, valid, None, cfg)
            self.last_stats, self.last_vec = stats, vec
            return tuple(vec.unbind(0))
        with torch.no_grad():       # matching, anet/multisegment_loss.py:142-190
            c = priors[:, 0].view(1, -1, 1)
            lvl = priors[:, 1].long()
            if loc.device not in self._bounds:      # cached: a host->device copy is not allowed inside a CUDA-graph capture
                self._bounds[loc.device] = torch.tensor(ANET_BOUNDS, dtype=loc.dtype, device=loc.device)
            bounds = self._bounds[loc.device]
            lb, rb = bounds[lvl, 0].view(1, -1, 1), bounds[lvl, 1].view(1, -1, 1)
            left = (c - tgt[:, None, :, 0]) * clip
            right = (tgt[:, None, :, 1] - c) * clip
            max_dis = torch.max(left, right)
            big = clip * 2
            area = left + right
            bad = (left < 0) | (right < 0) | (max_dis <= lb) | (max_dis > rb)
            area = torch.where(bad, torch.full_like(area, big), area)
            area = torch.where(valid[:, None, :], area, torch.full_like(area, big * 2))
            best, idx = area.min(dim=2)
            cc = priors[:, 0].view(1, -1)
            loc_t = torch.stack([(cc - tgt[:, :, 0].gather(1, idx)) * clip, (tgt[:, :, 1].gather(1, idx) - cc) * clip], dim=-1)
            lab = tgt[:, :, 2].long().gather(1, idx)
            conf_t = torch.where(best >= big, torch.zeros_like(lab), lab)
            ld = loc.detach()
            iou = _iou(ld, loc_t)[0]
            pos = conf_t > 0
            max_iou = torch.where(pos, iou, torch.full_like(iou, -1e30)).max(1).values
            max_iou = torch.where(pos.any(1), max_iou, torch.full_like(max_iou, 2.0))
            thr = torch.clamp(max_iou, max=self.overlap_thresh)                    # min(piou, max_iou)
            prop_conf_t = torch.where(iou < thr[:, None], torch.zeros_like(conf_t), conf_t)
            ppos = prop_conf_t > 0
            prop_loc_t = (loc_t - ld) / (0.5 * (ld[..., 0] + ld[..., 1]).unsqueeze(-1))
        zero = loc.new_zeros(())
        N = pos.sum(1).clamp(min=1).to(loc.dtype)
        PN = ppos.sum(1).clamp(min=1).to(loc.dtype)
        loss_l = torch.where(pos, _giou_loss(loc, loc_t), zero).sum(1) / N
        sl1 = F.smooth_l1_loss(ploc, prop_loc_t, reduction="none").sum(-1)
        loss_prop_l = torch.where(ppos, sl1, zero).sum(1) / PN
        cur = 0.5 * (loc[..., 0] + loc[..., 1]).unsqueeze(-1) * ploc + loc
        q = _iou(cur, loc_t)[0].clamp(min=0)
        ct = F.binary_cross_entropy_with_logits(center.reshape(B, P), q, reduction="none")
        loss_ct = torch.where(pos, ct, zero).sum(1) / N
        loss_c = self._edl(conf, conf_t - 1, pos) / N
        loss_prop_c = self._edl(pconf, prop_conf_t - 1, ppos) / PN
        if self.iou_aware:          # per-sample mean over ALL priors of the sample (anet/multisegment_loss.py:258-260)
            io = torch.where(iou < 0, torch.full_like(iou, 1e-3), iou)
            unc = K / (torch.exp(torch.clamp(pconf, -10, 10)) + 1).sum(-1)
            loss_prop_c = loss_prop_c + (-io * torch.log(1 - unc) - (1 - io) * torch.log(unc)).mean(1)
        loss_act = self._act(act.reshape(B, P), pos)
        loss_prop_act = self._act(pact.reshape(B, P), ppos)
        return tuple(t.mean() for t in (loss_l, loss_c, loss_prop_l, loss_prop_c, loss_ct, loss_act, loss_prop_act))


def calc_bce_loss(start, end, scores):
    """tanh -> mean over channels -> BCE (AFSD/thumos14/train.py:152-161).  CUDA fp32 inputs take the fused kernel
    (one launch per map forward, one backward); anything else the equivalent torch formulation."""
    if start.is_cuda and start.dtype == torch.float32 and scores.dtype == torch.float32 and scores.stride(2) == 1:
        from . import ops
        return ops.boundary_bce(start, scores[:, 0]), ops.boundary_bce(end, scores[:, 1])
    def bce(p, y):
        # F.binary_cross_entropy's arithmetic (logs clamped at -100) without its target-range check: the ActivityNet maps hold
        # class ids as targets, which torch 1.9 accepted and torch 2.x rejects on the CPU
        return -(y * torch.log(p).clamp(min=-100.0) + (1.0 - y) * torch.log(1.0 - p).clamp(min=-100.0)).mean()
    s = torch.tanh(start).mean(-1)
    e = torch.tanh(end).mean(-1)
    return bce(s.view(-1), scores[:, 0].contiguous().view(-1)), bce(e.view(-1), scores[:, 1].contiguous().view(-1))


_COST_W: dict = {}


def _cost_weights(device, lw, cw, ctw, actw):
    """[13] weights of (loss_l, loss_c, loss_prop_l, loss_prop_c, loss_ct, loss_act, loss_prop_act, the six boundary terms) in the total
    cost, and the [6] weights of loss_start / loss_end (thumos14/train.py:186-200)."""
    key = (device, lw, cw, ctw, actw)
    if key not in _COST_W:
        w = torch.tensor([lw, cw, lw, cw, ctw, actw, actw, 1.0, 1.0, 0.1, 0.1, 0.1, 0.1], dtype=torch.float32)
        sel = torch.tensor([[1.0, 0.0, 0.1, 0.0, 0.1, 0.0], [0.0, 1.0, 0.0, 0.1, 0.0, 0.1]], dtype=torch.float32)
        _COST_W[key] = (w.to(device), sel.to(device))
    return _COST_W[key]


def training_cost(output_dict, losses, scores, *, lw=1.0, cw=10.0, ctw=1.0, actw=1.0, score_scale=4, loss_vec=None):
    """Total cost of one (non-SSL) training step (thumos14/train.py:186-200, 226-235; anet/train.py:168-190 with the
    score maps down-sampled by 8 = score_scale).  On the GPU the six boundary terms are one fused op and the weighted sum is a dot
    product with a cached weight vector (loss_vec: the criterion's [7] loss vector, when it has one, instead of re-stacking the tuple)."""
    loss_l, loss_c, loss_prop_l, loss_prop_c, loss_ct, loss_act, loss_prop_act = losses
    scores = scores[:, -2:]      # the ActivityNet loader's maps are (action, start, end): rows 1, 2 are used (anet/train.py:134-143)
    sc = F.interpolate(scores, scale_factor=1.0 / score_scale)
    start = output_dict["start"]
    if start.is_cuda and start.dtype == torch.float32 and scores.dtype == torch.float32 and scores.stride(2) == 1:
        from . import ops
        maps = [output_dict[k] for k in ("start", "end", "start_loc_prop", "end_loc_prop", "start_conf_prop", "end_conf_prop")]
        tgs = [scores[:, 0], scores[:, 1], sc[:, 0], sc[:, 1], sc[:, 0], sc[:, 1]]
        bce = ops.boundary_bce_multi(maps, tgs)
        if loss_vec is None:
            zero = loss_l.new_zeros(())
            loss_vec = torch.stack([l if l is not None else zero for l in losses])
        w, sel = _cost_weights(start.device, float(lw), float(cw), float(ctw), float(actw))
        cost = torch.dot(torch.cat([loss_vec, bce]), w)
        lse = sel @ bce.detach()
        return cost, lse[0], lse[1]
    ls, le = calc_bce_loss(output_dict["start"], output_dict["end"], scores)
    a, b = calc_bce_loss(output_dict["start_loc_prop"], output_dict["end_loc_prop"], sc)
    c, d = calc_bce_loss(output_dict["start_conf_prop"], output_dict["end_conf_prop"], sc)
    ls = ls + 0.1 * (a + c)
    le = le + 0.1 * (b + d)
    cost = lw * loss_l + cw * loss_c + lw * loss_prop_l + cw * loss_prop_c + ctw * loss_ct + ls + le
    if loss_act is not None:
        cost = cost + actw * loss_act + actw * loss_prop_act
    return cost, ls, le
