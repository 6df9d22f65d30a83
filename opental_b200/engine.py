"""Training step of the OpenTAL hot path on one B200 per process: forward -> MultiSegmentLoss + boundary BCE ->
backward -> (data-parallel gradient all-reduce over NCCL) -> fused Adam.

Replaces the body of `run_one_epoch` / `forward_one_epoch` (AFSD/thumos14/train.py:164-252) for the non-SSL pass:
the cost is `lw*loss_l + cw*loss_c + lw*loss_prop_l + cw*loss_prop_c + ctw*loss_ct + loss_start + loss_end
(+ actw*loss_act + actw*loss_prop_act)` (train.py:226-235), the optimizer is Adam with L2-in-gradient weight decay
over every trainable parameter (train.py:321-323; frozen BN affine parameters are skipped, BDNet.py:47-49).

All trainable parameters live in two flat fp32 buffers (backbone conv weights in kernel layout, head parameters in
registration order); their gradients live in two matching flat buffers that the parameters' .grad alias.  A step is
therefore: zero two buffers, forward/backward, at most two all-reduces, two fused Adam launches.  The head's
all-reduce is launched when the backbone's backward starts, so the 130 MB of head gradients cross NVLink while
the tensor cores work through the backbone; the backbone's 49 MB follow at the end.  No value is read back to the
host inside a step; losses are returned as device tensors.

Data-parallel semantics: clips shard across ranks, gradients are summed by NCCL and divided by the
world size inside the Adam kernel.  Loss normalisers (N, PN, AN), the actionness top-M selection and the IBM EMA are
per-rank, which differs from "the reference at batch 8*W on one GPU" exactly as documented in DESIGN.md.
"""
from __future__ import annotations

import contextlib
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .bdnet import BDNet
from .multisegment_loss import MultiSegmentLoss, pad_targets, training_cost


class FlatParams:
    """Re-homes a list of parameters into one flat fp32 buffer (+ a flat gradient buffer their .grad alias)."""

    def __init__(self, params: list[nn.Parameter]):
        self.params = params
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4            # 16-byte alignment of every block
        self.offsets, self.total = offs, total
        self.w = torch.zeros(total, dtype=torch.float32, device=dev)
        self.g = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, o in zip(params, offs):
            view = self.w[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.g[o:o + p.numel()].view(p.shape)

    def rebind(self) -> None:
        """Undo a zero_grad(set_to_none=True) or a foreign .grad assignment."""
        for p, o in zip(self.params, self.offsets):
            want = self.g.data_ptr() + 4 * o
            if p.grad is None or p.grad.data_ptr() != want:
                view = self.g[o:o + p.numel()].view(p.shape)
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view


class GradReducer:
    """Data-parallel exchange step: sum the flat gradient buffers over the ranks of `process_group` (NCCL over NVLink on
    the GPUs; the same host logic runs over gloo in the CPU tests).  The 1/world scaling is NOT applied here — it is
    folded into the optimizer kernel (`grad_scale`)."""

    def __init__(self, buffers: list[torch.Tensor], process_group=None):
        self.buffers = buffers
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self._work: dict = {}             # bucket index -> the handle of its all-reduce in flight

    @property
    def grad_scale(self) -> float:
        return 1.0 / self.world

    def launch(self, which=None) -> None:
        """Start the all-reduce of buffers[i] for i in `which` (default: all) without waiting."""
        if self.world == 1:
            return
        for i in (range(len(self.buffers)) if which is None else which):
            self._work[i] = dist.all_reduce(self.buffers[i], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def wait(self, which=None) -> None:
        """Make the CURRENT stream wait for the all-reduces of the buckets in `which` (default: all in flight)."""
        for i in (sorted(self._work) if which is None else which):
            w = self._work.pop(i, None)
            if w is not None:
                w.wait()


def shutdown_distributed(trainers=(), grace_s: float = 12.0) -> None:
    """destroy_process_group() that cannot hang the launcher: a step graph that captured NCCL collectives holds references into
    its communicator, so the trainers' graphs are released first; if the teardown stalls anyway a watchdog ends the process
    (exit code 0: the work is done by the time this is called)."""
    import gc
    import threading
    if not (dist.is_available() and dist.is_initialized()):
        return
    timer = threading.Timer(grace_s, lambda: os._exit(0))
    timer.daemon = True
    timer.start()
    for tr in trainers:
        tr._graph = tr._graph_out = tr._static = None
        tr._graph_cache.clear()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dist.destroy_process_group()
    timer.cancel()


def shard_indices(num_items: int, rank: int, world: int) -> range:
    """Clip indices of `rank`: contiguous, disjoint, equal-sized shards (the remainder is dropped like drop_last)."""
    per = num_items // world
    return range(rank * per, (rank + 1) * per)


def global_normaliser_factors(counts: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """f = max(n_rank, 1) * world / max(sum over ranks of n_rank, 1), elementwise, for the raw counts (#positives, #refined
    positives, #actionness samples ...) a rank's loss terms were normalised by.  `term_rank * f` summed over ranks and divided
    by world — which is what the gradient all-reduce + the 1/world factor of the Adam kernel do — is the term normalised by the
    BATCH-GLOBAL count, i.e. the reference's value at batch = world x per-rank batch (multisegment_loss.py:243-254).  One
    all-reduce of a handful of floats."""
    total = counts.clone()
    if world > 1:
        dist.all_reduce(total, group=group)
    return counts.clamp(min=1) * float(world) / total.clamp(min=1)


def globalise_losses(losses, stats: torch.Tensor, factors: torch.Tensor, iouc_live: torch.Tensor | None = None):
    """Re-weight the 7 THUMOS14 loss terms of one rank from per-rank to batch-global normalisers.
    stats = MultiSegmentLoss.last_stats (#pos, #refined pos, AN, PAN, loss_iouc), factors = global_normaliser_factors(stats[:4]).
    loss_prop_c = A / PN + iouc mixes a count-normalised part with a plain mean over all priors (the IoU calibration, a mean of
    per-rank means is already global): only A / PN is re-weighted.  With `iouc_live` (the calibration term recomputed with
    autograd) the gradient is exact as well; without it the calibration term's gradient carries the factor f_PN (~1)."""
    l, c, pl, pc, ct, act, pact = losses
    fN, fPN, fAN, fPAN = factors.unbind(0)
    iouc = stats[4].detach()
    pc2 = pc * fPN + (1.0 - fPN) * (iouc_live if iouc_live is not None else iouc)
    return (l * fN, c * fN, pl * fPN, pc2, ct * fN, act * fAN if act is not None else None,
            pact * fPAN if pact is not None else None)


def live_iou_calibration(crit, out, targets):
    """The IoU-calibration term of loss_prop_c (multisegment_loss.py:233-238, cls_loss.py:120-129) recomputed with autograd —
    the fused loss kernel only returns its value — so that globalise_losses can keep its gradient at weight 1.  None when the
    loss has no such term."""
    if not getattr(crit, "iou_aware", False):
        return None
    tgt, valid = pad_targets(targets, out["loc"].device)
    iou = crit.match(out["loc"].detach(), out["priors"], tgt, valid)[4]
    K = out["prop_conf"].shape[-1]
    return crit.cls_loss.iou_calib(out["prop_conf"].reshape(-1, K), iou.t().reshape(-1), mean=True)


class Trainer:
    def __init__(self, net: BDNet, criterion: MultiSegmentLoss, *, lr=1e-5, weight_decay=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 lw=1.0, cw=10.0, ctw=1.0, actw=1.0, ssl_weight=0.001, process_group=None, backbone_lr_scale=1.0,
                 global_normalisers=False):
        """backbone_lr_scale: learning rate of the backbone group relative to `lr` — anet/train.py:303-310 trains the backbone
        at 0.1 x the head's rate; thumos14/train.py:321-323 uses one rate (1.0).
        global_normalisers (THUMOS14 loss, world > 1; off by default): normalise the loss terms by the batch-GLOBAL counts of
        positives instead of each rank's own — one all-reduce of 4 floats between the loss and the backward.
        The actionness top-M selection and the IBM EMA stay per rank (§8e (2), (3))."""
        self.net, self.criterion = net, criterion
        self.backbone_lr_scale = float(backbone_lr_scale)
        self.global_normalisers = bool(global_normalisers)
        self.lw, self.cw, self.ctw, self.actw, self.ssl_weight = lw, cw, ctw, actw, ssl_weight
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        p0 = next(net.coarse_pyramid_detection.parameters())
        dev = p0.device
        ops._require_cuda(p0)          # RuntimeError on anything but a CUDA model: the training step has no CPU fallback
        self.device = dev
        bb = net.backbone
        self.bb_w, self.bb_g = bb.flat_parameters(dev)
        bb._bind_grads()
        skip = {id(p) for p in bb.parameters()}
        # head conv weights live in the head's packed store (kernel layout), everything else (biases, GroupNorm,
        # ScaleExp) in a plain flat buffer
        self.hc = net.coarse_pyramid_detection.conv_store
        self.groups = [(self.bb_w, self.bb_g)]
        if self.hc is not None:
            self.hc.ensure(dev)
            self.hc.bind_grads()
            skip |= {id(r.weight) for r in self.hc.recs}
            self.groups.append((self.hc.flat_w, self.hc.flat_g))
        head = [p for p in net.parameters() if p.requires_grad and id(p) not in skip]
        self.head = FlatParams(head)
        self.groups.append((self.head.w, self.head.g))
        self.state = [dict(m=torch.zeros_like(w), v=torch.zeros_like(w)) for w, _ in self.groups]
        # Exchange buckets, in the order their gradients complete during the backward: the head's two buffers (complete when the
        # backbone backward starts), the backbone's Mixed_4b..5c (89 % of its parameters; complete after Mixed_4b's backward), and
        # the rest of the backbone at the end.  Each all-reduce is launched as soon as its bucket is complete and overlaps the
        # remaining backward; the launches, the wait and the Adam update are part of the captured step graph.
        split = bb.grad_split_offset()
        self._buckets = [g for _, g in self.groups[1:]] + [self.bb_g[split:], self.bb_g[:split]]
        self._n_head_buckets = len(self.groups) - 1
        self.reducer = GradReducer(self._buckets, process_group)
        self.step_count = 0
        self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev)      # Adam's step counter, incremented on the device
        self._head_launched = self._deep_launched = False
        self._graph_updates = False       # does the current graph contain the all-reduce + Adam update?
        self._nccl_captured = False       # has this Trainer captured NCCL work into a graph before? (see capture())
        self._defer_comm = False          # a graph WITHOUT the exchange is being captured: the hooks must not launch it
        self._graph = None
        self._graph_ssl = False
        self._graph_cache: dict = {}    # flavour (with / without the SSL pass) -> the captured graph that is not current
        self.target_slots = 8          # ground-truth slots per clip of a captured graph (more in a batch -> capture() again)
        self.graph_update = True       # capture the gradient exchange + Adam update into the step graph (False: they follow the replay)
        self._static = None
        self._capturing = self._warming_up = False
        # Optimizer launches off the step's tail: a parameter range is updated as soon as its (exchanged) gradient is final — the
        # head's two groups when the backbone backward starts (73 % of the parameters), Mixed_4b..5c when their bucket is back — on
        # a side stream next to the remaining backward; only the first layers' update is left behind the last weight gradient.
        # OFF by default: on one GPU it measures the same (17.28 / 17.28 / 17.25 / 17.25 ms, off / on / off / on,
        # profiles/r02_early_adam_n1.txt — the update competes with the HBM-bound first kernels of the backbone backward for the
        # same bandwidth it would use at the tail); OTAL_EARLY_ADAM=1 switches it on.
        self.early_update = os.environ.get("OTAL_EARLY_ADAM", "0") == "1"
        self._early = False               # inside step() / the captured body: the hooks may launch optimizer work
        self._updated: list = []          # (group, lo, hi) ranges already updated in this step
        self._counted = False             # has the device-side step counter been advanced in this step?
        self._opt_stream = None
        self._bb_split = split
        bb.on_backward_start = self._on_backbone_backward
        bb.on_deep_done = self._on_backbone_deep_done if self.world > 1 else None      # (its join of the weight-gradient stream
        #                                                                               is not worth an early update alone)

    # ---------------------------------------------------------------------------------------------- data parallel
    def _comm_allowed(self) -> bool:
        # the warm-up passes of capture() are real passes whose gradients are thrown away: no exchange, no update; a graph that is
        # captured without the exchange gets it from step() after the replay
        return not self._warming_up and not self._defer_comm

    def _on_backbone_backward(self) -> None:
        if self._comm_allowed():
            if self.world > 1:
                self._launch_head_allreduce()
            self._early_update("head")

    def _on_backbone_deep_done(self) -> None:
        if self._comm_allowed() and not self._deep_launched:
            self._launch_head_allreduce()
            self.reducer.launch([self._n_head_buckets])         # Mixed_4b..5c
            self._deep_launched = True
            self._early_update("deep")

    def _launch_head_allreduce(self) -> None:
        if self._head_launched:                                # once per step, whoever asks first
            return
        self.reducer.launch(range(self._n_head_buckets))       # head buffers: complete once the backbone backward starts
        self._head_launched = True

    def _adam(self, gi: int, lo: int = 0, hi: int | None = None) -> None:
        (w, g), st = self.groups[gi], self.state[gi]
        hi = w.numel() if hi is None else hi
        if hi <= lo:
            return
        lr = self.lr * self.backbone_lr_scale if gi == 0 else self.lr         # group 0 = the backbone's flat buffer
        ops.adam_step(w[lo:hi], g[lo:hi], st["m"][lo:hi], st["v"][lo:hi], lr=lr, betas=self.betas, eps=self.eps,
                      weight_decay=self.wd, grad_scale=self.reducer.grad_scale, step_dev=self._step_dev)
        self._updated.append((gi, lo, hi))

    def _count_step(self) -> None:
        if not self._counted:
            self._step_dev.add_(1)
            self._counted = True

    def _early_update(self, which: str) -> None:
        """Adam for the ranges whose gradients are final at this point of the backward, on the optimizer side stream (it waits for
        everything enqueued so far and, data parallel, for the ranges' all-reduces; `_finish_step` joins it).  Only inside a
        step: a bare forward_backward() leaves the weights alone."""
        if not (self._early and self.early_update):
            return
        cuda = self.device.type == "cuda"
        if cuda:
            if self._opt_stream is None:
                self._opt_stream = torch.cuda.Stream(self.device)
            self._opt_stream.wait_stream(torch.cuda.current_stream())
        with (torch.cuda.stream(self._opt_stream) if cuda else contextlib.nullcontext()):
            if not any(gi > 0 for gi, _, _ in self._updated):              # the head's groups (also when "deep" comes first)
                self.reducer.wait(range(self._n_head_buckets))
                self._count_step()
                for gi in range(1, len(self.groups)):
                    self._adam(gi)
            if which == "deep":
                self.reducer.wait([self._n_head_buckets])
                self._adam(0, self._bb_split, None)

    def _finish_step(self) -> None:
        """Everything after the backward: the buckets not exchanged yet, the wait, the step counter and the Adam launches of the
        ranges that were not updated during the backward.  Part of the captured graph (NCCL collectives and the device-side step
        counter capture like any other launch)."""
        ops.nvtx_push("otal.exchange + adam")
        if self.world > 1:
            self._launch_head_allreduce()
            if not self._deep_launched:
                self.reducer.launch([self._n_head_buckets])
            self.reducer.launch([self._n_head_buckets + 1])
        if self._updated and self._opt_stream is not None:
            torch.cuda.current_stream().wait_stream(self._opt_stream)          # join: early updates (and the waits they made)
        if self.world > 1:
            self.reducer.wait()
            self._head_launched = self._deep_launched = False
        self._count_step()
        done = {gi: (lo, hi) for gi, lo, hi in self._updated}
        for gi in range(len(self.groups)):
            if gi not in done:
                self._adam(gi)
            else:                                           # the part of the group in front of the early range (backbone: first layers)
                self._adam(gi, 0, done[gi][0])
        self._updated, self._counted = [], False
        ops.nvtx_pop()

    def broadcast_parameters(self, src: int = 0) -> None:
        if self.world > 1:
            for t in [w for w, _ in self.groups] + [self.net.backbone.flat_bn]:
                dist.broadcast(t, src, group=self.pg)

    # ---------------------------------------------------------------------------------------------- one step
    def zero_grad(self) -> None:
        self.net.backbone._pending_bwd = 0      # a forward without backward (evaluation) must not hold back the next step's hook
        self.net.backbone._bind_grads()
        if self.hc is not None:
            self.hc.bind_grads()
        self.head.rebind()
        for _, g in self.groups:
            g.zero_()

    @staticmethod
    def triplet_loss(anchor, positive, negative):
        """nn.TripletMarginLoss() per feature with weights (1, 0.1, 0.1): thumos14/train.py:174-184."""
        weights = (1.0, 0.1, 0.1)
        return torch.stack([F.triplet_margin_loss(anchor[i], positive[i], negative[i]) * weights[i] for i in range(3)]).sum(0)

    def forward_backward(self, clips: torch.Tensor, targets, scores: torch.Tensor, ssl_clips=None, ssl_targets=None,
                         ssl_frame_map=None):
        ops.nvtx_push("otal.forward (backbone + head)")
        try:
            out = self.net(clips)
        finally:
            ops.nvtx_pop()
        anet = getattr(self.net, "variant", "thumos") == "anet"
        ops.nvtx_push("otal.loss")
        if anet:      # the ActivityNet loss takes the list form (anet/train.py:168-172)
            losses = self.criterion([out[k] for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "priors", "act", "prop_act")],
                                    targets)
        else:
            losses = self.criterion(out, targets)
            if self.global_normalisers and self.world > 1:
                losses = self._globalise(out, losses, targets)
                self.criterion.last_vec = None           # the re-weighted tuple is the loss now
        cost, ls, le = training_cost(out, losses, scores, lw=self.lw, cw=self.cw, ctw=self.ctw, actw=self.actw,
                                     score_scale=8 if anet else 4, loss_vec=getattr(self.criterion, "last_vec", None))
        ops.nvtx_pop()
        if ssl_clips is not None or ssl_frame_map is not None:
            # second forward on the cut-paste augmented clip + triplet loss on its boundary features
            # (thumos14/train.py:237-242; BDNet.py:482-503); one backward for the sum.  With `ssl_frame_map` the augmented
            # clip is not materialised: the ingest kernel re-reads the uint8 frames of `clips` through the map
            # (opental_b200/augment.py).
            bb = self.net.backbone
            if ssl_frame_map is not None:
                assert ssl_clips is None and clips.dtype == torch.uint8, "ssl_frame_map re-reads the uint8 frames of `clips`"
                bb.frame_map, ssl_clips = ssl_frame_map, clips
            try:
                a, p, n = self.net(ssl_clips, proposals=ssl_targets, ssl=True)
            finally:
                bb.frame_map = None
            cost = cost + self.ssl_weight * self.triplet_loss(a, p, n)
        ops.nvtx_push("otal.backward (head + loss, backbone, gradient exchange hooks)")
        try:
            cost.backward()
        finally:
            ops.nvtx_pop()
        # Nothing non-detached may outlive this call: a live loss tensor keeps this step's autograd graph (and its AccumulateGrad
        # nodes, which remember the stream they were created on) alive, which breaks a later CUDA-graph capture ("dependency created
        # on uncaptured work in another stream" inside backward()).  The criterion's `last_vec` is such a tensor.
        if getattr(self.criterion, "last_vec", None) is not None:
            self.criterion.last_vec = self.criterion.last_vec.detach()
        return cost.detach(), tuple(l.detach() if l is not None else None for l in losses), ls.detach(), le.detach()

    def _globalise(self, out, losses, targets):
        crit = self.criterion
        stats = crit.last_stats.detach()
        if stats.numel() >= 12:              # the fused kernel's 16-float vector: counts and loss_iouc sit at [7:12]
            stats = stats[7:12]
        factors = global_normaliser_factors(stats[:4].detach().clone(), self.world, self.pg)
        return globalise_losses(losses, stats, factors, live_iou_calibration(crit, out, targets))

    # ---------------------------------------------------------------------------------------------- CUDA graph
    @staticmethod
    def _ssl_sources(ssl_clips, ssl_targets, ssl_frame_map) -> list:
        if ssl_clips is None and ssl_frame_map is None:
            return []
        tg = torch.stack(list(ssl_targets)) if isinstance(ssl_targets, (list, tuple)) else ssl_targets
        return [ssl_clips if ssl_clips is not None else ssl_frame_map, tg]

    def _ssl_args(self, ssl_clips, ssl_frame_map) -> tuple:
        """(ssl_clips, ssl_targets, ssl_frame_map) over the static buffers."""
        if len(self._static) == 4:
            return None, None, None
        src, tg = self._static[4], list(self._static[5].unbind(0))
        return (src, tg, None) if ssl_clips is not None else (None, tg, src)

    def capture(self, clips: torch.Tensor, targets, scores: torch.Tensor, ssl_clips=None, ssl_targets=None,
                ssl_frame_map=None) -> None:
        """Capture zero_grad + forward + loss + backward for this input geometry into ONE CUDA graph.  The head and the
        loss are ~2500 small launches whose host-side enqueue cost (~75 ms per step at batch 8) exceeds the GPU work; a
        graph replay removes it.  Inputs are copied into static buffers before every replay; the gradient all-reduce and
        the Adam launches are part of the graph (`graph_update`; Adam reads its step counter from device memory).
        The graph bakes in `criterion.cls_loss.epoch >= ibm_start`: re-capture when the epoch crosses ibm_start."""
        import gc
        self._release_for_capture(ssl_clips is not None or ssl_frame_map is not None)
        gc.collect()                                 # also drops dead autograd graphs of earlier eager steps (see forward_backward)
        # Data parallel: every captured graph holds the gradient exchange and the Adam update, re-captures included (2 GPUs:
        # tools/probe/recapture_dp.py, profiles/r02_recapture_probe_n2.txt).  OTAL_DP_RECAPTURE=0 keeps the exchange and the update of
        # a re-captured graph behind its replay instead (developer switch).  Single-GPU graphs always contain the update.
        behind = os.environ.get("OTAL_DP_RECAPTURE", "1") == "0" and self.world > 1 and self._nccl_captured
        in_graph = bool(self.graph_update) and not behind
        tgt, valid = pad_targets(targets, clips.device, slots=max(self.target_slots, self._max_segments(targets)))
        srcs = [clips, tgt, valid, scores] + self._ssl_sources(ssl_clips, ssl_targets, ssl_frame_map)
        self._static = [torch.empty_like(t) for t in srcs]
        for d, s in zip(self._static, srcs):
            d.copy_(s)
        c, t, v, sc = self._static[:4]
        ssl_args = self._ssl_args(ssl_clips, ssl_frame_map)
        # one capture stream per Trainer, reused by every re-capture (OTAL_CAP_STREAM=new: a fresh stream per capture, probe switch)
        if getattr(self, "_cap_stream", None) is None or os.environ.get("OTAL_CAP_STREAM", "persist") == "new":
            self._cap_stream = torch.cuda.Stream()
        stream = self._cap_stream
        stream.wait_stream(torch.cuda.current_stream())
        self._capturing = True
        # the warm-ups are real passes over the live criterion: its stateful buffers (the IBM EMA `weight_accum`, GHM's
        # `acc_sum`) must come out of capture() as they went in — the reference never makes these two extra updates
        crit_state = [(b, b.detach().clone()) for b in self.criterion.buffers()]
        self._step_dev.fill_(self.step_count)        # the graph increments it from here (the host mirror follows in step())
        try:
            self._warming_up = True
            with torch.cuda.stream(stream):
                for _ in range(2):                   # warm-up on the capture stream (lazy initialisations, allocator)
                    self.zero_grad()
                    self.forward_backward(c, (t, v), sc, *ssl_args)
                for b, saved in crit_state:
                    b.copy_(saved)
            self._warming_up = False
            torch.cuda.current_stream().wait_stream(stream)
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread may query events while this thread captures
            self._defer_comm = not in_graph
            with torch.cuda.graph(self._graph, stream=stream, capture_error_mode="thread_local"):
                self.zero_grad()
                self._early = in_graph
                self._graph_out = self.forward_backward(c, (t, v), sc, *ssl_args)
                if in_graph:
                    self._finish_step()
        finally:
            self._capturing = self._warming_up = self._defer_comm = self._early = False
            self._updated, self._counted = [], False
            self._head_launched = self._deep_launched = False
        self._graph_updates = in_graph
        self._nccl_captured = self._nccl_captured or (in_graph and self.world > 1)
        self._graph_epoch_flag = self._ibm_flag()
        self._graph_ssl = ssl_clips is not None or ssl_frame_map is not None

    def _release_for_capture(self, new_ssl: bool) -> None:
        """Graph bookkeeping at the start of capture(): a current graph of the OTHER flavour is kept for select_graph; a graph
        of the flavour being captured — current or cached, i.e. one with the other IBM switch or too few target slots — is
        released before the new one takes its memory."""
        if self._graph is not None and self._graph_ssl != new_ssl:
            self._stash()
        self._graph_cache.pop(new_ssl, None)
        self._graph = self._static = self._graph_out = None

    @staticmethod
    def _max_segments(targets) -> int:
        """Largest number of ground-truth segments of a sample (0 for an already padded (tensor, mask) pair: its geometry is
        compared as is)."""
        if isinstance(targets, (tuple, list)) and len(targets) == 2 and torch.is_tensor(targets[0]) and targets[0].dim() == 3:
            return 0
        return max(int(t.shape[0]) for t in targets)

    def graph_matches(self, ssl: bool, targets=None) -> bool:
        """Is there a captured step graph for this flavour of batch (with / without the SSL pass), the current epoch's IBM
        switch and (when `targets` is given) enough ground-truth slots?  (Clip geometry is checked by step() itself.)"""
        if self._graph is None or self._graph_ssl != bool(ssl) or self._ibm_flag() != self._graph_epoch_flag:
            return False
        return targets is None or self._max_segments(targets) <= self._static[1].shape[1]

    def _stash(self) -> None:
        if self._graph is not None:
            self._graph_cache[self._graph_ssl] = (self._graph, self._static, self._graph_out, self._graph_epoch_flag, self._graph_updates)

    def select_graph(self, ssl: bool, targets=None) -> bool:
        """Make a captured graph that fits this batch current (see graph_matches); False = the caller must capture().
        One graph per flavour is kept, so batches alternating between with / without the SSL pass (train.py:237: only when
        the first sample could be augmented) do not re-capture."""
        if self.graph_matches(ssl, targets):
            return True
        ent = self._graph_cache.get(bool(ssl))
        if ent is None or ent[3] != self._ibm_flag() or (targets is not None and self._max_segments(targets) > ent[1][1].shape[1]):
            return False
        self._stash()
        self._graph, self._static, self._graph_out, self._graph_epoch_flag, self._graph_updates = ent
        self._graph_ssl = bool(ssl)
        return True

    def _ibm_flag(self):
        c = self.criterion.cls_loss
        return bool(getattr(c, "with_ibm", False) and c.epoch >= getattr(c, "ibm_start", 0))

    def step(self, clips: torch.Tensor, targets, scores: torch.Tensor, ssl_clips=None, ssl_targets=None, ssl_frame_map=None):
        """clips [B,3,T,H,W] fp32 (or uint8 frames [B,T,Hs,Ws,3]) on the device, targets: list of [N_i,3] or padded
        (tensor, mask), scores [B,2,T].  ssl_clips / ssl_targets: the cut-paste augmented clip and its [3,2] (anchor,
        positive, negative) segments for the triplet pass (thumos14/train.py:237-242), or None.  ssl_frame_map (int32
        [B,T], from opental_b200.augment.cut_paste) replaces ssl_clips when `clips` are uint8 frames."""
        if self._graph is not None:
            if self._max_segments(targets) > self._static[1].shape[1]:
                raise RuntimeError("captured training graph does not match this batch geometry / epoch: call capture() again")
            tgt, valid = pad_targets(targets, clips.device, slots=self._static[1].shape[1])
            srcs = [clips, tgt, valid, scores] + self._ssl_sources(ssl_clips, ssl_targets, ssl_frame_map)
            if (len(srcs) != len(self._static) or any(tuple(a.shape) != tuple(b.shape) or a.dtype != b.dtype for a, b in zip(srcs, self._static))
                    or self._ibm_flag() != self._graph_epoch_flag):
                raise RuntimeError("captured training graph does not match this batch geometry / epoch: call capture() again")
            for d, s in zip(self._static, srcs):
                if d.data_ptr() != s.data_ptr():
                    d.copy_(s, non_blocking=True)
            self._graph.replay()
            cost, losses, ls, le = self._graph_out
            if not self._graph_updates:
                self._finish_step()
        else:
            self.zero_grad()
            self._early, self._updated, self._counted = True, [], False
            try:
                cost, losses, ls, le = self.forward_backward(clips, targets, scores, ssl_clips, ssl_targets, ssl_frame_map)
            finally:
                self._early = False
            self._finish_step()
        self.step_count += 1
        return cost, losses, ls, le


    # ---------------------------------------------------------------------------------------------- bookkeeping
    def grad_norm(self) -> torch.Tensor:
        """`get_grad_norm` (thumos14/train.py:132-139): global L2 norm of the gradients of all trainable parameters, as a
        0-dim device tensor (no synchronisation).  Three reductions over the flat buffers (their alignment padding is
        zero) instead of 275 per-parameter norms.  After `step()` the buffers hold the rank-summed gradients; the 1/world
        factor the Adam kernel applies is applied here too, so the value is the norm of the averaged gradient."""
        sq = torch.stack([torch.linalg.vector_norm(g) ** 2 for _, g in self.groups]).sum()
        return sq.sqrt() * self.reducer.grad_scale

    def _param_groups(self):
        """None for the single-rate optimizer (thumos14/train.py:321); the ActivityNet script's two groups — backbone at
        backbone_lr_scale x the rate first, then the head (anet/train.py:304-311) — otherwise."""
        if self.backbone_lr_scale == 1.0:
            return None
        return [(list(self.net.backbone.parameters()), self.lr * self.backbone_lr_scale),
                (list(self.net.coarse_pyramid_detection.parameters()), self.lr)]

    def optimizer_state_dict(self) -> dict:
        """The `torch.optim.Adam(...).state_dict()` a reference run would hold at this point (train.py:115)."""
        from . import checkpoint
        return checkpoint.adam_state_dict(list(self.net.parameters()), self.groups, self.state, step=self.step_count,
                                          lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.wd,
                                          param_groups=self._param_groups())

    def load_optimizer_state_dict(self, sd: dict) -> None:
        from . import checkpoint
        self.step_count = checkpoint.load_adam_state_dict(sd, list(self.net.parameters()), self.groups, self.state,
                                                          param_groups=self._param_groups())
        g = sd["param_groups"][-1]                  # the head's group carries the base rate
        self.lr, self.betas, self.eps, self.wd = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]
        self._step_dev.fill_(self.step_count)
        # a captured step graph has the optimizer's hyper-parameters baked into its Adam launches: capture again
        self._graph = self._static = self._graph_out = None
        self._graph_cache.clear()

    def save_checkpoint(self, epoch: int, checkpoint_path: str, train_state_path: str):
        """`save_model` (train.py:106-118): model state_dict + {'optimizer', 'state'} in the reference's file layout."""
        from . import checkpoint
        return checkpoint.save(self, epoch, checkpoint_path, train_state_path)

    def resume(self, resume_epoch: int, checkpoint_path: str, train_state_path: str) -> int:
        """`resume_training` (train.py:121-131)."""
        from . import checkpoint
        return checkpoint.resume(self, resume_epoch, checkpoint_path, train_state_path)


# ----------------------------------------------------------------------------------------------------------
# synthetic workload (SURVEY §8d conventions) — shared by bench.py, smoke() and the tests
# ----------------------------------------------------------------------------------------------------------
def synthetic_clip_u8(index: int, rank: int = 0, frames: int = 256) -> torch.Tensor:
    """uint8 [T,112,112,3] i.i.d. uniform pixels, the npy format of AFSD/common/video2npy.py:61-74."""
    g = torch.Generator().manual_seed(1000 * rank + index)
    return torch.randint(0, 256, (frames, 112, 112, 3), generator=g, dtype=torch.uint8)


def normalise_clip(px: torch.Tensor, crop: int = 96) -> torch.Tensor:
    """centre crop + (x/255)*2-1 -> fp32 [3,T,crop,crop] (thumos_dataset.py:261-263)."""
    o = (112 - crop) // 2
    px = px[:, o:o + crop, o:o + crop, :]
    return (px.permute(3, 0, 1, 2).float() / 255.0) * 2.0 - 1.0


def synthetic_targets(index: int, rank: int = 0, num_classes: int = 15) -> torch.Tensor:
    g = torch.Generator().manual_seed(7_000_000 + 1000 * rank + index)
    rows = []
    for j in range(2):
        s = 0.10 + 0.45 * j + (torch.rand((), generator=g).item() * 0.06 - 0.03)
        e = s + 0.25 + (torch.rand((), generator=g).item() * 0.06 - 0.03)
        lab = int(torch.randint(1, num_classes + 1, (), generator=g).item())
        rows.append([s, e, float(lab)])
    return torch.tensor(rows, dtype=torch.float32)


def synthetic_scores(targets: torch.Tensor, frames: int = 256) -> torch.Tensor:
    """start/end score maps [2,T] of a clip's (normalised) targets — the loader's rule, thumos_dataset.py:109-120."""
    from .windows import boundary_score_maps
    start, end = boundary_score_maps([[s * frames, e * frames, l] for s, e, l in targets.tolist()], frames)
    return torch.from_numpy(np.stack([start, end])).float()


OPENTAL_EDL_CONFIG = dict(evidence="exp", loss_type="log", with_ibm=True, ibm_start=10, momentum=0.99, num_bins=50,
                          iou_aware=True)          # configs/thumos14_opental_final.yaml:38-49
OPENTAL_ACT_CONFIG = dict(weight=0.0, margin=1.0)  # :50-52


def build_opental_anet(device="cuda", precision="bf16x3", epoch=1):
    """BDNet + MultiSegmentLoss as constructed for configs/anet_opental.yaml --open_set (768-frame clips, 150 classes)."""
    from .multisegment_loss import MultiSegmentLossANet
    net = BDNet(in_channels=3, training=True, use_edl=True, num_classes=151, os_head=True, frame_num=768, precision=precision,
                variant="anet").to(device)
    crit = MultiSegmentLossANet(150, 0.5, 1.0, cls_loss_type="edl", edl_config=OPENTAL_EDL_CONFIG, os_head=True).to(device)
    crit.cls_loss.epoch = epoch
    net.train()
    return net, crit


def build_opental(device="cuda", precision="bf16x3", frame_num=256, epoch=1):
    """BDNet + MultiSegmentLoss as constructed for configs/thumos14_opental_final.yaml --open_set."""
    net = BDNet(in_channels=3, training=True, use_edl=True, num_classes=16, os_head=True, frame_num=frame_num,
                precision=precision).to(device)
    crit = MultiSegmentLoss(15, 0.5, 1.0, cls_loss_type="edl", edl_config=OPENTAL_EDL_CONFIG, os_head=True,
                            act_config=OPENTAL_ACT_CONFIG, clip_length=frame_num).to(device)
    crit.cls_loss.epoch = epoch
    net.train()
    return net, crit


def smoke_step() -> None:
    """One tiny BDNet training step on cuda:0 (short 64-frame clip would break the 256-frame head, so: one full clip)."""
    torch.manual_seed(0)
    net, crit = build_opental()
    tr = Trainer(net, crit)
    clip = normalise_clip(synthetic_clip_u8(0)).unsqueeze(0).cuda()
    tgt = [synthetic_targets(0).cuda()]
    sc = synthetic_scores(tgt[0].cpu()).unsqueeze(0).cuda()
    cost, losses, ls, le = tr.step(clip, tgt, sc)
    torch.cuda.synchronize()
    vals = [float(cost)] + [float(v) for v in losses]
    assert all(v == v and abs(v) < 1e6 for v in vals), vals
    print("smoke training step ok: cost %.4f, losses %s" % (vals[0], ["%.4f" % v for v in vals[1:]]))
