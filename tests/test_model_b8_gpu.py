"""GPU parity of the BENCH configuration and of a multi-step trajectory, against goldens produced by the reference's own code:

* tests/golden/model_thumos_b8.* (oracle/make_golden.py --batch8): reference BDNet forward + MultiSegmentLoss (epoch 1, 11) +
  backward on synthetic clips 0..7 with ragged targets (1, 2 and 3 segments) — batch 8 per GPU is what bench.py measures
  (BASELINE configs[1]/[2]); B > 1 exercises the `[P,B]`-vs-`[B,P]` IoU-calibration pairing of the reference
  through the whole model.
* tests/golden/trajectory_thumos.* (--trajectory): five optimizer steps of the reference modules under torch.optim.Adam
  (thumos14/train.py:226-252, 321-323) on a batch of two clips at epoch 11 (IBM EMA evolving).  `Trainer.step` must follow the
  cost step by step — eagerly and through the captured CUDA graph.

Tolerances: outputs / losses 1e-3 relative (BASELINE.json); gradient fingerprints 5e-2 (discontinuous: ReLU / arg-max flips);
trajectory cost max(1e-3, 3 x the largest drift recorded between the reference and its fp32 CPU restatement up to that step)."""
import json
import math
import os

import numpy as np
import pytest
import torch

import opental_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def build(epoch):
    from opental_b200 import engine
    net, crit = engine.build_opental(epoch=epoch)
    net.load_state_dict(O.synthetic_state_dict(O.OracleConfig(), loc_bias_shift=math.log(32.0)))
    return net, crit


def b8_targets():
    targets = [O.synthetic_targets(i, num_classes=15) for i in range(8)]
    targets[3] = targets[3][:1].clone()
    targets[5] = torch.cat([targets[5], torch.tensor([[0.42, 0.47, 3.0]])])
    return targets


def test_batch8_forward_loss_backward_match_reference_golden(golden_dir):
    from opental_b200.prop_pooling import BoundaryMaxPoolingFunction
    arrays = np.load(os.path.join(golden_dir, "model_thumos_b8.npz"))
    with open(os.path.join(golden_dir, "model_thumos_b8.json")) as fh:
        summary = json.load(fh)
    net, crit = build(11)
    x = torch.stack([O.synthetic_clip(i) for i in range(8)]).cuda()
    targets = [t.cuda() for t in b8_targets()]
    assert [int(t.shape[0]) for t in targets] == summary["n_segments"]
    out = net(x)
    # Coarse outputs: max-norm 1e-3.  Refined outputs sit behind the proposal windows, whose ends are rounded to frames
    # (BDNet.py:355-384): a 1e-6 difference in `loc` can move a window end across an integer and change that prior's pooled
    # maxima — the reference and its fp32 restatement differ that way in 1..2 of the 1008 rows (oracle/make_golden.py).  So the
    # refined outputs are compared row-wise: all but a handful of (clip, prior) rows within 1e-3, every row within 5e-2.
    errs, flipped = {}, set()
    for k in ("loc", "conf", "act", "unct"):
        errs[k] = rel(out[k].detach().cpu(), torch.from_numpy(arrays[f"b8.{k}"]))
    for k in ("prop_loc", "prop_conf", "center", "prop_act", "prop_unct"):
        ref = torch.from_numpy(arrays[f"b8.{k}"])
        d = (out[k].detach().cpu() - ref).abs().reshape(8, 126, -1).amax(-1) / ref.abs().max()
        flipped |= {tuple(i) for i in (d > 1e-3).nonzero().tolist()}
        assert float(d.max()) < 5e-2, (k, float(d.max()))
    assert len(flipped) <= 8, sorted(flipped)
    for k in ("start", "end", "start_loc_prop", "end_loc_prop", "start_conf_prop", "end_conf_prop"):
        errs[k] = rel(out[k].detach().cpu()[:, ::8, ::8], torch.from_numpy(arrays[f"b8.{k}.sample"]))
    assert max(errs.values()) < 1e-3, errs
    for epoch in (1, 11):
        crit.cls_loss.epoch = epoch
        crit.cls_loss.weight_accum = torch.ones(50, device="cuda")
        losses = crit(out, targets)
        for a, b in zip(losses, summary[f"e{epoch}"]["losses"]):
            assert abs(float(a) - b) <= 1e-3 * max(abs(b), 1.0), (epoch, float(a), b)
        assert np.allclose(crit.cls_loss.weight_accum.cpu().numpy(), arrays[f"b8.e{epoch}.weight_accum"], atol=1e-5)
    net.backbone.flat_parameters()[1].zero_()
    cost = losses[0] + 10 * losses[1] + losses[2] + 10 * losses[3] + losses[4] + losses[5] + losses[6]
    assert abs(float(cost) - summary["e11"]["cost"]) < 1e-3 * abs(summary["e11"]["cost"])
    BoundaryMaxPoolingFunction.compat_tscale_bug = True        # the golden gradients come from the reference kernel
    try:
        cost.backward()
    finally:
        BoundaryMaxPoolingFunction.compat_tscale_bug = False
    params = dict(net.named_parameters())
    bad = {}
    for k, (s, a) in summary["e11"]["grad_fingerprint"].items():
        g = params[k].grad
        assert g is not None, k
        if a > 0 and abs(float(g.abs().sum()) - a) / a > 5e-2:
            bad[k] = abs(float(g.abs().sum()) - a) / a
        smp = torch.from_numpy(arrays[f"b8.e11.grad.{k}"])
        got = g.detach().cpu().reshape(-1)[:: max(1, g.numel() // 64)][:64]
        if smp.abs().max() > 0 and rel(got, smp) > 0.2:
            bad[k + ":sample"] = rel(got, smp)
    assert not bad, bad


@pytest.mark.parametrize("graph", [False, True])
def test_five_step_trajectory_follows_the_reference(golden_dir, graph):
    from opental_b200 import engine
    from opental_b200.prop_pooling import BoundaryMaxPoolingFunction
    with open(os.path.join(golden_dir, "trajectory_thumos.json")) as fh:
        gold = json.load(fh)
    w_acc = np.load(os.path.join(golden_dir, "trajectory_thumos.npz"))["weight_accum"]
    net, crit = build(11)
    tr = engine.Trainer(net, crit, lr=gold["lr"], weight_decay=gold["weight_decay"])
    w0 = {k: p.detach().clone() for k, p in net.named_parameters() if p.requires_grad}
    x = torch.stack([O.synthetic_clip(i) for i in range(2)]).cuda()
    tg = [O.synthetic_targets(i, num_classes=15).cuda() for i in range(2)]
    sc = torch.stack([O.synthetic_scores(t.cpu()) for t in tg]).cuda()
    BoundaryMaxPoolingFunction.compat_tscale_bug = True        # the reference trajectory ran the reference kernel's backward
    try:
        if graph:
            acc0 = crit.cls_loss.weight_accum.clone()
            tr.capture(x, tg, sc)
            crit.cls_loss.weight_accum.copy_(acc0)             # capture() warm-ups touch the EMA (restored by the Trainer too)
        for s, want in enumerate(gold["steps"]):
            cost, losses, ls, le = tr.step(x, tg, sc)
            # once two implementations have bifurcated (a matching flip), every later step inherits the difference.  From the
            # fourth step on the product also bifurcates against ITSELF from run to run (the atomically reduced weight gradients
            # are summed in a different order, Adam's first steps turn sign noise into +-lr, and a prior whose IoU sits at the
            # refinement threshold changes sides: measured 5.5e-3 on the cost in one run out of four, 1e-4 in the others), so
            # those steps get the bar of a flip (1e-2) — the optimizer itself is pinned by the first three steps and by delta_total
            tol = max(1e-3 if s < 3 else 1e-2, 3 * max(w["oracle_rel"] for w in gold["steps"][:s + 1]))
            assert abs(float(cost) - want["cost"]) <= tol * abs(want["cost"]), (s, float(cost), want["cost"])
            if s < 3:
                for a, b in zip(losses, want["losses"]):
                    assert abs(float(a) - b) <= 2e-3 * max(abs(b), 1.0), (s, float(a), b)
                assert abs(float(ls) - want["loss_start"]) <= 1e-3 and abs(float(le) - want["loss_end"]) <= 1e-3
    finally:
        BoundaryMaxPoolingFunction.compat_tscale_bug = False
    assert tr.step_count == len(gold["steps"])
    # the 50-bin IBM EMA after five steps.  The bin of a sample is ceil(grad_norm * 50) (cls_loss.py:263): a sample whose
    # grad_norm sits within the accumulated fp32 difference of a bin edge lands in the neighbouring bin and moves BOTH bins' EMA
    # by about (1 - momentum) x the difference of their per-bin means (measured on B200: 0.023..0.035 per event, in +/- pairs;
    # the number of such events varies from run to run with the summation order of the atomically reduced weight gradients,
    # which Adam's first steps amplify: lr * sign(g) for the elements whose gradient is noise).  The exact single-update check
    # of the EMA is test_batch8_forward_loss_backward_match_reference_golden (atol 1e-5); here: no bin far off, and small on average.
    d_acc = np.abs(crit.cls_loss.weight_accum.cpu().numpy() - w_acc)
    assert d_acc.max() < 0.12 and d_acc.mean() < 1.5e-2, np.round(d_acc, 4).tolist()
    # how far the parameters moved: Adam's first steps are ~lr per element per step, so the total |delta w| is a tight check of
    # the optimizer (bias correction, L2-in-gradient, step counter) even where single elements flip sign
    total = sum(float((p.detach() - w0[k]).abs().sum()) for k, p in net.named_parameters() if p.requires_grad)
    assert abs(total - gold["delta_total"]) <= 2e-2 * gold["delta_total"], (total, gold["delta_total"])
