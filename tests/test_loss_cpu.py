"""CPU: the product's fixed-shape, sync-free MultiSegmentLoss (opental_b200/multisegment_loss.py; plain torch ops, so it
runs on CPU tensors too) against the oracle restatement of the reference loss, which itself is pinned to the
reference-generated golden values (tests/test_oracle_golden.py).  Tolerance 1e-5 relative: only summation order differs."""
import os
import math

import pytest
import torch

import opental_oracle as O
from opental_b200.multisegment_loss import MultiSegmentLoss, training_cost
from opental_b200.engine import OPENTAL_EDL_CONFIG


def fake_outputs(B, seed, loc_scale=30.0):
    g = torch.Generator().manual_seed(seed)
    P = 126
    cfg = O.OracleConfig()
    out = dict(loc=(torch.rand(B, P, 2, generator=g) * loc_scale + 1).requires_grad_(True),
               conf=torch.randn(B, P, 15, generator=g).requires_grad_(True),
               prop_loc=(0.3 * torch.randn(B, P, 2, generator=g)).requires_grad_(True),
               prop_conf=torch.randn(B, P, 15, generator=g).requires_grad_(True),
               center=torch.randn(B, P, 1, generator=g).requires_grad_(True),
               priors=torch.cat(O.level_priors(cfg), 0),
               act=torch.randn(B, P, 1, generator=g).requires_grad_(True),
               prop_act=torch.randn(B, P, 1, generator=g).requires_grad_(True))
    return out, cfg


def make_crit(epoch, act_weight=0.0):
    crit = MultiSegmentLoss(15, 0.5, 1.0, cls_loss_type="edl", edl_config=OPENTAL_EDL_CONFIG, os_head=True,
                            act_config=dict(weight=act_weight, margin=1.0))
    crit.cls_loss.epoch = epoch
    return crit


@pytest.mark.parametrize("B,epoch,act_weight", [(1, 1, 0.0), (1, 11, 0.0), (3, 11, 0.0), (2, 11, 0.1), (4, 1, 0.1)])
def test_loss_matches_oracle(B, epoch, act_weight):
    out, cfg = fake_outputs(B, seed=10 * B + epoch)
    cfg.act_weight = act_weight
    targets = [O.synthetic_targets(i, num_classes=15) for i in range(B)]
    if B >= 3:
        targets[1] = targets[1][:1]                      # ragged number of ground-truth segments
    state = O.LossState(epoch=epoch)
    ref = O.multisegment_loss(out, targets, state, cfg)
    crit = make_crit(epoch, act_weight)
    got = crit(out, targets)
    for a, b in zip(got, ref):
        assert abs(float(a) - float(b)) <= 1e-5 * max(1.0, abs(float(b))), (float(a), float(b))
    assert torch.allclose(crit.cls_loss.weight_accum, state.weight_accum, atol=1e-6)
    # gradients of the weighted cost w.r.t. every head output
    keys = ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act")
    w = (1, 10, 1, 10, 1, 1, 1)
    g_ref = torch.autograd.grad(sum(wi * li for wi, li in zip(w, ref)), [out[k] for k in keys], allow_unused=True)
    g_got = torch.autograd.grad(sum(wi * li for wi, li in zip(w, got)), [out[k] for k in keys], allow_unused=True)
    for k, a, b in zip(keys, g_got, g_ref):
        assert torch.allclose(a, b, atol=1e-6, rtol=1e-4), k


def test_no_positive_priors():
    out, cfg = fake_outputs(2, seed=3)
    targets = [torch.tensor([[1.2, 1.4, 3.0]]), torch.tensor([[1.5, 1.9, 4.0]])]     # outside [0,1]: no prior inside
    got = make_crit(11)(out, targets)
    ref = O.multisegment_loss(out, targets, O.LossState(epoch=11), cfg)
    assert float(got[0]) == 0 and float(got[1]) == 0 and float(got[2]) == 0
    for a, b in zip(got, ref):
        assert math.isfinite(float(a)) and abs(float(a) - float(b)) <= 1e-5 * max(1.0, abs(float(b)))


def test_training_cost_matches_oracle():
    out, cfg = fake_outputs(2, seed=5)
    g = torch.Generator().manual_seed(6)
    for k, shp in (("start", (2, 256, 256)), ("end", (2, 256, 256)), ("start_loc_prop", (2, 64, 512)),
                   ("end_loc_prop", (2, 64, 512)), ("start_conf_prop", (2, 64, 512)), ("end_conf_prop", (2, 64, 512))):
        out[k] = torch.randn(*shp, generator=g).relu()
    targets = [O.synthetic_targets(i, num_classes=15) for i in range(2)]
    scores = torch.stack([O.synthetic_scores(t) for t in targets])
    ref_cost, _ = O.training_cost(out, targets, scores, O.LossState(epoch=1), cfg)
    crit = make_crit(1)
    cost, ls, le = training_cost(out, crit(out, targets), scores)
    assert abs(float(cost) - float(ref_cost)) < 1e-5 * abs(float(ref_cost))


def test_focal_configuration():
    """configs/thumos14.yaml flavour (config 1): 21 classes incl. background, softmax focal loss on every prior."""
    g = torch.Generator().manual_seed(9)
    P = 126
    cfg = O.OracleConfig()
    out = dict(loc=torch.rand(1, P, 2, generator=g) * 30 + 1, conf=torch.randn(1, P, 21, generator=g),
               prop_loc=0.3 * torch.randn(1, P, 2, generator=g), prop_conf=torch.randn(1, P, 21, generator=g),
               center=torch.randn(1, P, 1, generator=g), priors=torch.cat(O.level_priors(cfg), 0), act=None, prop_act=None)
    crit = MultiSegmentLoss(21, 0.5, 1.0, cls_loss_type="focal")
    l = crit(out, [O.synthetic_targets(0, num_classes=20)])
    assert l[5] is None and l[6] is None and all(math.isfinite(float(v)) for v in l[:5])
    # hand check of the focal term on the coarse logits
    loc_t, conf_t, *_ = crit.match(out["loc"], out["priors"], *__import__("opental_b200.multisegment_loss", fromlist=["pad_targets"]).pad_targets([O.synthetic_targets(0, num_classes=20)]))
    p = torch.softmax(out["conf"].view(-1, 21), 1).gather(1, conf_t.view(-1, 1)).view(-1) + 1e-6
    alpha = torch.where(conf_t.view(-1) == 0, torch.tensor(0.25), torch.tensor(0.75))
    want = (-(1 - p) ** 2 * alpha * p.log()).sum() / max(int((conf_t > 0).sum()), 1)
    assert abs(float(l[1]) - float(want)) < 1e-5 * abs(float(want))


@pytest.mark.parametrize("B,epoch", [(1, 1), (2, 11), (3, 11)])
def test_anet_loss_matches_oracle(B, epoch):
    """ActivityNet flavour (per-sample normalisation, level-gated matching, smooth-L1, exp-form IBM) vs the oracle
    restatement, which oracle/make_golden.py --anet pins to the reference's own AFSD/anet code."""
    from opental_b200.multisegment_loss import MultiSegmentLossANet
    cfg = O.anet_config()
    g = torch.Generator().manual_seed(40 + B + epoch)
    P, K = 189, cfg.num_classes
    priors = torch.cat(O.level_priors(cfg), 0)
    stride = torch.tensor([O.ANET_FPN_STRIDES[int(l)] for l in priors[:, 1]])
    out = dict(loc=((torch.rand(B, P, 2, generator=g) * 6 + 0.5) * stride[None, :, None]).requires_grad_(True),
               conf=torch.randn(B, P, K, generator=g).requires_grad_(True),
               prop_loc=(0.3 * torch.randn(B, P, 2, generator=g)).requires_grad_(True),
               prop_conf=torch.randn(B, P, K, generator=g).requires_grad_(True),
               center=torch.randn(B, P, 1, generator=g).requires_grad_(True), priors=priors,
               act=torch.randn(B, P, 1, generator=g).requires_grad_(True),
               prop_act=torch.randn(B, P, 1, generator=g).requires_grad_(True))
    targets = [O.synthetic_targets(i, num_classes=K) for i in range(B)]
    if B >= 2:
        targets[1] = torch.cat([targets[1], torch.tensor([[0.40, 0.44, 17.0]])])
    if B >= 3:
        targets[2] = torch.tensor([[1.2, 1.4, 3.0]])                 # no prior inside: no positives in this sample
    ref = O.multisegment_loss_anet(out, targets, O.LossState(epoch=epoch), cfg)
    crit = MultiSegmentLossANet(K, 0.5, 1.0, cls_loss_type="edl", edl_config=OPENTAL_EDL_CONFIG, os_head=True)
    crit.cls_loss.epoch = epoch
    keys = ("loc", "conf", "prop_loc", "prop_conf", "center", "priors", "act", "prop_act")
    got = crit([out[k] for k in keys], targets)
    for a, b in zip(got, ref):
        assert abs(float(a) - float(b)) <= 2e-5 * max(1.0, abs(float(b))), (float(a), float(b))
    gk = ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act")
    w = (1, 10, 1, 10, 1, 1, 1)
    g_ref = torch.autograd.grad(sum(wi * li for wi, li in zip(w, ref)), [out[k] for k in gk], allow_unused=True)
    g_got = torch.autograd.grad(sum(wi * li for wi, li in zip(w, got)), [out[k] for k in gk], allow_unused=True)
    for k, a, b in zip(gk, g_got, g_ref):
        b = torch.zeros_like(a) if b is None else b
        assert torch.allclose(a, b, atol=2e-6, rtol=2e-4), (k, float((a - b).abs().max()))


@pytest.mark.parametrize("name", ["plain_log_exp", "digamma_softplus", "mse_relu", "soft_label", "focal", "ghm_momentum", "ghm_plain",
                                  "ibloss", "ibm", "ibm_digamma_mean"])
def test_evidence_loss_branches_match_reference_golden(name, golden_dir):
    """Every EvidenceLoss branch (cls_loss.py:212-285): the masked / sync-free formulation of the product,
    fed the real rows interleaved with padding rows, and the oracle's gather formulation, both against values the
    reference produced (oracle/make_golden.py --edl).  Called repeatedly so the GHM / IBM state evolves."""
    import numpy as np
    from opental_b200.multisegment_loss import EvidenceLoss
    variants, inputs = O.EDL_VARIANTS, O.edl_inputs
    cfg, epochs = variants[name]
    gold = np.load(os.path.join(golden_dir, "edl_variants.npz"))
    K = 15
    size_average = name.endswith("_mean")
    crit = EvidenceLoss(K, dict(cfg), size_average=size_average)
    st = O.EdlVariantState(cfg.get("num_bins", 50))
    for call, epoch in enumerate(epochs):
        logit, target = inputs(name, call, K)
        want, want_g = float(gold[f"{name}.{call}.loss"]), torch.from_numpy(gold[f"{name}.{call}.grad"])
        # oracle
        st.epoch = epoch
        lo = logit.clone().requires_grad_(True)
        loss_o = O.evidence_loss_variant(lo, target, st, K, cfg, size_average=size_average)
        loss_o.backward()
        assert abs(float(loss_o) - want) <= 2e-5 * max(1.0, abs(want))
        assert float((lo.grad - want_g).abs().max()) <= 2e-5 * float(want_g.abs().max())
        # product: real rows at even positions, junk rows (masked out) at odd positions
        M = logit.shape[0]
        padded = torch.zeros(2 * M, K)
        padded[0::2] = logit
        padded[1::2] = 7.0 * torch.randn(M, K, generator=torch.Generator().manual_seed(call))
        tgt = torch.full((2 * M,), -1, dtype=torch.long)     # as MultiSegmentLoss passes them: conf_t - 1 = -1 on padding
        tgt[0::2] = target
        mask = torch.zeros(2 * M, dtype=torch.bool)
        mask[0::2] = True
        padded.requires_grad_(True)
        crit.epoch = epoch
        loss_p = crit(padded, tgt, mask)
        loss_p.backward()
        assert abs(float(loss_p) - want) <= 2e-5 * max(1.0, abs(want)), (call, float(loss_p), want)
        assert float((padded.grad[0::2] - want_g).abs().max()) <= 2e-5 * float(want_g.abs().max())
        assert float(padded.grad[1::2].abs().max()) == 0.0
    if cfg.get("with_ibm"):
        assert torch.allclose(crit.weight_accum, torch.from_numpy(gold[f"{name}.weight_accum"]), atol=1e-6)
    if cfg.get("with_ghm") and cfg.get("momentum", 0) > 0:
        assert np.allclose(crit.acc_sum.numpy(), gold[f"{name}.acc_sum"], rtol=1e-12)
        assert np.allclose(st.acc_sum, gold[f"{name}.acc_sum"], rtol=1e-12)


@pytest.mark.parametrize("name", ["mse_relu", "focal", "ghm_momentum", "ibloss", "digamma_softplus"])
def test_multisegment_loss_with_ablation_branches_matches_reference_golden(name, golden_dir):
    """The whole MultiSegmentLoss with an ablation edl_config (masked formulation; the fused kernel declines these)."""
    import numpy as np
    gold = np.load(os.path.join(golden_dir, "edl_variants.npz"))
    cfg, _ = O.EDL_VARIANTS[name]
    crit = MultiSegmentLoss(15, 0.5, 1.0, cls_loss_type="edl", edl_config=dict(cfg, iou_aware=True), os_head=True,
                            act_config=dict(weight=0.1, margin=1.0))
    crit.cls_loss.epoch = 11
    keys = ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act")
    for it in range(2):
        out = {k: v.requires_grad_(True) for k, v in O.fake_head_outputs(3, 5 + it).items()}
        assert not crit._fused_ok(out["loc"])
        out["priors"] = torch.cat(O.level_priors(O.OracleConfig()), 0)
        losses = crit(out, [O.synthetic_targets(i, num_classes=15) for i in range(3)])
        want = gold[f"{name}.msl.{it}.losses"]
        for a, b in zip(losses, want):
            assert abs(float(a) - b) <= 1e-5 * max(1.0, abs(b))
        grads = torch.autograd.grad(sum(w * l for w, l in zip((1, 10, 1, 10, 1, 1, 1), losses)), [out[k] for k in keys])
        for k, g in zip(keys, grads):
            w = torch.from_numpy(gold[f"{name}.msl.{it}.grad.{k}"])
            assert float((g - w).abs().max()) <= 2e-5 * max(float(w.abs().max()), 1e-6), k


@pytest.mark.parametrize("tag", ["init", "biased"])
def test_closed_set_focal_loss_matches_reference_golden(tag, golden_dir):
    """BASELINE configs[0] (configs/thumos14.yaml: focal, 21 classes incl. background, no actionness heads): the
    reference's head outputs (golden) through the product's MultiSegmentLoss and the oracle's restatement -> the
    reference's loss values."""
    import json
    import numpy as np
    arrays = np.load(os.path.join(golden_dir, "model_thumos_closed.npz"))
    with open(os.path.join(golden_dir, "model_thumos_closed.json")) as fh:
        want = json.load(fh)[tag]["losses"]
    cfg = O.OracleConfig(num_classes=21, os_head=False, use_edl=False)
    out = {k: torch.from_numpy(arrays[f"{tag}.{k}"]) for k in ("loc", "conf", "prop_loc", "prop_conf", "center")}
    out.update(priors=torch.cat(O.level_priors(cfg), 0), act=None, prop_act=None)
    targets = [O.synthetic_targets(0, num_classes=20)]
    got_o = O.multisegment_loss_closed(out, targets, cfg)
    got_p = MultiSegmentLoss(21, 0.5, 1.0, cls_loss_type="focal")(out, targets)
    assert got_p[5] is None and got_p[6] is None
    for a, b, c in zip(got_o, got_p[:5], want):
        assert abs(float(a) - c) <= 2e-5 * max(1.0, abs(c)) and abs(float(b) - c) <= 2e-5 * max(1.0, abs(c)), (float(a), float(b), c)


def test_pad_targets_fixed_slots():
    from opental_b200.multisegment_loss import pad_targets
    t = [torch.tensor([[0.1, 0.2, 3.0]]), torch.tensor([[0.3, 0.5, 1.0], [0.6, 0.9, 2.0]])]
    p, v = pad_targets(t)
    assert tuple(p.shape) == (2, 2, 3) and v.tolist() == [[True, False], [True, True]]
    p8, v8 = pad_targets(t, slots=8)
    assert tuple(p8.shape) == (2, 8, 3) and int(v8.sum()) == 3 and torch.equal(p8[:, :2], p)
    assert pad_targets((p8, v8), slots=4)[0] is p8                    # an already padded pair passes through
    with pytest.raises(ValueError):
        pad_targets(t, slots=1)
    # the loss does not depend on the number of (invalid) padding slots
    out = {k: v_ for k, v_ in O.fake_head_outputs(2, 3).items()}
    out["priors"] = torch.cat(O.level_priors(O.OracleConfig()), 0)
    crit = MultiSegmentLoss(15, 0.5, 1.0, cls_loss_type="edl", edl_config=OPENTAL_EDL_CONFIG, os_head=True, act_config=dict(weight=0.1, margin=1.0))
    crit.cls_loss.epoch = 1
    a = crit(out, pad_targets(t))
    b = crit(out, pad_targets(t, slots=8))
    assert all(torch.equal(x, y) for x, y in zip(a, b))
