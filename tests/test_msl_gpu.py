"""GPU parity of the single-CTA MultiSegmentLoss kernel (opental_b200/csrc/msl.cu, through the C ABI) against the CPU
oracle restatement of the reference loss (oracle/opental_oracle.py, pinned to reference-generated golden values by
tests/test_oracle_golden.py): the 7 losses, the IBM EMA buffer and the gradient of the weighted cost w.r.t. every head
output.  Tolerance: 2e-5 relative on the losses (fp32 summation order and expf/logf rounding), 1e-4 relative / 2e-6
absolute on the gradients."""
import math

import pytest
import torch

import opental_oracle as O
from opental_b200.engine import OPENTAL_EDL_CONFIG
from opental_b200.multisegment_loss import MultiSegmentLoss

pytestmark = pytest.mark.gpu

KEYS = ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act")
W = (1.0, 10.0, 1.0, 10.0, 1.0, 1.0, 1.0)


def fake_outputs(B, seed, loc_scale=30.0, P=126):
    g = torch.Generator().manual_seed(seed)
    cfg = O.OracleConfig()
    out = dict(loc=(torch.rand(B, P, 2, generator=g) * loc_scale + 1), conf=2 * torch.randn(B, P, 15, generator=g),
               prop_loc=0.3 * torch.randn(B, P, 2, generator=g), prop_conf=2 * torch.randn(B, P, 15, generator=g),
               center=torch.randn(B, P, 1, generator=g), act=torch.randn(B, P, 1, generator=g),
               prop_act=torch.randn(B, P, 1, generator=g))
    return out, torch.cat(O.level_priors(cfg), 0), cfg


def make_crit(epoch, act_weight=0.0):
    crit = MultiSegmentLoss(15, 0.5, 1.0, cls_loss_type="edl", edl_config=OPENTAL_EDL_CONFIG, os_head=True,
                            act_config=dict(weight=act_weight, margin=1.0)).cuda()
    crit.cls_loss.epoch = epoch
    return crit


def run_both(B, epoch, act_weight, seed, targets=None, loc_scale=30.0, mutate=None):
    out, priors, cfg = fake_outputs(B, seed, loc_scale)
    if mutate:
        mutate(out)
    cfg.act_weight = act_weight
    targets = targets or [O.synthetic_targets(i, num_classes=15) for i in range(B)]
    # oracle (CPU)
    ref_in = {k: v.clone().requires_grad_(True) for k, v in out.items()}
    ref_in["priors"] = priors
    state = O.LossState(epoch=epoch)
    ref = O.multisegment_loss(ref_in, targets, state, cfg)
    g_ref = torch.autograd.grad(sum(w * l for w, l in zip(W, ref)), [ref_in[k] for k in KEYS], allow_unused=True)
    # fused CUDA kernel
    dev_in = {k: v.clone().cuda().requires_grad_(True) for k, v in out.items()}
    dev_in["priors"] = priors.cuda()
    crit = make_crit(epoch, act_weight)
    assert crit._fused_ok(dev_in["loc"])
    got = crit(dev_in, [t.cuda() for t in targets])
    g_got = torch.autograd.grad(sum(w * l for w, l in zip(W, got)), [dev_in[k] for k in KEYS], allow_unused=True)
    return ref, g_ref, state, got, g_got, crit


def check(ref, g_ref, state, got, g_got, crit):
    for i, (a, b) in enumerate(zip(got, ref)):
        assert abs(float(a) - float(b)) <= 2e-5 * max(1.0, abs(float(b))), (i, float(a), float(b))
    assert torch.allclose(crit.cls_loss.weight_accum.cpu(), state.weight_accum, atol=1e-6)
    for k, a, b in zip(KEYS, g_got, g_ref):
        b = torch.zeros_like(a.cpu()) if b is None else b
        assert torch.allclose(a.cpu(), b, atol=2e-6, rtol=1e-4), (k, float((a.cpu() - b).abs().max()))


@pytest.mark.parametrize("B,epoch,act_weight", [(1, 1, 0.0), (1, 11, 0.0), (3, 11, 0.0), (2, 11, 0.1), (8, 11, 0.0), (8, 1, 0.1),
                                                (16, 11, 0.1)])
def test_fused_loss_matches_oracle(B, epoch, act_weight):
    targets = [O.synthetic_targets(i, num_classes=15) for i in range(B)]
    if B >= 3:
        targets[1] = targets[1][:1]                       # ragged number of ground-truth segments
    check(*run_both(B, epoch, act_weight, seed=10 * B + epoch, targets=targets))


def test_fused_loss_small_extents_few_refined_positives():
    # random-init-like loc (extent ~2 frames): IoU with the GT below piou almost everywhere -> PN tiny or zero
    check(*run_both(2, 11, 0.0, seed=77, loc_scale=1.0))


def test_fused_loss_no_positive_priors():
    targets = [torch.tensor([[1.2, 1.4, 3.0]]), torch.tensor([[1.5, 1.9, 4.0]])]     # outside [0,1]: no prior inside
    ref, g_ref, state, got, g_got, crit = run_both(2, 11, 0.1, seed=3, targets=targets)
    assert float(got[0]) == 0 and float(got[1]) == 0 and float(got[2]) == 0 and float(got[4]) == 0
    check(ref, g_ref, state, got, g_got, crit)


def test_fused_loss_clamped_logits():
    # logits beyond +-10 hit the clamp of the exp evidence: zero gradient there (cls_loss.py:182-190)
    def mutate(out):
        out["conf"][:, ::3, 2] = 14.0
        out["prop_conf"][:, ::4, 5] = -13.0
    check(*run_both(2, 11, 0.0, seed=5, mutate=mutate))


def test_fused_matches_masked_torch_formulation_on_device():
    """Same inputs through the torch formulation of the same module (fused = False), all on the GPU."""
    out, priors, _ = fake_outputs(8, 123)
    targets = [O.synthetic_targets(i, num_classes=15).cuda() for i in range(8)]
    res = []
    for fused in (True, False):
        crit = make_crit(11)
        crit.fused = fused
        d = {k: v.clone().cuda().requires_grad_(True) for k, v in out.items()}
        d["priors"] = priors.cuda()
        l = crit(d, targets)
        g = torch.autograd.grad(sum(w * x for w, x in zip(W, l)), [d[k] for k in KEYS])
        res.append((l, g, crit.cls_loss.weight_accum.clone()))
    for a, b in zip(res[0][0], res[1][0]):
        assert abs(float(a) - float(b)) <= 2e-5 * max(1.0, abs(float(b)))
    for k, a, b in zip(KEYS, res[0][1], res[1][1]):
        assert torch.allclose(a, b, atol=2e-6, rtol=1e-4), k
    assert torch.allclose(res[0][2], res[1][2], atol=1e-6)


def test_fused_loss_is_deterministic_and_sync_free():
    out, priors, _ = fake_outputs(8, 9)
    targets = [O.synthetic_targets(i, num_classes=15) for i in range(8)]
    from opental_b200.multisegment_loss import pad_targets
    tp, tv = pad_targets(targets, device="cpu")
    tp, tv = tp.cuda(), tv.cuda()
    d = {k: v.clone().cuda().requires_grad_(True) for k, v in out.items()}
    d["priors"] = priors.cuda()
    vals = []
    torch.cuda.synchronize()
    for _ in range(2):
        crit = make_crit(11)
        torch.cuda.set_sync_debug_mode("error")          # any host<->device synchronisation raises
        try:
            l = crit(d, (tp, tv))
            cost = sum(w * x for w, x in zip(W, l))
            g = torch.autograd.grad(cost, [d[k] for k in KEYS])
        finally:
            torch.cuda.set_sync_debug_mode("default")
        vals.append(torch.cat([torch.stack(list(l))] + [x.reshape(-1) for x in g]))
    assert torch.equal(vals[0], vals[1])
    assert all(math.isfinite(v) for v in vals[0][:7].tolist())


def test_boundary_bce_kernel_matches_torch():
    """calc_bce_loss (train.py:152-161) fused kernel vs the torch formulation the oracle uses (oracle.boundary_bce)."""
    from opental_b200.multisegment_loss import calc_bce_loss
    g = torch.Generator().manual_seed(8)
    for B, T, C in ((2, 256, 256), (3, 64, 512), (1, 768, 256)):
        start = torch.randn(B, T, C, generator=g).relu()
        end = torch.randn(B, T, C, generator=g).relu()
        scores = (torch.rand(B, 2, T, generator=g) > 0.7).float()
        sr, er = start.clone().requires_grad_(True), end.clone().requires_grad_(True)
        ls_r, le_r = O.boundary_bce(sr, er, scores)
        (ls_r + 2 * le_r).backward()
        sd, ed = start.cuda().requires_grad_(True), end.cuda().requires_grad_(True)
        ls, le = calc_bce_loss(sd, ed, scores.cuda())
        (ls + 2 * le).backward()
        assert abs(float(ls) - float(ls_r)) < 1e-5 * max(1.0, abs(float(ls_r))) and abs(float(le) - float(le_r)) < 1e-5 * max(1.0, abs(float(le_r)))
        assert torch.allclose(sd.grad.cpu(), sr.grad, atol=1e-9, rtol=1e-4) and torch.allclose(ed.grad.cpu(), er.grad, atol=1e-9, rtol=1e-4)


# ---------------------------------------------------------------------------------------------------------------------
# the other two flavours of the kernel: ActivityNet (per-sample normalisation, level-gated matching, smooth-L1, stateless IBM
# weight with its gradient through ||z||_1) and the closed-set softmax focal loss of configs/thumos14.yaml
# ---------------------------------------------------------------------------------------------------------------------
def anet_outputs(B, seed, loc_scale=60.0):
    cfg = O.anet_config()
    P, K = sum(len(p) for p in O.level_priors(cfg)), cfg.num_classes
    g = torch.Generator().manual_seed(seed)
    out = dict(loc=(torch.rand(B, P, 2, generator=g) * loc_scale + 1), conf=2 * torch.randn(B, P, K, generator=g),
               prop_loc=0.8 * torch.randn(B, P, 2, generator=g), prop_conf=2 * torch.randn(B, P, K, generator=g),
               center=torch.randn(B, P, 1, generator=g), act=torch.randn(B, P, 1, generator=g),
               prop_act=torch.randn(B, P, 1, generator=g))
    return out, torch.cat(O.level_priors(cfg), 0), cfg


@pytest.mark.parametrize("B,epoch", [(1, 1), (1, 11), (2, 11), (3, 1), (8, 11), (16, 11)])
def test_fused_anet_loss_matches_oracle(B, epoch):
    from opental_b200.multisegment_loss import MultiSegmentLossANet
    out, priors, cfg = anet_outputs(B, seed=100 + 10 * B + epoch)
    targets = [O.synthetic_targets(i, num_classes=cfg.num_classes) for i in range(B)]
    if B >= 3:
        targets[1] = targets[1][:1]
        targets[2] = torch.tensor([[0.30, 0.34, 5.0]])               # a short action: only the fine levels' ranges admit it
    ref_in = {k: v.clone().requires_grad_(True) for k, v in out.items()}
    ref_in["priors"] = priors
    ref = O.multisegment_loss_anet(ref_in, targets, O.LossState(epoch=epoch), cfg)
    g_ref = torch.autograd.grad(sum(w * l for w, l in zip(W, ref)), [ref_in[k] for k in KEYS], allow_unused=True)
    dev_in = {k: v.clone().cuda().requires_grad_(True) for k, v in out.items()}
    crit = MultiSegmentLossANet(cfg.num_classes, 0.5, 1.0, cls_loss_type="edl", edl_config=OPENTAL_EDL_CONFIG, os_head=True).cuda()
    crit.cls_loss.epoch = epoch
    from opental_b200 import _lib
    n0 = _lib.LAUNCHES.get("otal_msl_forward", 0)
    got = crit([dev_in[k] for k in ("loc", "conf", "prop_loc", "prop_conf", "center")] + [priors.cuda(), dev_in["act"], dev_in["prop_act"]],
               [t.cuda() for t in targets])
    assert _lib.LAUNCHES.get("otal_msl_forward", 0) == n0 + 1, "the ActivityNet loss must take the fused kernel"
    g_got = torch.autograd.grad(sum(w * l for w, l in zip(W, got)), [dev_in[k] for k in KEYS], allow_unused=True)
    for i, (a, b) in enumerate(zip(got, ref)):
        assert abs(float(a) - float(b)) <= 2e-5 * max(1.0, abs(float(b))), (i, float(a), float(b))
    for k, a, b in zip(KEYS, g_got, g_ref):
        b = torch.zeros_like(a.cpu()) if b is None else b
        assert torch.allclose(a.cpu(), b, atol=2e-6, rtol=1e-4), (k, float((a.cpu() - b).abs().max()))
    # and the masked torch formulation (the fallback for B x P > 4096) agrees with the kernel
    crit.fused = False
    dev2 = {k: v.clone().cuda().requires_grad_(True) for k, v in out.items()}
    got2 = crit([dev2[k] for k in ("loc", "conf", "prop_loc", "prop_conf", "center")] + [priors.cuda(), dev2["act"], dev2["prop_act"]],
                [t.cuda() for t in targets])
    for a, b in zip(got, got2):
        assert abs(float(a) - float(b)) <= 2e-5 * max(1.0, abs(float(b)))


@pytest.mark.parametrize("B", [1, 2, 8])
def test_fused_closed_set_focal_loss_matches_oracle(B):
    cfg = O.OracleConfig(num_classes=21)
    P = 126
    g = torch.Generator().manual_seed(300 + B)
    out = dict(loc=(torch.rand(B, P, 2, generator=g) * 30 + 1), conf=2 * torch.randn(B, P, 21, generator=g),
               prop_loc=0.3 * torch.randn(B, P, 2, generator=g), prop_conf=2 * torch.randn(B, P, 21, generator=g),
               center=torch.randn(B, P, 1, generator=g))
    priors = torch.cat(O.level_priors(cfg), 0)
    targets = [O.synthetic_targets(i, num_classes=20) for i in range(B)]
    keys = KEYS[:5]
    ref_in = {k: v.clone().requires_grad_(True) for k, v in out.items()}
    ref_in["priors"] = priors
    ref = O.multisegment_loss_closed(ref_in, targets, cfg)
    g_ref = torch.autograd.grad(sum(w * l for w, l in zip(W, ref)), [ref_in[k] for k in keys])
    dev_in = {k: v.clone().cuda().requires_grad_(True) for k, v in out.items()}
    dev_in["priors"] = priors.cuda()
    crit = MultiSegmentLoss(21, 0.5, 1.0, cls_loss_type="focal").cuda()
    assert crit._fused_ok(dev_in["loc"])
    got = crit(dev_in, [t.cuda() for t in targets])
    assert got[5] is None and got[6] is None
    g_got = torch.autograd.grad(sum(w * l for w, l in zip(W, got[:5])), [dev_in[k] for k in keys])
    for i, (a, b) in enumerate(zip(got[:5], ref)):
        assert abs(float(a) - float(b)) <= 2e-5 * max(1.0, abs(float(b))), (i, float(a), float(b))
    for k, a, b in zip(keys, g_got, g_ref):
        assert torch.allclose(a.cpu(), b, atol=2e-6, rtol=1e-4), (k, float((a.cpu() - b).abs().max()))


# ---------------------------------------------------------------------------------------------------------------------
# the re-weighting ablations of configs/ablations/*.yaml inside the fused kernel: focal-EDL, GHM (fp64 per-bin EMA state), IB —
# against the REFERENCE's own losses and gradients (tests/golden/edl_variants.npz, generated by oracle/make_golden.py --edl from
# the imported reference: two consecutive calls, so that the GHM state carries over) — and the no-os_head form against the
# product's masked torch formulation (itself pinned to the reference on the CPU, tests/test_loss_cpu.py)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["focal", "ghm_momentum", "ibloss"])
def test_fused_ablation_branches_match_reference_golden(name, golden_dir):
    import os

    import numpy as np
    gold = np.load(os.path.join(golden_dir, "edl_variants.npz"))
    cfg, _ = O.EDL_VARIANTS[name]
    crit = MultiSegmentLoss(15, 0.5, 1.0, cls_loss_type="edl", edl_config=dict(cfg, iou_aware=True), os_head=True,
                            act_config=dict(weight=0.1, margin=1.0)).cuda()
    crit.cls_loss.epoch = 11
    priors = torch.cat(O.level_priors(O.OracleConfig()), 0).cuda()
    from opental_b200 import _lib
    for it in range(2):
        out = {k: v.cuda().requires_grad_(True) for k, v in O.fake_head_outputs(3, 5 + it).items()}
        assert crit._fused_ok(out["loc"])
        out["priors"] = priors
        n0 = _lib.LAUNCHES.get("otal_msl_forward", 0)
        losses = crit(out, [O.synthetic_targets(i, num_classes=15).cuda() for i in range(3)])
        assert _lib.LAUNCHES.get("otal_msl_forward", 0) == n0 + 1
        want = gold[f"{name}.msl.{it}.losses"]
        for a, b in zip(losses, want):
            assert abs(float(a) - b) <= 2e-5 * max(1.0, abs(b)), (name, it, float(a), b)
        grads = torch.autograd.grad(sum(w * l for w, l in zip(W, losses)), [out[k] for k in KEYS])
        for k, g in zip(KEYS, grads):
            w = torch.from_numpy(gold[f"{name}.msl.{it}.grad.{k}"])
            assert float((g.cpu() - w).abs().max()) <= 1e-4 * max(float(w.abs().max()), 1e-6), (name, it, k)


@pytest.mark.parametrize("edl", [dict(), dict(with_ibm=True, ibm_start=0, momentum=0.9, num_bins=50), dict(with_ghm=True, num_bins=10, momentum=0.0)])
def test_fused_loss_without_os_head_matches_the_torch_formulation(edl):
    """configs/ablations/thumos14_opental_noACT.yaml: EDL over all priors with a background class, no actionness head."""
    cfg = dict(loss_type="log", evidence="exp", iou_aware=True, **edl)
    g = torch.Generator().manual_seed(41)
    B, P, K = 3, 126, 16
    base = dict(loc=(torch.rand(B, P, 2, generator=g) * 30 + 1), conf=2 * torch.randn(B, P, K, generator=g),
                prop_loc=0.3 * torch.randn(B, P, 2, generator=g), prop_conf=2 * torch.randn(B, P, K, generator=g),
                center=torch.randn(B, P, 1, generator=g))
    priors = torch.cat(O.level_priors(O.OracleConfig()), 0).cuda()
    targets = [O.synthetic_targets(i, num_classes=15).cuda() for i in range(B)]
    keys = KEYS[:5]
    res = {}
    for fused in (True, False):
        crit = MultiSegmentLoss(K, 0.5, 1.0, cls_loss_type="edl", edl_config=cfg, os_head=False).cuda()
        crit.cls_loss.epoch = 11
        crit.fused = fused
        for it in range(2):                                  # two calls: the second sees the first one's state
            out = {k: v.clone().cuda().requires_grad_(True) for k, v in base.items()}
            out["priors"] = priors
            losses = crit(out, targets)
            assert losses[5] is None and losses[6] is None
            grads = torch.autograd.grad(sum(w * l for w, l in zip(W, losses[:5])), [out[k] for k in keys])
        res[fused] = ([float(l) for l in losses[:5]], [g_.cpu() for g_ in grads])
    for a, b in zip(res[True][0], res[False][0]):
        assert abs(a - b) <= 2e-5 * max(1.0, abs(b)), (a, b)
    for k, a, b in zip(keys, res[True][1], res[False][1]):
        assert float((a - b).abs().max()) <= 1e-4 * max(float(b.abs().max()), 1e-6), k
