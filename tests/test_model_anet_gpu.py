"""GPU parity of the ActivityNet flavour (configs/anet_opental.yaml --open_set: 768-frame clips, 150 classes, 189 priors,
single-source pyramid, per-sample loss) against golden vectors generated from the reference's own AFSD/anet code
(tests/golden/model_anet_opental.*, oracle/make_golden.py --anet).  Same tolerances as tests/test_model_gpu.py."""
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


def test_anet_forward_loss_backward_match_reference_golden(golden_dir):
    from opental_b200.bdnet import BDNet
    from opental_b200.engine import OPENTAL_EDL_CONFIG
    from opental_b200.multisegment_loss import MultiSegmentLossANet
    from opental_b200.prop_pooling import BoundaryMaxPoolingFunction
    arrays = np.load(os.path.join(golden_dir, "model_anet_opental.npz"))
    with open(os.path.join(golden_dir, "model_anet_opental.json")) as fh:
        summary = json.load(fh)
    cfg = O.anet_config()
    net = BDNet(in_channels=3, training=True, use_edl=True, num_classes=151, os_head=True, frame_num=768, variant="anet").cuda()
    net.load_state_dict(O.synthetic_state_dict(cfg, loc_bias_shift=math.log(8.0)))
    net.train()
    x = torch.stack([O.synthetic_clip(i, frames=768) for i in range(2)]).cuda()
    targets = [O.synthetic_targets(i, num_classes=cfg.num_classes) for i in range(2)]
    targets[1] = torch.cat([targets[1], torch.tensor([[0.40, 0.44, 17.0]])])
    targets = [t.cuda() for t in targets]
    out = net(x)
    errs = {}
    for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act", "unct", "prop_unct"):
        errs[k] = rel(out[k].detach().cpu(), torch.from_numpy(arrays[f"anet.{k}"]))
    for k in ("start", "end", "start_loc_prop", "end_loc_prop", "start_conf_prop", "end_conf_prop"):
        errs[k] = rel(out[k].detach().cpu()[:, ::8, ::8], torch.from_numpy(arrays[f"anet.{k}.sample"]))
    assert max(errs.values()) < 1e-3, errs
    assert torch.equal(out["priors"].cpu(), torch.from_numpy(arrays["anet.priors"]))
    crit = MultiSegmentLossANet(cfg.num_classes, 0.5, 1.0, cls_loss_type="edl", edl_config=OPENTAL_EDL_CONFIG, os_head=True).cuda()
    keys = ("loc", "conf", "prop_loc", "prop_conf", "center", "priors", "act", "prop_act")
    for epoch in (1, 11):
        crit.cls_loss.epoch = epoch
        losses = crit([out[k] for k in keys], targets)
        for a, b in zip(losses, summary[f"anet.e{epoch}"]["losses"]):
            assert abs(float(a) - b) <= 1e-3 * max(abs(b), 1.0), (epoch, float(a), b)
    net.backbone.flat_parameters()[1].zero_()
    cost = losses[0] + 10 * losses[1] + losses[2] + 10 * losses[3] + losses[4] + losses[5] + losses[6]
    assert abs(float(cost) - summary["anet.e11"]["cost"]) < 1e-3 * abs(summary["anet.e11"]["cost"])
    BoundaryMaxPoolingFunction.compat_tscale_bug = True        # the golden gradients come from the reference kernel's arithmetic
    try:
        cost.backward()
    finally:
        BoundaryMaxPoolingFunction.compat_tscale_bug = False
    fp = summary["anet.e11"]["grad_fingerprint"]
    params = dict(net.named_parameters())
    bad = {}
    for k, (s, a) in fp.items():
        g = params[k].grad
        assert g is not None, k
        if a > 0:
            e = abs(float(g.abs().sum()) - a) / a
            if e > 5e-2:
                bad[k] = e
    assert not bad, bad


def test_anet_ssl_triplet_pass_matches_reference_golden(golden_dir):
    """SSL second pass of the ActivityNet flavour (anet/BDNet.py:453-474 + the triplet loss of anet/train.py:159-166) vs the
    reference's own run (tests/golden/model_anet_ssl.*, oracle/make_golden.py --anet-ssl)."""
    from opental_b200 import engine
    from opental_b200.prop_pooling import BoundaryMaxPoolingFunction
    arrays = np.load(os.path.join(golden_dir, "model_anet_ssl.npz"))
    with open(os.path.join(golden_dir, "model_anet_ssl.json")) as fh:
        summary = json.load(fh)
    net, _ = engine.build_opental_anet(epoch=1)
    net.load_state_dict(O.synthetic_state_dict(O.anet_config(), loc_bias_shift=math.log(8.0)))
    x = O.synthetic_clip(1, frames=768).unsqueeze(0).cuda()
    proposals = [torch.tensor(summary["proposals"]).cuda()]
    a, p, n = net(x, proposals=proposals, ssl=True)
    for name, lst in (("anchor", a), ("positive", p), ("negative", n)):
        for i, t in enumerate(lst):
            assert rel(t.detach().cpu(), torch.from_numpy(arrays[f"ssl.{name}.{i}"])) < 1e-3, (name, i)
    trip = engine.Trainer.triplet_loss(a, p, n)
    assert abs(float(trip) - summary["triplet"]) < 1e-3 * max(1.0, abs(summary["triplet"]))
    net.backbone.flat_parameters()[1].zero_()
    BoundaryMaxPoolingFunction.compat_tscale_bug = True        # golden gradients: the reference kernel's arithmetic
    try:
        trip.backward()
    finally:
        BoundaryMaxPoolingFunction.compat_tscale_bug = False
    params = dict(net.named_parameters())
    bad = {}
    for k, (s, a_) in summary["grad_fingerprint"].items():
        g = params[k].grad
        if a_ > 0:
            assert g is not None, k
            e = abs(float(g.abs().sum()) - a_) / a_
            if e > 5e-2:
                bad[k] = e
    assert not bad, bad
