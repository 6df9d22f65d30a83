"""GPU parity for BASELINE configs[0]: the closed-set baseline `configs/thumos14.yaml` — 21 classes
incl. background, softmax focal loss, no EDL, no actionness heads — on the seed-0 clip, against goldens produced by the
reference's own code (oracle/make_golden.py --closed).  Same tolerances as tests/test_model_gpu.py: head / loss outputs
< 1e-3 relative, gradient |sum| fingerprints < 5e-2 (ReLU / arg-max discontinuities)."""
import json
import math
import os

import numpy as np
import pytest
import torch

import opental_oracle as O

pytestmark = pytest.mark.gpu


def config1_clip():
    g = torch.Generator().manual_seed(0)
    return (torch.randint(0, 256, [1, 3, 256, 96, 96], generator=g).float() / 255) * 2 - 1


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


@pytest.mark.parametrize("tag,shift", [("init", 0.0), ("biased", math.log(32.0))])
def test_closed_set_forward_focal_loss_backward_match_reference_golden(golden_dir, tag, shift):
    from opental_b200.bdnet import BDNet
    from opental_b200.multisegment_loss import MultiSegmentLoss
    from opental_b200.prop_pooling import BoundaryMaxPoolingFunction
    arrays = np.load(os.path.join(golden_dir, "model_thumos_closed.npz"))
    with open(os.path.join(golden_dir, "model_thumos_closed.json")) as fh:
        summary = json.load(fh)[tag]
    cfg = O.OracleConfig(num_classes=21, os_head=False, use_edl=False)
    net = BDNet(in_channels=3, training=False, num_classes=21, os_head=False, use_edl=False).cuda()      # train.py / test.py ctor for thumos14.yaml
    net.load_state_dict(O.synthetic_state_dict(cfg, loc_bias_shift=shift))
    net.train()
    out = net(config1_clip().cuda())
    assert out["act"] is None and out["prop_act"] is None and "unct" not in out
    assert tuple(out["conf"].shape) == (1, 126, 21) and tuple(out["start"].shape) == (1, 256, 256) and tuple(out["priors"].shape) == (126, 1)
    errs = {}
    for k in ("loc", "conf", "prop_loc", "prop_conf", "center"):
        errs[k] = rel(out[k].detach().cpu(), torch.from_numpy(arrays[f"{tag}.{k}"]))
    for k in ("start", "end", "start_loc_prop", "end_loc_prop", "start_conf_prop", "end_conf_prop"):
        errs[k] = rel(out[k].detach().cpu()[:, ::8, ::8], torch.from_numpy(arrays[f"{tag}.{k}.sample"]))
    assert max(errs.values()) < 1e-3, errs
    crit = MultiSegmentLoss(21, 0.5, 1.0, cls_loss_type="focal").cuda()
    losses = crit(out, [O.synthetic_targets(0, num_classes=20).cuda()])
    assert losses[5] is None and losses[6] is None
    for a, b in zip(losses[:5], summary["losses"]):
        assert abs(float(a) - b) <= 1e-3 * max(abs(b), 1.0), (float(a), b)
    cost = losses[0] + 10 * losses[1] + losses[2] + 10 * losses[3] + losses[4]
    assert abs(float(cost) - summary["cost"]) < 1e-3 * abs(summary["cost"])
    net.backbone.flat_parameters()[1].zero_()
    BoundaryMaxPoolingFunction.compat_tscale_bug = True        # the golden gradients come from the reference kernel
    try:
        cost.backward()
    finally:
        BoundaryMaxPoolingFunction.compat_tscale_bug = False
    params = dict(net.named_parameters())
    bad = {}
    for k, (s, a) in summary["grad_fingerprint"].items():
        g = params[k].grad
        assert g is not None, k
        if a > 0:
            e = abs(float(g.abs().sum()) - a) / a
            if e > 5e-2:
                bad[k] = e
        smp = torch.from_numpy(arrays[f"{tag}.grad.{k}"])
        got = g.detach().cpu().reshape(-1)[:: max(1, g.numel() // 64)][:64]
        if smp.abs().max() > 0 and rel(got, smp) > 0.2:
            bad[k + ":sample"] = rel(got, smp)
    assert not bad, bad
