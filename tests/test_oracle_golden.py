"""CPU: the oracle restatement reproduces the golden vectors generated from the reference's own code
(oracle/make_golden.py, run in the build container where /root/reference exists)."""
import json
import math
import os

import numpy as np
import pytest
import torch

import opental_oracle as O


@pytest.fixture(scope="module")
def bmp_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "bmp_cases.npz"))


@pytest.fixture(scope="module")
def model_golden(golden_dir):
    arrays = np.load(os.path.join(golden_dir, "model_thumos_opental.npz"))
    with open(os.path.join(golden_dir, "model_thumos_opental.json")) as fh:
        summary = json.load(fh)
    return arrays, summary


BMP_CASES = ["level", "frame", "ssl", "tiny", "wide"]


@pytest.mark.parametrize("name", BMP_CASES)
def test_bmp_oracle_matches_kernel_emulation(bmp_golden, name):
    g = bmp_golden
    inp = torch.from_numpy(g[f"{name}.inp"]).requires_grad_(True)
    seg = torch.from_numpy(g[f"{name}.seg"])
    gout = torch.from_numpy(g[f"{name}.gout"])
    y = O.boundary_max_pooling(inp, seg, False)
    assert torch.equal(y.detach(), torch.from_numpy(g[f"{name}.fwd"]))          # pure max: bit exact
    (gx,) = torch.autograd.grad(y, inp, gout)
    assert torch.allclose(gx, torch.from_numpy(g[f"{name}.bwd_fixed"]), atol=1e-6)
    if f"{name}.bwd_compat" in g:
        (gx,) = torch.autograd.grad(O.boundary_max_pooling(inp, seg, True), inp, gout)
        assert torch.allclose(gx, torch.from_numpy(g[f"{name}.bwd_compat"]), atol=1e-6)


def test_bmp_compat_equals_fixed_when_T_equals_K(bmp_golden):
    assert np.array_equal(bmp_golden["level.bwd_compat"], bmp_golden["level.bwd_fixed"])
    assert not np.array_equal(bmp_golden["frame.bwd_compat"], bmp_golden["frame.bwd_fixed"])
    # gradient mass is conserved by the quirk
    assert math.isclose(bmp_golden["frame.bwd_compat"].sum(), bmp_golden["frame.bwd_fixed"].sum(), rel_tol=1e-5)


def test_state_dict_spec_counts():
    cfg = O.OracleConfig()
    spec = O.model_spec(cfg)
    assert len(spec) == 446
    assert sum(math.prod(s) for _, s, _ in spec) == 44_750_436
    assert len({k for k, _, _ in spec}) == 446


def test_same_padding_rule():
    assert O.same_pad(256, 7, 2) == (2, 3)      # conv1a: 256 -> 261
    assert O.same_pad(96, 7, 2) == (2, 3)
    assert O.same_pad(24, 3, 1) == (1, 1)
    assert O.same_pad(32, 3, 2) == (0, 1)       # stride-2 k3 Unit1D pads (0,1)
    assert O.same_pad(3, 2, 2) == (0, 1)        # MaxPool3d_5a on odd extent
    assert O.same_pad(64, 1, 1) == (0, 0)


@pytest.mark.parametrize("tag,shift", [("init", 0.0), ("biased", math.log(32.0))])
def test_model_forward_loss_backward_match_reference_golden(model_golden, tag, shift):
    arrays, summary = model_golden
    cfg = O.OracleConfig()
    sd = O.synthetic_state_dict(cfg, loc_bias_shift=shift)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    x = O.synthetic_clip(0).unsqueeze(0)
    targets = [O.synthetic_targets(0, num_classes=cfg.num_classes)]
    out = O.bdnet_forward(x, sd, cfg, compat=True)
    for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act", "unct", "prop_unct"):
        ref = torch.from_numpy(arrays[f"{tag}.{k}"])
        err = float((out[k].detach() - ref).abs().max() / ref.abs().max())
        assert err < 2e-5, (k, err)
    for k in ("start", "end", "start_loc_prop", "end_loc_prop", "start_conf_prop", "end_conf_prop"):
        ref = torch.from_numpy(arrays[f"{tag}.{k}.sample"])
        got = out[k].detach()[:, ::8, ::8]
        assert float((got - ref).abs().max() / ref.abs().max()) < 2e-5, k
    for epoch in (1, 11):
        state = O.LossState(epoch=epoch)
        losses = O.multisegment_loss(out, targets, state, cfg)
        ref_losses = summary[f"{tag}.e{epoch}"]["losses"]
        for a, b in zip(losses, ref_losses):
            assert abs(float(a) - b) <= 5e-5 * max(abs(b), 1.0), (epoch, float(a), b)
        assert np.allclose(state.weight_accum.numpy(), arrays[f"{tag}.e{epoch}.weight_accum"], atol=1e-6)
    if tag == "biased":
        assert summary["biased.e1"]["losses"][2] > 0          # refined positives exist in this fixture
    # backward of the epoch-11 cost: per-tensor gradient fingerprints; gradients are discontinuous in the
    # activations (ReLU / argmax flips), so the bound is loose for the backbone (see oracle/make_golden.py)
    cost = losses[0] + 10 * losses[1] + losses[2] + 10 * losses[3] + losses[4] + losses[5] + losses[6]
    cost.backward()
    fp = summary[f"{tag}.e11"]["grad_fingerprint"]
    worst = 0.0
    for k, (s, a) in fp.items():
        g = sd[k].grad
        assert g is not None, k
        if a > 0:
            worst = max(worst, abs(float(g.abs().sum()) - a) / a)
    assert worst < 5e-2, worst
    head = [k for k in fp if k.startswith("coarse_pyramid_detection.prop_") or "conf_tower" in k]
    for k in head:
        a = fp[k][1]
        if a > 0:
            assert abs(float(sd[k].grad.abs().sum()) - a) / a < 1e-3, k


def test_backbone_endpoints_match_golden(model_golden):
    arrays, _ = model_golden
    cfg = O.OracleConfig()
    sd = O.synthetic_state_dict(cfg)
    x = O.synthetic_clip(0).unsqueeze(0)
    with torch.no_grad():
        feats = O.i3d_features(x, sd, keep=None)
    assert tuple(feats["Mixed_4f"].shape) == (1, 832, 64, 6, 6) and tuple(feats["Mixed_5c"].shape) == (1, 1024, 32, 3, 3)
    for name, f in feats.items():
        ref = torch.from_numpy(arrays[f"init.feat.{name}.sample"])
        got = f[0, ::7, ::5, ::3, ::3]
        assert float((got - ref).abs().max() / ref.abs().max()) < 2e-5, name


def test_empty_positive_case():
    """No ground truth inside the clip: every loss is finite, loc/cls terms are zero (App. C4)."""
    cfg = O.OracleConfig()
    P = 126
    g = torch.Generator().manual_seed(3)
    out = dict(loc=torch.rand(1, P, 2, generator=g) * 10 + 1, conf=torch.randn(1, P, 15, generator=g),
               prop_loc=torch.randn(1, P, 2, generator=g), prop_conf=torch.randn(1, P, 15, generator=g),
               center=torch.randn(1, P, 1, generator=g), priors=torch.cat(O.level_priors(cfg), 0),
               act=torch.randn(1, P, 1, generator=g), prop_act=torch.randn(1, P, 1, generator=g))
    targets = [torch.tensor([[1.2, 1.4, 3.0]])]     # outside [0,1]: no prior falls inside
    losses = O.multisegment_loss(out, targets, O.LossState(epoch=11), cfg)
    assert float(losses[0]) == 0 and float(losses[1]) == 0 and float(losses[2]) == 0
    assert all(math.isfinite(float(v)) for v in losses)


def test_anet_oracle_matches_reference_golden(golden_dir):
    """ActivityNet flavour: the oracle's forward (2 clips x 768 frames) and per-sample loss reproduce the golden values the
    reference's own AFSD/anet code produced (oracle/make_golden.py --anet)."""
    arrays = np.load(os.path.join(golden_dir, "model_anet_opental.npz"))
    with open(os.path.join(golden_dir, "model_anet_opental.json")) as fh:
        summary = json.load(fh)
    cfg = O.anet_config()
    assert len(O.model_spec(cfg)) == 446 - 4 + 4      # one source conv instead of two, one more stride-2 level: same count
    sd = O.synthetic_state_dict(cfg, loc_bias_shift=math.log(8.0))
    x = torch.stack([O.synthetic_clip(i, frames=768) for i in range(2)])
    targets = [O.synthetic_targets(i, num_classes=cfg.num_classes) for i in range(2)]
    targets[1] = torch.cat([targets[1], torch.tensor([[0.40, 0.44, 17.0]])])
    with torch.no_grad():
        out = O.bdnet_forward(x, sd, cfg, compat=True)
    for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act", "unct", "prop_unct", "priors"):
        ref = torch.from_numpy(arrays[f"anet.{k}"])
        assert float((out[k] - ref).abs().max() / ref.abs().max()) < 2e-5, k
    for epoch in (1, 11):
        losses = O.multisegment_loss_anet(out, targets, O.LossState(epoch=epoch), cfg)
        for a, b in zip(losses, summary[f"anet.e{epoch}"]["losses"]):
            assert abs(float(a) - b) <= 5e-5 * max(abs(b), 1.0), (epoch, float(a), b)


def test_ssl_oracle_matches_reference_golden(golden_dir):
    """SSL / triplet pass (BDNet.py:482-503, train.py:174-184): oracle vs the reference-generated golden values."""
    arrays = np.load(os.path.join(golden_dir, "model_thumos_ssl.npz"))
    with open(os.path.join(golden_dir, "model_thumos_ssl.json")) as fh:
        summary = json.load(fh)
    cfg = O.OracleConfig()
    sd = O.synthetic_state_dict(cfg, loc_bias_shift=math.log(32.0))
    x = O.synthetic_clip(1).unsqueeze(0)
    proposals = [torch.tensor(summary["proposals"])]
    with torch.no_grad():
        a, p, n = O.bdnet_forward_ssl(x, sd, cfg, proposals, compat=True)
    for name, lst in (("anchor", a), ("positive", p), ("negative", n)):
        for i, t in enumerate(lst):
            ref = torch.from_numpy(arrays[f"ssl.{name}.{i}"])
            assert float((t - ref).abs().max() / ref.abs().max()) < 2e-5, (name, i)
    assert abs(float(O.triplet_loss(a, p, n)) - summary["triplet"]) < 1e-5 * max(1.0, abs(summary["triplet"]))


def test_inference_postprocessing_oracle_matches_reference_golden(golden_dir):
    """decode_predictions / softnms_v2 restatements vs the values the reference's own functions produced."""
    g = np.load(os.path.join(golden_dir, "infer_cases.npz"))
    cfg = O.OracleConfig()
    gen = torch.Generator().manual_seed(77)
    P, K = 126, cfg.num_classes
    out = dict(loc=torch.rand(1, P, 2, generator=gen) * 40 + 1, conf=2 * torch.randn(1, P, K, generator=gen),
               prop_loc=0.3 * torch.randn(1, P, 2, generator=gen), prop_conf=2 * torch.randn(1, P, K, generator=gen),
               center=torch.randn(1, P, 1, generator=gen), priors=torch.cat(O.level_priors(cfg), 0),
               act=2 * torch.randn(1, P, 1, generator=gen), prop_act=2 * torch.randn(1, P, 1, generator=gen))
    seg, scores, unct, act = O.decode_predictions(out, 0, 384, 10.0, cfg)
    assert torch.allclose(seg, torch.from_numpy(g["seg"]), atol=1e-5) and torch.allclose(scores, torch.from_numpy(g["scores"]), atol=1e-7)
    assert torch.allclose(unct, torch.from_numpy(g["unct"]), atol=1e-6) and torch.allclose(act, torch.from_numpy(g["act"]), atol=1e-6)
    for name in "abcd":
        cand = torch.from_numpy(g[f"nms.{name}.cand"])
        kept, cnt, mask = O.softnms_v2(cand, sigma=float(g[f"nms.{name}.cfg"][1]), top_k=int(g[f"nms.{name}.cfg"][0]))
        assert torch.equal(mask, torch.from_numpy(g[f"nms.{name}.mask"])) and torch.allclose(kept, torch.from_numpy(g[f"nms.{name}.kept"]), atol=1e-7)


def test_closed_set_config1_oracle_matches_reference_golden(golden_dir):
    """BASELINE configs[0] / SURVEY §8d config 1 (`configs/thumos14.yaml`, closed set, eval forward on the seed-0 clip):
    keys / shapes of §8(a9) and values the reference produced (oracle/make_golden.py --closed)."""
    arrays = np.load(os.path.join(golden_dir, "model_thumos_closed.npz"))
    cfg = O.OracleConfig(num_classes=21, os_head=False, use_edl=False)
    spec = O.model_spec(cfg)
    assert len(spec) == 442                              # no actionness heads: 446 - 2 x (weight, bias)
    g = torch.Generator().manual_seed(0)
    x = (torch.randint(0, 256, [1, 3, 256, 96, 96], generator=g).float() / 255) * 2 - 1
    with torch.no_grad():
        out = O.bdnet_forward(x, O.synthetic_state_dict(cfg), cfg, compat=True)
    shapes = dict(loc=(1, 126, 2), conf=(1, 126, 21), prop_loc=(1, 126, 2), prop_conf=(1, 126, 21), center=(1, 126, 1),
                  priors=(126, 1), start=(1, 256, 256), end=(1, 256, 256), start_loc_prop=(1, 64, 512), end_loc_prop=(1, 64, 512),
                  start_conf_prop=(1, 64, 512), end_conf_prop=(1, 64, 512))
    for k, s in shapes.items():
        assert tuple(out[k].shape) == s, k
    assert out["act"] is None and out["prop_act"] is None and "unct" not in out
    for k in ("loc", "conf", "prop_loc", "prop_conf", "center"):
        w = torch.from_numpy(arrays[f"init.{k}"])
        assert float((out[k] - w).abs().max() / w.abs().max()) < 2e-5, k
    for k in ("start", "end", "start_loc_prop", "end_loc_prop", "start_conf_prop", "end_conf_prop"):
        w = torch.from_numpy(arrays[f"init.{k}.sample"])
        assert float((out[k][:, ::8, ::8] - w).abs().max() / w.abs().max()) < 2e-5, k


def test_batch8_loss_oracle_matches_reference_golden(golden_dir):
    """The bench configuration (batch 8, ragged targets): the oracle's MultiSegmentLoss on the reference's own B = 8 head outputs
    must give the reference's losses at epoch 1 and 11 — pins the `[P,B]`-vs-`[B,P]` IoU-calibration pairing for B > 1."""
    arrays = np.load(os.path.join(golden_dir, "model_thumos_b8.npz"))
    with open(os.path.join(golden_dir, "model_thumos_b8.json")) as fh:
        summary = json.load(fh)
    cfg = O.OracleConfig()
    out = {k: torch.from_numpy(arrays[f"b8.{k}"]) for k in ("loc", "conf", "prop_loc", "prop_conf", "center", "act", "prop_act")}
    out["priors"] = torch.cat(O.level_priors(cfg), 0)
    targets = [O.synthetic_targets(i, num_classes=15) for i in range(8)]
    targets[3] = targets[3][:1].clone()
    targets[5] = torch.cat([targets[5], torch.tensor([[0.42, 0.47, 3.0]])])
    for epoch in (1, 11):
        state = O.LossState(epoch=epoch)
        losses = O.multisegment_loss(out, targets, state, cfg)
        for a, b in zip(losses, summary[f"e{epoch}"]["losses"]):
            assert abs(float(a) - b) <= 5e-5 * max(abs(b), 1.0), (epoch, float(a), b)
        assert np.allclose(state.weight_accum.numpy(), arrays[f"b8.e{epoch}.weight_accum"], atol=1e-6)
