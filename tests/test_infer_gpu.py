"""GPU parity of the device-side inference post-processing (opental_b200/csrc/postproc.cu) against the oracle restatement
of decode_predictions / filtering / softnms_v2 and the golden cases produced by the reference's own functions
(tests/golden/infer_cases.npz, oracle/make_golden.py --infer)."""
import os

import numpy as np
import pytest
import torch

import opental_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "infer_cases.npz"))


def fake_out(B, seed):
    cfg = O.OracleConfig()
    g = torch.Generator().manual_seed(seed)
    P, K = 126, cfg.num_classes
    return dict(loc=torch.rand(B, P, 2, generator=g) * 40 + 1, conf=2 * torch.randn(B, P, K, generator=g),
                prop_loc=0.3 * torch.randn(B, P, 2, generator=g), prop_conf=2 * torch.randn(B, P, K, generator=g),
                center=torch.randn(B, P, 1, generator=g), priors=torch.cat(O.level_priors(cfg), 0),
                act=2 * torch.randn(B, P, 1, generator=g), prop_act=2 * torch.randn(B, P, 1, generator=g)), cfg


def test_decode_scores_matches_reference_golden(golden):
    from opental_b200 import ops
    out, cfg = fake_out(1, 77)                       # the inputs of oracle/make_golden.py infer_cases()
    dev = {k: v.cuda() for k, v in out.items()}
    seg, scores, unct, act = ops.decode_scores(dev, torch.tensor([384.0]), 256, 10.0)
    assert torch.allclose(seg[0].cpu(), torch.from_numpy(golden["seg"]), atol=1e-4, rtol=1e-5)
    assert torch.allclose(scores[0].cpu(), torch.from_numpy(golden["scores"]), atol=1e-7, rtol=1e-4)
    assert torch.allclose(unct[0].cpu(), torch.from_numpy(golden["unct"]), atol=1e-6, rtol=1e-5)
    assert torch.allclose(act[0].cpu(), torch.from_numpy(golden["act"]), atol=1e-6, rtol=1e-5)


def test_decode_scores_batch_matches_oracle():
    from opental_b200 import ops
    out, cfg = fake_out(5, 3)
    offs = torch.tensor([0.0, 128.0, 256.0, 300.0, 77.0])
    dev = {k: v.cuda() for k, v in out.items()}
    seg, scores, unct, act = ops.decode_scores(dev, offs, 256, 25.0)
    for b in range(5):
        s_o, sc_o, u_o, a_o = O.decode_predictions(out, b, float(offs[b]), 25.0, cfg)
        assert torch.allclose(seg[b].cpu(), s_o, atol=1e-4, rtol=1e-5)
        assert torch.allclose(scores[b].cpu(), sc_o, atol=1e-7, rtol=1e-4)
        assert torch.allclose(unct[b].cpu(), u_o, atol=1e-6, rtol=1e-5) and torch.allclose(act[b].cpu(), a_o, atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_softnms_matches_reference_golden(golden, name):
    from opental_b200 import ops
    cand = torch.from_numpy(golden[f"nms.{name}.cand"])
    top_k, sigma = int(golden[f"nms.{name}.cfg"][0]), float(golden[f"nms.{name}.cfg"][1])
    decayed, keep, count = ops.softnms(cand[:, :2].cuda(), cand[:, 2].unsqueeze(0).cuda(), sigma=sigma, top_k=top_k)
    mask = torch.from_numpy(golden[f"nms.{name}.mask"])
    assert torch.equal(keep[0].cpu(), mask) and int(count[0]) == int(mask.sum())
    kept = torch.from_numpy(golden[f"nms.{name}.kept"])
    assert torch.allclose(decayed[0].cpu()[mask], kept[:, 2], atol=1e-6, rtol=1e-5)


def test_softnms_many_classes_matches_oracle():
    from opental_b200 import ops
    g = torch.Generator().manual_seed(12)
    C, M = 15, 1500
    st = torch.rand(M, generator=g) * 300
    seg = torch.stack([st, st + torch.rand(M, generator=g) * 40 + 0.2], -1)
    scores = torch.rand(C, M, generator=g) ** 3
    scores[scores < 0.05] = 0.0
    decayed, keep, count = ops.softnms(seg.cuda(), scores.cuda(), sigma=0.5, top_k=200)
    for c in (0, 7, 14):
        cand = torch.cat([seg, scores[c][:, None]], -1)
        kept_o, cnt_o, mask_o = O.softnms_v2(cand, sigma=0.5, top_k=200)
        # the arg-max order is identical unless two decayed scores agree to the last bit; allow no mismatch here
        assert int(count[c]) == cnt_o and torch.equal(keep[c].cpu(), mask_o)
        assert torch.allclose(decayed[c].cpu()[mask_o], kept_o[:, 2], atol=1e-6, rtol=1e-4)


def test_detect_video_runs_end_to_end():
    from opental_b200 import engine
    from opental_b200.inference import clip_offsets, detect_video
    assert clip_offsets(700, 256, 128) == [0, 128, 256, 384, 444] and clip_offsets(100, 256, 128) == [0]
    net, _ = engine.build_opental(epoch=1)
    net.load_state_dict(O.synthetic_state_dict(O.OracleConfig(), loc_bias_shift=3.4657))
    net.eval()
    frames = torch.stack([O.synthetic_clip(i) for i in range(2)], 1).reshape(3, 512, 96, 96)[:, :450].cuda()
    res = detect_video(net, frames, sample_fps=10.0, conf_thresh=0.001, top_k=100)
    assert isinstance(res, dict)
    for cl, rows in res.items():
        # (with random weights a refined extent can be negative, i.e. end < start: the reference does not forbid it)
        assert rows.shape[1] == 5 and rows.shape[0] <= 100 and bool(torch.isfinite(rows).all())
        assert float(rows[:, :2].max()) <= 450 / 10.0 + 1e-3 and float(rows[:, :2].min()) >= 0.0
        assert bool((rows[:, 2] >= 0.001).all()) and bool((rows[:, 4] > 0.5).all())


def test_closed_set_detect_video_never_reports_the_background_class():
    """without the open-set head the reference iterates `range(1, num_classes)` (test.py:208): class 0 is the
    background and never reaches filtering / soft-NMS / the proposal list, whose name map has no key 0."""
    from opental_b200.bdnet import BDNet
    from opental_b200.inference import detect_video, to_proposal_list
    cfg = O.OracleConfig(num_classes=21, os_head=False, use_edl=False)
    net = BDNet(in_channels=3, training=False, num_classes=21, os_head=False, use_edl=False).cuda()
    net.load_state_dict(O.synthetic_state_dict(cfg, loc_bias_shift=3.4657))
    net.eval()
    frames = O.synthetic_clip(0).cuda()
    res = detect_video(net, frames, sample_fps=10.0, conf_thresh=0.001, top_k=50)
    assert res and 0 not in res and all(1 <= cl <= 20 for cl in res)
    names = {i: f"class{i}" for i in range(1, 21)}                    # get_class_index_map: keys 1..K only
    props = to_proposal_list(res, names, os_head=False, use_edl=False)
    assert props and all(p["label"] in names.values() and p["actionness"] == 0.0 and p["uncertainty"] == 0.0 for p in props)
