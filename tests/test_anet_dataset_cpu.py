"""CPU: the ActivityNet training set (opental_b200/anet_dataset.py, augment variant="anet") against fixtures produced by the
reference's own `ANET_Dataset` / `augment_` (AFSD/common/anet_dataset.py:32-257; oracle/make_golden.py --anet-dataset,
--augment-anet).  Files are regenerated from seeds by the fixture generator's own function."""
import json
import os
import random
import zlib

import numpy as np
import pytest
import torch

import make_golden
from opental_b200 import anet_dataset as AD
from opental_b200 import augment as A
from opental_b200 import dataset as D


def crc(t) -> int:
    a = t.contiguous().numpy() if torch.is_tensor(t) else np.ascontiguousarray(t)
    return zlib.crc32(a.tobytes())


@pytest.fixture(scope="module")
def cases(golden_dir):
    with open(os.path.join(golden_dir, "anet_dataset_cases.json")) as fh:
        return json.load(fh)


def test_samples_match_reference_getitem(cases, tmp_path_factory):
    n = 0
    for c in cases:
        root = str(tmp_path_factory.mktemp(f"anet_{c['seed']}"))
        info, npy = make_golden.anet_dataset_case_files(root, c["seed"])
        ds = AD.AnetWindows(info, npy, 768, 96, training=c["training"])
        assert len(ds) == c["n_windows"] and ds.th == c["th"]
        for want in c["samples"]:
            s = ds.sample(want["idx"], random.Random(want["rng_seed"]))
            assert list(s["crop"]) == want["crop"] and s["flag"] == want["flag"]
            assert s["frames"].dtype == np.uint8 and s["frames"].shape == (768, 112, 112, 3)
            assert np.array_equal(s["target"], np.asarray(want["target"], dtype=np.float32))
            assert np.array_equal(s["ssl_target"], np.asarray(want["ssl_target"], dtype=np.float32)[:, :2])
            assert crc(np.asarray(s["frame_map"], dtype=np.int32)) == want["frame_map_crc"]
            assert s["scores"].shape == (3, 768) and crc(torch.from_numpy(s["scores"])) == want["scores_crc"]
            assert s["scores"].max() > 1                                   # the maps hold class ids, not 1
            k = want["frame_num"]
            assert crc(D.host_clip(s["frames"], s["crop"], 96)[:, :k]) == want["clip_crc"]
            if want["ssl_clip_crc"] is not None:
                assert crc(D.host_clip(s["frames"], s["crop"], 96, s["frame_map"])) == want["ssl_clip_crc"]
            else:                                                           # the short video: padded with 128 (reference: 127.5)
                assert k < 768 and (s["frames"][k:] == 128).all()
            n += 1
    assert n >= 7


def test_videos_without_npy_or_valid_annotation_are_skipped(tmp_path):
    info, npy = make_golden.anet_dataset_case_files(str(tmp_path), 0)
    table = AD.get_video_info(info, "training")
    tl, th = AD.split_videos(table, 768, npy)
    names = {w["video_name"] for w in tl}
    assert "v_00004" in table and "v_00004" not in names                   # no npy file
    assert "v_00003" not in table                                          # validation subset
    assert all(w["offset"] == 0 for w in tl) and set(th) == names
    # the end <= start annotation of video 2 is dropped
    assert all(a[1] > a[0] for w in tl for a in w["annos"])
    tb, _ = AD.split_videos(table, 768, npy, binary_class=True)
    assert all(a[2] in (0, 1) for w in tb for a in w["annos"])


def test_anet_cut_paste_matches_reference(golden_dir):
    with open(os.path.join(golden_dir, "augment_cases_anet.json")) as fh:
        cases = json.load(fh)
    assert len(cases) >= 100 and any(not c["flag"] for c in cases)
    for c in cases:
        random.seed(c["seed"])
        fmap, annos, flag = A.cut_paste([list(a) for a in c["annos"]], c["th"], 256, 1, variant="anet")
        assert flag == c["flag"] and fmap.tolist() == c["frame_map"], c["seed"]
        assert [list(map(float, a)) for a in annos] == c["new_annos"], c["seed"]


def test_anet_variant_accepts_actions_of_exactly_twice_the_threshold():
    # `>=` (anet_dataset.py:177) vs `>` (thumos_dataset.py:193): an action of exactly 2*th frames qualifies for ActivityNet only —
    # and then the reference's own range of cut points is empty (IndexError, :179-180), which is reproduced
    annos = [[40, 56, 1]]
    assert A.cut_paste(annos, 8, 256, 1, rng=random.Random(0))[2] is False
    with pytest.raises(IndexError):
        A.cut_paste(annos, 8, 256, 1, rng=random.Random(0), variant="anet")


def test_training_cost_reads_start_end_rows_of_the_three_row_score_maps():
    """anet/train.py:134-143, 168-190: BCE targets are scores[:, 1] / scores[:, 2] (class ids, App. D7), down-sampled by 8 for the
    proposal-level maps."""
    import torch.nn.functional as F
    from opental_b200.multisegment_loss import training_cost
    g = torch.Generator().manual_seed(0)
    B, T = 2, 64
    out = {k: torch.rand(B, T if k in ("start", "end") else T // 8, 16, generator=g)          # post-ReLU features: tanh >= 0
           for k in ("start", "end", "start_loc_prop", "end_loc_prop", "start_conf_prop", "end_conf_prop")}
    scores = torch.zeros(B, 3, T)
    scores[:, 0, 10:40] = 7.0
    scores[:, 1, 8:13] = 7.0
    scores[:, 2, 37:43] = 7.0
    losses = tuple(torch.tensor(float(i + 1)) for i in range(7))
    cost, ls, le = training_cost(out, losses, scores, score_scale=8)

    def bce(x, y):          # F.binary_cross_entropy's formula; torch 2.x rejects targets > 1 on the CPU, torch 1.9 (the reference's) did not
        p, y = torch.tanh(x).mean(-1).view(-1).double(), y.contiguous().view(-1).double()
        return (-(y * torch.log(p).clamp(min=-100) + (1 - y) * torch.log(1 - p).clamp(min=-100)).mean()).float()
    sc = F.interpolate(scores, scale_factor=1.0 / 8)
    want_ls = bce(out["start"], scores[:, 1]) + 0.1 * (bce(out["start_loc_prop"], sc[:, 1]) + bce(out["start_conf_prop"], sc[:, 1]))
    want_le = bce(out["end"], scores[:, 2]) + 0.1 * (bce(out["end_loc_prop"], sc[:, 2]) + bce(out["end_conf_prop"], sc[:, 2]))
    assert torch.allclose(ls, want_ls, rtol=1e-5) and torch.allclose(le, want_le, rtol=1e-5)
    # with {0,1} targets the formula is torch's own BCE
    z = torch.rand(4, 9, 5, generator=g)
    y01 = (torch.rand(4, 2, 9, generator=g) > 0.5).float()
    from opental_b200.multisegment_loss import calc_bce_loss
    a, b = calc_bce_loss(z, z * 0.5, y01)
    assert torch.allclose(a, F.binary_cross_entropy(torch.tanh(z).mean(-1), y01[:, 0]), rtol=1e-6)
    assert torch.allclose(b, F.binary_cross_entropy(torch.tanh(z * 0.5).mean(-1), y01[:, 1]), rtol=1e-6)
    assert torch.allclose(cost, 1 * 1.0 + 10 * 2.0 + 1 * 3.0 + 10 * 4.0 + 1 * 5.0 + want_ls + want_le + 6.0 + 7.0)
    # two-row (THUMOS14) maps are used as they are
    c2, ls2, le2 = training_cost(out, losses, scores[:, 1:].contiguous(), score_scale=8)
    assert torch.equal(ls2, ls) and torch.equal(le2, le)
