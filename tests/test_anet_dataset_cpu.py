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
