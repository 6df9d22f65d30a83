"""CPU: the training set in the reference's file formats (csv / txt / npy) through opental_b200.dataset, against fixtures
produced by the reference's own loader (`get_video_info`, `get_video_anno`, `THUMOS_Dataset.__getitem__`,
thumos_dataset.py:13-56,133-275; oracle/make_golden.py --dataset).  The files are regenerated from seeds by the same
function the fixture generator used; the reference's fp32 clips are pinned by CRC-32 of their bytes."""
import json
import os
import random
import zlib

import numpy as np
import pytest
import torch

import make_golden
from opental_b200 import dataset as D


@pytest.fixture(scope="module")
def cases(golden_dir):
    with open(os.path.join(golden_dir, "dataset_cases.json")) as fh:
        return json.load(fh)


def crc(t) -> int:
    a = t.contiguous().numpy() if torch.is_tensor(t) else np.ascontiguousarray(t)
    return zlib.crc32(a.tobytes())


@pytest.fixture(scope="module")
def trees(tmp_path_factory, cases):
    out = {}
    for c in cases:
        root = str(tmp_path_factory.mktemp(f"thumos_{c['seed']}"))
        out[c["seed"]] = make_golden.dataset_case_files(root, c["seed"])
    return out


def build(c, tree):
    info, anno, cls, npy = tree
    vi = D.get_video_info(info)
    va = D.get_video_anno(vi, anno, cls)
    return vi, va, D.ThumosWindows(D.load_video_data(vi, npy), vi, va, clip_length=256, crop_size=96, stride=30, training=c["training"])


def test_csv_parsing_matches_reference(cases, trees):
    for c in cases:
        vi, va, _ = build(c, trees[c["seed"]])
        assert {k: {a: float(b) for a, b in v.items()} for k, v in vi.items()} == c["video_infos"]
        assert {k: [list(map(float, a)) for a in v] for k, v in va.items()} == c["video_annos"]
        o2i, i2c = D.get_class_index_map(trees[c["seed"]][2])
        assert len(o2i) == 20 and sorted(o2i.values()) == list(range(1, 21)) and i2c[1] == "Class0"


def test_samples_match_reference_getitem(cases, trees):
    n = 0
    for c in cases:
        _, _, ds = build(c, trees[c["seed"]])
        assert len(ds) == c["n_windows"]
        for want in c["samples"]:
            s = ds.sample(want["idx"], random.Random(want["rng_seed"]))
            assert list(s["crop"]) == want["crop"] and s["flag"] == want["flag"]
            assert s["frames"].dtype == np.uint8 and s["frames"].shape == (256, 112, 112, 3)
            assert np.array_equal(s["target"], np.asarray(want["target"], dtype=np.float32))
            assert np.array_equal(s["ssl_target"], np.asarray(want["ssl_target"], dtype=np.float32)[:, :2])
            assert crc(np.asarray(s["frame_map"], dtype=np.int32)) == want["frame_map_crc"]
            assert crc(torch.from_numpy(s["scores"])) == want["scores_crc"]
            # what the ingest kernel makes of (frames, crop, mirror, frame map) == the reference's fp32 clips, bit for bit
            assert crc(D.host_clip(s["frames"], s["crop"], 96)) == want["clip_crc"]
            assert crc(D.host_clip(s["frames"], s["crop"], 96, s["frame_map"])) == want["ssl_clip_crc"]
            if not c["training"]:
                assert s["crop"] == (8, 8, 0)
            n += 1
    assert n >= 20


def test_short_video_is_zero_padded(tmp_path):
    vi = {"v": dict(fps=30.0, sample_fps=10.0, count=300, sample_count=100)}
    va = {"v": [[10.0, 60.0, 2]]}
    data = {"v": np.full((100, 112, 112, 3), 7, dtype=np.uint8)}
    ds = D.ThumosWindows(data, vi, va, training=False)
    assert len(ds) == 1
    s = ds.sample(0)
    assert s["frames"].shape[0] == 256 and (s["frames"][:100] == 7).all() and (s["frames"][100:] == 0).all()


def test_collate_and_epoch_batches_shard_across_ranks(cases, trees):
    c = cases[0]
    _, _, ds = build(c, trees[c["seed"]])
    seen = []
    for rank in range(2):
        off = torch.zeros(2, 3, dtype=torch.int32)
        n_batches = 0
        for b in D.epoch_batches(ds, 2, epoch=3, rank=rank, world=2, crop_offsets=off):   # lazily: `off` is refreshed per batch
            n_batches += 1
            assert b["clips"].dtype == torch.uint8 and tuple(b["clips"].shape) == (2, 256, 112, 112, 3)
            assert tuple(b["scores"].shape) == (2, 2, 256) and tuple(b["ssl_frame_map"].shape) == (2, 256)
            assert b["ssl_frame_map"].dtype == torch.int32 and len(b["targets"]) == 2 and len(b["flags"]) == 2
            assert all(tuple(t.shape) == (3, 2) for t in b["ssl_targets"])
            assert torch.equal(off, b["crop_offsets"])
            seen.append((rank, crc(b["clips"])))
        assert n_batches == len(ds) // 4                                      # drop_last over the global batch
    assert len({h for _, h in seen}) == len(seen)                             # ranks see disjoint windows
    again = [crc(b["clips"]) for b in D.epoch_batches(ds, 2, epoch=3, rank=0, world=2)]
    assert again == [h for r, h in seen if r == 0]                            # deterministic in (seed, epoch, rank)
    other = [crc(b["clips"]) for b in D.epoch_batches(ds, 2, epoch=4, rank=0, world=2)]
    assert other != again
