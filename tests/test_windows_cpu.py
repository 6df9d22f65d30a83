"""CPU: the training-set window index and the start / end score maps (opental_b200/windows.py) against fixtures produced by
the reference's own `split_videos` (thumos_dataset.py:69-130; oracle/make_golden.py --windows)."""
import json
import os

import numpy as np
import pytest
import torch

import opental_oracle as O
from opental_b200 import engine, windows


@pytest.fixture(scope="module")
def cases(golden_dir):
    with open(os.path.join(golden_dir, "window_cases.json")) as fh:
        return json.load(fh)


def test_split_videos_matches_reference(cases):
    assert len(cases) == 4
    for c in cases:
        tl, th = windows.split_videos(c["infos"], c["annos"], c["clip_length"], c["stride"])
        assert th == c["th"]
        assert len(tl) == len(c["windows"]) > 0
        for got, want in zip(tl, c["windows"]):
            assert got["video_name"] == want["video_name"] and got["offset"] == want["offset"]
            assert [list(map(float, a)) for a in got["annos"]] == want["annos"]
            assert np.nonzero(got["start"])[0].tolist() == want["start"] and np.nonzero(got["end"])[0].tolist() == want["end"]
            assert got["start"].shape == (c["clip_length"],) and set(np.unique(got["start"])) <= {0.0, 1.0}


def test_window_invariants(cases):
    for c in cases:
        L = c["clip_length"]
        for w in c["windows"]:
            count = c["infos"][w["video_name"]]["sample_count"]
            assert 0 <= w["offset"] <= max(count - L, 0)                    # windows never run past the video
            assert all(1 <= a[0] < a[1] <= L for a in w["annos"])           # kept segments clipped to the window
            assert w["start"] and w["end"]                                  # a kept window has a complete segment


def test_synthetic_scores_follow_the_loader_rule():
    """engine.synthetic_scores (product side, through windows.boundary_score_maps) == the oracle's restatement."""
    for i in range(12):
        for frames in (128, 256, 768):
            t = engine.synthetic_targets(i)
            assert torch.equal(engine.synthetic_scores(t, frames), O.synthetic_scores(t, frames))
    # band width d = max(len/10, 2): a 100-frame action gets 10-frame bands centred on its boundaries
    s = engine.synthetic_scores(torch.tensor([[50 / 256, 150 / 256, 1.0]]), 256)
    assert s[0].nonzero().flatten().tolist() == list(range(45, 56)) and s[1].nonzero().flatten().tolist() == list(range(145, 156))


def test_annos_transform():
    assert windows.annos_transform([[64, 128, 3]], 256) == [[0.25, 0.5, 3]]
