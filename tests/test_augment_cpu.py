"""CPU: the SSL cut-paste augmentation as a frame map (opental_b200/augment.py) against fixtures produced by the
reference's own `THUMOS_Dataset.augment` (thumos_dataset.py:187-237) with the same `random` seed: the reference's
augmented clip, built from a clip whose pixel value is its frame index, IS the frame map (oracle/make_golden.py --augment)."""
import json
import os
import random

import numpy as np
import pytest

from opental_b200 import augment as A


@pytest.fixture(scope="module")
def cases(golden_dir):
    with open(os.path.join(golden_dir, "augment_cases.json")) as fh:
        return json.load(fh)


def test_frame_map_equals_reference_augmented_clip(cases):
    assert len(cases) == 60 and sum(c["flag"] for c in cases) >= 40 and any(not c["flag"] for c in cases)
    for c in cases:
        random.seed(c["seed"])
        fmap, annos, flag = A.cut_paste([list(a) for a in c["annos"]], c["th"], 256, 1)
        assert flag == c["flag"], c["seed"]
        assert fmap.dtype == np.int32 and fmap.tolist() == c["frame_map"], c["seed"]
        assert [list(map(float, a)) for a in annos] == c["new_annos"], c["seed"]


def test_frame_map_properties(cases):
    for c in cases:
        fmap = np.array(c["frame_map"])
        assert fmap.min() >= 0 and fmap.max() < 256
        if not c["flag"]:
            assert (fmap == np.arange(256)).all()            # no augmentation: identity, annotations unchanged
            continue
        th = c["th"]
        (a0, a1), (p0, p1), (n0, n1) = c["new_annos"]
        # the pasted background snippet is th consecutive source frames lying outside every annotated action
        lo = int(n0) - 1 if n1 - n0 == th - 2 else None
        assert lo is not None
        snippet = fmap[lo:lo + th]
        assert (np.diff(snippet) == 1).all()
        for s, e, _ in c["annos"]:
            assert snippet[-1] <= s or snippet[0] >= e      # gaps include the boundary frames (get_bg)
        # anchor and positive are the two halves of one action, th frames apart after the paste
        assert p0 - a1 == th
        # everything outside the rewritten span is untouched
        changed = np.nonzero(fmap != np.arange(256))[0]
        assert changed.size > 0 and changed.max() - changed.min() + 1 <= 256


def test_own_rng_is_reproducible_and_local():
    annos, th = [[30, 120, 3], [170, 230, 7]], 8
    a = A.cut_paste(annos, th, 256, 1, rng=random.Random(5))
    state = random.getstate()
    b = A.cut_paste(annos, th, 256, 1, rng=random.Random(5))
    assert random.getstate() == state                        # a private generator leaves the global one alone
    assert a[0].tolist() == b[0].tolist() and a[1] == b[1] and a[2] == b[2] is True


def test_no_long_action_or_no_background_means_no_augmentation():
    assert A.cut_paste([[10, 20, 1]], 8)[2] is False                        # action shorter than 2*th
    assert A.cut_paste([[1, 254, 1]], 8)[2] is False                        # no background gap longer than th
