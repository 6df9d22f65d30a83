"""CPU: the threaded batch prefetcher (opental_b200/loader.py) — same batches as the in-thread `dataset.epoch_batches`,
independent of the number of threads; bounded read-ahead; errors of a loader thread reach the consumer."""
import threading
import time

import numpy as np
import pytest
import torch

from opental_b200 import dataset as D
from opental_b200.loader import Prefetcher, step_indices, step_rng


class ToyWindows:
    """Stand-in dataset with the `sample(idx, rng)` contract of ThumosWindows: tiny frames whose bytes encode the window index."""

    def __init__(self, n=23, T=16, slow=0.0, fail_at=None):
        self.training_list = [dict(annos=[[1, 5, 1]] * (1 + i % 3)) for i in range(n)]
        self.T, self.slow, self.fail_at = T, slow, fail_at
        self.calls, self.lock = [], threading.Lock()

    def __len__(self):
        return len(self.training_list)

    def sample(self, idx, rng):
        if self.fail_at == idx:
            raise ValueError(f"window {idx} is broken")
        if self.slow:
            time.sleep(self.slow)
        with self.lock:
            self.calls.append(idx)
        n = 1 + idx % 3
        return dict(frames=np.full((self.T, 4, 4, 3), idx, dtype=np.uint8), crop=(rng.randint(0, 16), rng.randint(0, 16), int(rng.random() < 0.5)),
                    target=np.full((n, 3), idx, dtype=np.float32), scores=np.full((2, self.T), idx, dtype=np.float32),
                    frame_map=np.arange(self.T, dtype=np.int32)[::-1].copy() if idx % 2 else np.arange(self.T, dtype=np.int32),
                    ssl_target=np.full((3, 2), idx, dtype=np.float32), flag=bool(idx % 2))


def as_lists(b):
    tg = b["targets"]
    if isinstance(tg, tuple):
        tg = [tg[0][i][tg[1][i]] for i in range(tg[0].shape[0])]
    return dict(clips=b["clips"].clone(), targets=[t.clone() for t in tg], scores=b["scores"].clone(), flags=list(b["flags"]),
                fmap=b["ssl_frame_map"].clone(), crop=b["crop_offsets"].clone())


@pytest.mark.parametrize("workers", [1, 4])
def test_prefetcher_equals_in_thread_batches(workers):
    ds = ToyWindows()
    ref = [as_lists(b) for b in D.epoch_batches(ds, 3, epoch=2, rank=1, world=2, seed=5)]
    got = [as_lists(b) for b in Prefetcher(ds, 3, 2, rank=1, world=2, seed=5, workers=workers, depth=2)]
    assert len(ref) == len(got) == 23 // 6
    for a, b in zip(ref, got):
        assert torch.equal(a["clips"], b["clips"]) and torch.equal(a["scores"], b["scores"]) and a["flags"] == b["flags"]
        assert torch.equal(a["fmap"], b["fmap"]) and torch.equal(a["crop"], b["crop"])
        assert len(a["targets"]) == len(b["targets"]) and all(torch.equal(x, y) for x, y in zip(a["targets"], b["targets"]))


def test_targets_are_padded_to_a_fixed_geometry_and_ssl_targets_are_placeholders_when_not_augmented():
    ds = ToyWindows()
    pf = Prefetcher(ds, 2, 0, workers=2)
    assert pf.target_slots == 8 and len(pf) == 11
    for b in pf:
        tgt, valid = b["targets"]
        assert tuple(tgt.shape) == (2, 8, 3) and tuple(valid.shape) == (2, 8) and valid.dtype == torch.bool
        for i in range(2):
            idx = int(b["clips"][i, 0, 0, 0, 0])
            assert int(valid[i].sum()) == 1 + idx % 3 and (tgt[i][valid[i]] == idx).all() and (tgt[i][~valid[i]] == 0).all()
            want = torch.full((3, 2), float(idx)) if idx % 2 else torch.tensor([[0.0, 1.0], [1.0, 2.0], [2.0, 3.0]])
            assert torch.equal(b["ssl_targets"][i], want)
    many = ToyWindows()
    many.training_list[3]["annos"] = [[1, 2, 1]] * 11
    assert Prefetcher(many, 2, 0).target_slots == 11


def test_read_ahead_is_bounded_by_the_ring():
    ds = ToyWindows(n=40)
    pf = Prefetcher(ds, 2, 0, workers=4, depth=2, shuffle=False)          # ring of 3 host slots
    it = iter(pf)
    first = next(it)
    time.sleep(0.3)                                                       # the workers run ahead as far as they may
    with ds.lock:
        built = len(ds.calls) // 2
    assert first["clips"][0, 0, 0, 0, 0] == 0
    assert built <= 1 + 3, built                                          # step 0 handed out + at most `ring` steps in the slots
    rest = list(it)
    assert len(rest) == 19 and int(rest[-1]["clips"][1, 0, 0, 0, 0]) == 39


def test_crop_offsets_buffer_is_refreshed_per_batch_and_shards_are_disjoint():
    ds = ToyWindows()
    off = torch.zeros(3, 3, dtype=torch.int32)
    seen = []
    for rank in range(2):
        for b in Prefetcher(ds, 3, 1, rank=rank, world=2, crop_offsets=off, workers=2):
            assert torch.equal(off, b["crop_offsets"])
            seen += b["clips"][:, 0, 0, 0, 0].tolist()
    assert len(seen) == len(set(seen)) == 18
    idx = step_indices(23, 3, 1, rank=0, world=2)
    assert len(idx) == 3 and all(len(s) == 3 for s in idx)
    assert step_rng(0, 1, 0, 0).random() != step_rng(0, 1, 0, 1).random() != step_rng(0, 1, 1, 0).random()


def test_loader_thread_errors_reach_the_consumer():
    ds = ToyWindows(fail_at=7)
    with pytest.raises(RuntimeError, match="loader thread") as e:
        list(Prefetcher(ds, 2, 0, workers=2, shuffle=False))
    assert isinstance(e.value.__cause__, ValueError) and "window 7" in str(e.value.__cause__)
    # a sample with more segments than slots is reported the same way
    ds = ToyWindows()
    with pytest.raises(RuntimeError):
        list(Prefetcher(ds, 2, 0, target_slots=2, workers=1))
