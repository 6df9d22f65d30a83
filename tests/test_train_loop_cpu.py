"""CPU: the training driver's bookkeeping (opental_b200/train_loop.py; semantics of AFSD/thumos14/train.py:203-300, 370-380)
with a recording stand-in for the Trainer (the real one needs a GPU: tests/test_model_gpu.py, test_checkpoint_gpu.py)."""
import types

import torch

from opental_b200 import train_loop


class FakeTrainer:
    def __init__(self, ibm_start=3):
        self.criterion = types.SimpleNamespace(cls_loss=types.SimpleNamespace(epoch=0, total_epoch=25, ibm_start=ibm_start))
        self.calls, self.captures, self.saved, self.resumed = [], [], [], None
        self._graph = None
        self.cache = set()

    def _flag(self):
        return self.criterion.cls_loss.epoch >= self.criterion.cls_loss.ibm_start

    def select_graph(self, ssl, targets=None):
        key = (bool(ssl), self._flag())
        if key in self.cache:
            self._graph = key
            return True
        return False

    def capture(self, clips, targets, scores, **kw):
        self._graph = (bool(kw), self._flag())
        self.cache.add(self._graph)
        self.captures.append((self.criterion.cls_loss.epoch, sorted(kw)))

    def step(self, clips, targets, scores, **kw):
        assert self._graph is None or self._graph == (bool(kw), self._flag())
        self.calls.append((self.criterion.cls_loss.epoch, sorted(kw)))
        v = float(len(self.calls))
        t = torch.tensor
        return t(v), (t(1.0), t(2.0), t(3.0), t(4.0), t(5.0), t(6.0), t(7.0)), t(0.5), t(0.25)

    def grad_norm(self):
        return torch.tensor(2.0)

    def save_checkpoint(self, epoch, a, b):
        self.saved.append((epoch, a, b))

    def resume(self, epoch, a, b):
        self.resumed = epoch
        return epoch + 1


def batch(flag, with_map=True):
    b = dict(clips=torch.zeros(1), targets=[torch.zeros(1, 3)], scores=torch.zeros(1, 2, 4), flags=[flag, True],
             ssl_targets=[torch.zeros(3, 2)])
    if with_map:
        b["ssl_frame_map"] = torch.zeros(1, 4, dtype=torch.int32)
    return b


def test_epoch_means_ssl_flag_and_recapture():
    tr = FakeTrainer()
    tr.criterion.cls_loss.epoch = 1
    m = train_loop.run_one_epoch(tr, [batch(True), batch(False), batch(True), batch(True, with_map=False)], 1)
    # SSL pass only when the FIRST sample's flag is set (train.py:237) and an augmented clip / frame map is present
    assert [c[1] for c in tr.calls] == [["ssl_frame_map", "ssl_targets"], [], ["ssl_frame_map", "ssl_targets"], []]
    assert m["steps"] == 4 and m["ssl_steps"] == 2
    assert len(tr.captures) == 2                                   # one per flavour even when the flavours alternate
    assert m["cost"] == 2.5 and m["loc"] == 1.0 and m["prop_conf"] == 4.0 and m["start"] == 0.5 and m["end"] == 0.25
    assert m["act"] == 6.0 and m["prop_act"] == 7.0 and m["grad_norm"] == 2.0
    line = train_loop.summary_line(1, m)
    assert line.startswith("Epoch-1 Train Loss: Total - 2.50000, loc - 1.00000, conf - 2.00000")


def test_fit_sets_epochs_recaptures_on_ibm_switch_and_checkpoints_after_epoch_10(tmp_path):
    tr = FakeTrainer(ibm_start=3)
    logs = []
    hist = train_loop.fit(tr, lambda e: [batch(False), batch(False)], max_epoch=12, checkpoint_path=str(tmp_path / "c"),
                          train_state_path=str(tmp_path / "s"), log=logs.append)
    assert [h["epoch"] for h in hist] == list(range(1, 13)) and len(logs) == 12
    assert tr.criterion.cls_loss.total_epoch == 12
    assert [c[0] for c in tr.calls] == [e for e in range(1, 13) for _ in range(2)]
    assert [c[0] for c in tr.captures] == [1, 3]                    # first step, then when the IBM switch flips at epoch 3
    assert [s[0] for s in tr.saved] == [11, 12]                     # `if training and epoch > 10: save_model` (train.py:291-293)


def test_fit_resumes_from_the_reference_layout():
    tr = FakeTrainer()
    hist = train_loop.fit(tr, lambda e: [batch(False)], max_epoch=6, resume=4, checkpoint_path="c", train_state_path="s",
                          use_graph=False, log=lambda s: None)
    assert tr.resumed == 4 and [h["epoch"] for h in hist] == [5, 6] and not tr.captures


def test_closed_set_losses_without_actionness_terms():
    tr = FakeTrainer()
    tr.step = lambda *a, **k: (torch.tensor(1.0), (torch.tensor(1.0),) * 5 + (None, None), torch.tensor(0.0), torch.tensor(0.0))
    m = train_loop.run_one_epoch(tr, [batch(False)], 1, use_graph=False)
    assert m["act"] == 0.0 and m["prop_act"] == 0.0


def _fit_worker(rank, world, port, out_dir):
    import os
    import sys
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from opental_b200 import train_loop as tl
    from opental_b200.engine import shard_indices
    tr = FakeTrainer()
    logs = []
    windows_ = list(range(10))                                       # 10 windows, batch 1 per rank: 5 steps per epoch per rank

    def make_batches(epoch):
        for i in shard_indices(len(windows_), rank, world):
            yield batch(False)

    hist = tl.fit(tr, make_batches, max_epoch=11, checkpoint_path=os.path.join(out_dir, "c"), train_state_path=os.path.join(out_dir, "s"),
                  use_graph=False, log=logs.append)
    torch.save(dict(saved=tr.saved, logs=len(logs), steps=[h["steps"] for h in hist]), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_fit_world_size_2_rank0_writes_checkpoints_and_logs(tmp_path):
    """gloo, world size 2: every rank runs its shard of the epoch, only rank 0 logs and writes checkpoints."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_fit_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in range(2))
    assert r0["steps"] == r1["steps"] == [5] * 11
    assert [s[0] for s in r0["saved"]] == [11] and r1["saved"] == []
    assert r0["logs"] == 11 and r1["logs"] == 0


def test_trainer_graph_cache_bookkeeping():
    """Trainer.select_graph / graph_matches / _release_for_capture without a GPU: the methods only move (graph, static buffers,
    outputs, IBM flag) tuples around, so stand-in objects are enough.  One graph per flavour (with / without the SSL pass);
    a flavour's graph is dropped when it is re-captured; IBM switch and target slots gate the reuse."""
    import types
    from opental_b200.engine import Trainer
    tr = Trainer.__new__(Trainer)
    tr._graph, tr._graph_ssl, tr._graph_cache, tr._static, tr._graph_out, tr.target_slots = None, False, {}, None, None, 8
    tr._graph_updates = True                            # does the current graph contain the exchange + Adam (travels with the graph)
    cls = types.SimpleNamespace(with_ibm=True, epoch=1, ibm_start=3)
    tr.criterion = types.SimpleNamespace(cls_loss=cls)

    def fake_capture(ssl, slots=8):                      # what capture() leaves behind
        tr._release_for_capture(ssl)
        tr._graph = object()
        tr._static = [None, torch.zeros(2, slots, 3)] + ([None] * (4 if ssl else 2))
        tr._graph_out, tr._graph_epoch_flag, tr._graph_ssl = ("out", ssl), tr._ibm_flag(), ssl
        return tr._graph

    tg2 = [torch.zeros(2, 3), torch.zeros(1, 3)]
    tg9 = [torch.zeros(9, 3), torch.zeros(1, 3)]
    assert not tr.select_graph(False, tg2)               # nothing captured yet
    g_plain = fake_capture(False)
    assert tr.select_graph(False, tg2) and tr._graph is g_plain
    assert not tr.select_graph(True, tg2)                # other flavour: capture needed ...
    g_ssl = fake_capture(True)
    assert tr._graph is g_ssl and tr._graph_cache[False][0] is g_plain          # ... and the plain graph is kept
    assert tr.select_graph(False, tg2) and tr._graph is g_plain and tr._graph_out == ("out", False)
    assert tr._graph_cache[True][0] is g_ssl
    assert tr.select_graph(True, tg2) and tr._graph is g_ssl                   # alternating batches: no re-capture
    assert not tr.select_graph(True, tg9)                # 9 segments > 8 slots
    assert tr.graph_matches(True, tg2) and not tr.graph_matches(True, tg9) and tr.graph_matches(True)
    g_ssl9 = fake_capture(True, slots=9)
    assert tr._graph is g_ssl9 and True not in tr._graph_cache                 # the 8-slot SSL graph is gone, not cached
    assert tr.select_graph(True, tg9) and tr.select_graph(False, tg2) and tr._graph is g_plain
    cls.epoch = 3                                        # the IBM switch flips: neither graph fits any more
    assert not tr.select_graph(False, tg2) and not tr.select_graph(True, tg2)
    g_plain_ibm = fake_capture(False)
    assert tr._graph is g_plain_ibm and tr._graph_cache.get(False) is None     # stale plain graph released
    assert tr._graph_cache[True][0] is g_ssl9 and not tr.select_graph(True, tg9)   # cached SSL graph has the old switch
    # an already padded (tensor, mask) pair carries its own geometry: compared by step(), accepted here
    assert tr.select_graph(False, (torch.zeros(2, 8, 3), torch.zeros(2, 8, dtype=torch.bool)))
