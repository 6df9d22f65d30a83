"""CPU: the CUDA-graph mode of engine.Trainer (capture / step / select_graph: static input buffers, fixed ground-truth slots, one
graph per batch flavour, re-capture on slot overflow and on the IBM switch) driven by train_loop.run_one_epoch, with the CUDA
graph machinery replaced by an eager stand-in: "capturing" runs the body once, "replay" re-runs it on the static buffers and
writes the new results into the tensors the capture returned — what a real replay does.  The model's forward / backward is
replaced by a cheap function of the step's inputs, so every step's reported cost tells which data the replay really saw."""
import contextlib

import pytest
import torch

import abi_emu
import opental_oracle as O


class FakeStream:
    def wait_stream(self, other):
        pass


class FakeGraph:
    instances = []

    def __init__(self):
        FakeGraph.instances.append(self)
        self.body = None

    def replay(self):
        self.body()


@pytest.fixture
def trainer(monkeypatch):
    from opental_b200 import engine
    abi_emu.install(monkeypatch)
    FakeGraph.instances.clear()
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)

    @contextlib.contextmanager
    def fake_graph(graph, **kw):
        yield

    monkeypatch.setattr(torch.cuda, "graph", fake_graph)
    net, crit = engine.build_opental(device="cpu", epoch=1)
    crit.cls_loss.ibm_start = 3
    tr = engine.Trainer(net, crit, lr=0.0)
    calls = []

    def cheap_forward_backward(clips, targets, scores, ssl_clips=None, ssl_targets=None, ssl_frame_map=None):
        tgt, valid = targets                                   # a captured step always sees the padded pair
        cost = clips.float().mean() + (tgt[..., 2] * valid).sum() + scores.sum()
        if ssl_frame_map is not None:
            cost = cost + 1000.0 + ssl_frame_map.float().mean() + torch.stack(list(ssl_targets)).sum()
        calls.append((tuple(tgt.shape), ssl_frame_map is not None))
        z = torch.zeros(())
        return cost.detach(), tuple(z.clone() for _ in range(7)), z.clone(), z.clone()

    monkeypatch.setattr(tr, "forward_backward", cheap_forward_backward)
    real_capture = tr.capture

    def capture(clips, targets, scores, **kw):
        real_capture(clips, targets, scores, **kw)
        graph, static, out = tr._graph, tr._static, tr._graph_out
        c, t, v, sc = static[:4]
        extra = {}
        if len(static) > 4:
            extra = dict(ssl_targets=list(static[5].unbind(0)), ssl_frame_map=static[4])

        def body():
            new = cheap_forward_backward(c, (t, v), sc, **extra)
            out[0].copy_(new[0])

        graph.body = body

    monkeypatch.setattr(tr, "capture", capture)
    return tr, crit, calls


def batch(i, n_seg=2, ssl=False, flag=True):
    g = torch.Generator().manual_seed(i)
    b = dict(clips=torch.randint(0, 256, (1, 4, 8, 8, 3), generator=g, dtype=torch.uint8),
             targets=[torch.cat([torch.rand(n_seg, 2, generator=g), torch.full((n_seg, 1), float(i + 1))], 1)],
             scores=torch.rand(1, 2, 4, generator=g), flags=[flag])
    if ssl:
        b.update(ssl_frame_map=torch.randint(0, 4, (1, 4), generator=g, dtype=torch.int32), ssl_targets=[torch.rand(3, 2, generator=g)])
    return b


def expected(b, ssl_ran):
    cost = b["clips"].float().mean() + b["targets"][0][:, 2].sum() + b["scores"].sum()
    if ssl_ran:
        cost = cost + 1000.0 + b["ssl_frame_map"].float().mean() + torch.stack(b["ssl_targets"]).sum()
    return float(cost)


def test_graph_mode_replays_see_the_current_batch_and_capture_only_when_needed(trainer):
    from opental_b200 import train_loop
    tr, crit, calls = trainer
    seq = [batch(0), batch(1, ssl=True), batch(2), batch(3, ssl=True), batch(4, ssl=True, flag=False),     # flag False: no SSL pass
           batch(5, n_seg=9), batch(6), batch(7, ssl=True)]
    costs = []
    m = train_loop.run_one_epoch(tr, seq, epoch=1, use_graph=True, on_step=lambda it, info: costs.append((info["ssl"], float(info["cost"]))))
    assert m["steps"] == 8 and m["ssl_steps"] == 3
    ran_ssl = [True if (b.get("ssl_frame_map") is not None and b["flags"][0]) else False for b in seq]
    assert [s for s, _ in costs] == ran_ssl
    for (s, c), b in zip(costs, seq):
        assert abs(c - expected(b, s)) <= 1e-4 * max(1.0, abs(expected(b, s))), (c, expected(b, s))
    # captures: plain (8 slots), ssl (8 slots), then the 9-segment clip forces a 9-slot plain graph; the ssl graph is reused
    assert len(FakeGraph.instances) == 3
    assert tr._static[1].shape[1] == 8 and tr._graph_ssl is True                  # batch 7 selected the cached 8-slot ssl graph
    assert tr._graph_cache[False][1][1].shape[1] == 9
    # the IBM switch flips at epoch 3: every flavour is captured again, the stale graphs are dropped
    crit.cls_loss.epoch = 3
    n0 = len(FakeGraph.instances)
    train_loop.run_one_epoch(tr, [batch(8), batch(9, ssl=True), batch(10)], epoch=3, use_graph=True)
    assert len(FakeGraph.instances) == n0 + 2
    # eager mode never captures
    train_loop.run_one_epoch(tr, [batch(11)], epoch=3, use_graph=False)
    assert len(FakeGraph.instances) == n0 + 2


def test_step_rejects_a_batch_that_does_not_fit_the_captured_graph(trainer):
    tr, crit, calls = trainer
    b = batch(0)
    tr.capture(b["clips"], b["targets"], b["scores"])
    tr.step(b["clips"], b["targets"], b["scores"])
    with pytest.raises(RuntimeError, match="capture"):
        tr.step(b["clips"], batch(1, n_seg=9)["targets"], b["scores"])                      # more segments than slots
    with pytest.raises(RuntimeError, match="capture"):
        tr.step(torch.zeros(2, 4, 8, 8, 3, dtype=torch.uint8), b["targets"] * 2, b["scores"].repeat(2, 1, 1))   # other batch size
    s = batch(2, ssl=True)
    with pytest.raises(RuntimeError, match="capture"):
        tr.step(s["clips"], s["targets"], s["scores"], ssl_targets=s["ssl_targets"], ssl_frame_map=s["ssl_frame_map"])   # other flavour
