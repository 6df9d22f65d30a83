"""CPU, world_size 2 over gloo: the host side of the data-parallel path — clip sharding, the flat-buffer
gradient exchange (`engine.GradReducer`, the same code that runs over NCCL on the GPUs) with the 1/world scaling folded
into the optimizer, and the `bench.py --impl reference` launch contract (rank 0 prints one JSON line, other ranks exit
0 without work).  No CUDA kernel runs here; the Adam arithmetic is restated in torch for the equivalence check."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _adam_reference(p, g, m, v, lr, b1, b2, eps, wd, grad_scale, step):
    """otal_adam_step semantics (include/opental_b200.h): g*grad_scale, L2-in-gradient weight decay, bias correction."""
    g = g * grad_scale + wd * p
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    p.sub_(lr * (m / (1 - b1 ** step)) / ((v / (1 - b2 ** step)).sqrt() + eps))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from opental_b200.engine import FlatParams, GradReducer, shard_indices
    torch.manual_seed(0)                                     # identical parameters on every rank
    params = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2, 2, 2))]
    flat = FlatParams(params)
    assert all(p.data_ptr() == flat.w.data_ptr() + 4 * o for p, o in zip(params, flat.offsets))
    red = GradReducer([flat.g], None)
    assert red.world == world and red.grad_scale == 1.0 / world
    # every rank differentiates its own shard of a 6-"clip" batch
    data = torch.arange(6 * 3, dtype=torch.float32).view(6, 3) / 10.0
    idx = list(shard_indices(6, rank, world))
    loss = sum(((params[0] @ data[i]).sum() + params[1].sum() * data[i, 0] + (params[2] ** 2).sum()) for i in idx)
    flat.g.zero_()
    loss.backward()                                          # accumulates into the flat buffer through the .grad views
    local = flat.g.clone()
    red.launch()
    red.wait()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(flat.g, sum(gathered))
    m, v = torch.zeros_like(flat.w), torch.zeros_like(flat.w)
    with torch.no_grad():
        _adam_reference(flat.w, flat.g, m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-3, red.grad_scale, 1)
    torch.save(dict(w=flat.w.clone(), idx=idx), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_exchange_and_sharding_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in range(2))
    assert sorted(r0["idx"] + r1["idx"]) == list(range(6)) and not set(r0["idx"]) & set(r1["idx"])
    assert torch.equal(r0["w"], r1["w"])                     # ranks stay bit-identical after the update
    # ... and equal the single-process update on the mean gradient of the whole batch
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2, 2, 2))]
    sys.path.insert(0, ROOT)
    from opental_b200.engine import FlatParams
    flat = FlatParams(params)
    data = torch.arange(6 * 3, dtype=torch.float32).view(6, 3) / 10.0
    loss = sum(((params[0] @ data[i]).sum() + params[1].sum() * data[i, 0] + (params[2] ** 2).sum()) for i in range(6))
    flat.g.zero_()
    loss.backward()
    m, v = torch.zeros_like(flat.w), torch.zeros_like(flat.w)
    with torch.no_grad():
        _adam_reference(flat.w, flat.g, m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-3, 0.5, 1)
    assert torch.allclose(flat.w, r0["w"], atol=1e-6)


@pytest.mark.timeout(600)
def test_bench_reference_arm_under_torchrun_world2():
    """`bench.py --impl reference` launched like the driver does for N = 2: exactly one JSON line, from rank 0."""
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"]
    pr = subprocess.run(cmd, capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert pr.returncode == 0, pr.stderr[-2000:]
    lines = [l for l in pr.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, pr.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "clips/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def _globalise_worker(rank, world, port, out_dir):
    """Each rank: the THUMOS14 loss (torch formulation, CPU) on its half of a 4-clip batch, re-weighted to batch-global
    normalisers; saves its terms and the gradients w.r.t. its own head outputs."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from opental_b200.engine import global_normaliser_factors, globalise_losses
    out, targets, crit = _globalise_case()
    sl = slice(2 * rank, 2 * rank + 2)
    mine = {k: (v[sl].detach().clone().requires_grad_(True) if k != "priors" else v) for k, v in out.items()}
    losses = crit(mine, targets[sl])
    stats = crit.last_stats
    factors = global_normaliser_factors(stats[:4].clone(), world)
    g = globalise_losses(losses, stats, factors)
    cost = g[0] + 10 * g[1] + g[2] + 10 * g[3] + g[4]
    cost.backward()
    torch.save(dict(terms=[float(t) for t in g[:5]], counts=stats[:2].tolist(), factors=factors.tolist(),
                    grads={k: v.grad.clone() for k, v in mine.items() if k != "priors" and v.grad is not None}),
               os.path.join(out_dir, f"g{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _globalise_case():
    import opental_oracle as O
    from opental_b200.multisegment_loss import MultiSegmentLoss
    g = torch.Generator().manual_seed(11)
    B, P, K = 4, 126, 15
    priors = O.make_priors() if hasattr(O, "make_priors") else None
    if priors is None:
        centers = torch.cat([(torch.arange(t) + 0.5) / t for t in (64, 32, 16, 8, 4, 2)])
        priors = centers.view(-1, 1)
    out = dict(loc=torch.rand(B, P, 2, generator=g) * 40 + 2, conf=torch.randn(B, P, K, generator=g),
               prop_loc=torch.randn(B, P, 2, generator=g) * 0.1, prop_conf=torch.randn(B, P, K, generator=g),
               center=torch.randn(B, P, 1, generator=g), act=torch.randn(B, P, 1, generator=g), prop_act=torch.randn(B, P, 1, generator=g),
               priors=priors.float())
    # very different numbers of positives per rank: clips 0, 1 carry short actions, clips 2, 3 long ones
    targets = [torch.tensor([[0.10, 0.16, 3.0]]), torch.tensor([[0.70, 0.74, 5.0]]),
               torch.tensor([[0.05, 0.60, 7.0], [0.62, 0.98, 2.0]]), torch.tensor([[0.20, 0.95, 9.0]])]
    edl = dict(evidence="exp", loss_type="log", with_ibm=False, iou_aware=False, num_bins=50, momentum=0.99, ibm_start=10)
    crit = MultiSegmentLoss(K, 0.5, 1.0, cls_loss_type="edl", edl_config=edl, os_head=True, act_config=dict(margin=1.0, weight=0))
    crit.cls_loss.epoch = 1
    crit.fused = False
    return out, targets, crit


def test_global_normalisers_reproduce_the_single_process_batch(tmp_path):
    """with the per-rank terms re-weighted by max(n_r,1) * world / max(sum n_r,1), the mean over ranks of every
    count-normalised loss term equals the term of the whole batch on one process, and each rank's input gradients are world x
    the whole-batch gradients of its clips — so the summing all-reduce + Adam's 1/world give exactly the single-process update."""
    world, port = 2, _free_port()
    mp.spawn(_globalise_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"g{k}.pt") for k in range(world)]
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    out, targets, crit = _globalise_case()
    full = {k: (v.detach().clone().requires_grad_(True) if k != "priors" else v) for k, v in out.items()}
    losses = crit(full, targets)
    cost = losses[0] + 10 * losses[1] + losses[2] + 10 * losses[3] + losses[4]
    cost.backward()
    assert r[0]["counts"] != r[1]["counts"] and abs(r[0]["factors"][0] - 1.0) > 0.2          # the ranks really are unbalanced
    for i in range(5):
        got = sum(x["terms"][i] for x in r) / world
        assert abs(got - float(losses[i])) <= 1e-5 * max(1.0, abs(float(losses[i]))), (i, got, float(losses[i]))
    for k in ("loc", "conf", "prop_loc", "prop_conf", "center"):
        for rank in range(world):
            want = full[k].grad[2 * rank: 2 * rank + 2] * world
            got = r[rank]["grads"][k]
            assert torch.allclose(got, want, rtol=1e-4, atol=1e-7), (k, rank, float((got - want).abs().max()))
    # without the re-weighting the per-rank mean is NOT the batch value (this is the documented default)
    plain = [float(crit({k: (v[2 * q: 2 * q + 2] if k != "priors" else v) for k, v in out.items()}, targets[2 * q: 2 * q + 2])[0]) for q in range(2)]
    assert abs(sum(plain) / 2 - float(losses[0])) > 1e-3


def test_live_iou_calibration_is_the_term_inside_loss_prop_c():
    """engine.live_iou_calibration == the loss's own calibration value (last_stats[4]); globalise_losses re-weights only the
    count-normalised part of loss_prop_c and leaves the calibration term (value and gradient) at weight 1."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from opental_b200.engine import globalise_losses, live_iou_calibration
    out, targets, crit = _globalise_case()
    crit.iou_aware = crit.cls_loss.iou_aware = True
    live = {k: (v.detach().clone().requires_grad_(True) if k != "priors" else v) for k, v in out.items()}
    losses = crit(live, targets)
    stats = crit.last_stats
    iouc = live_iou_calibration(crit, live, targets)
    assert float(stats[4]) > 0 and torch.allclose(iouc, stats[4], rtol=1e-6)
    factors = torch.tensor([0.5, 2.0, 1.0, 1.0])
    g = globalise_losses(losses, stats, factors, iouc)
    a_over_pn = losses[3] - iouc
    assert torch.allclose(g[3], a_over_pn * 2.0 + iouc, rtol=1e-6)
    (gr,) = torch.autograd.grad(g[3], live["prop_conf"], retain_graph=True)
    (ga,) = torch.autograd.grad(a_over_pn, live["prop_conf"], retain_graph=True)
    (gi,) = torch.autograd.grad(iouc, live["prop_conf"], retain_graph=True)
    assert torch.allclose(gr, 2.0 * ga + gi, rtol=1e-5, atol=1e-9)
    assert torch.allclose(g[0], losses[0] * 0.5) and torch.allclose(g[2], losses[2] * 2.0)
    crit.iou_aware = False
    assert live_iou_calibration(crit, live, targets) is None


def test_trainer_globalise_glue_with_the_fused_loss_stats(monkeypatch):
    """Trainer._globalise on the FUSED loss's 16-float stats vector (counts + loss_iouc at [7:12]) — the fused entry points
    emulated by the oracle's loss (tests/abi_emu.py) — in a 1-process gloo group with the world size forced to 2: every
    count-normalised term doubles (n * 2 / n), the IoU-calibration term inside loss_prop_c keeps weight 1."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import abi_emu
    from opental_b200.engine import Trainer, live_iou_calibration
    abi_emu.install(monkeypatch)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        out, targets, crit = _globalise_case()
        crit.fused = True
        crit.iou_aware = crit.cls_loss.iou_aware = True
        live = {k: (v.detach().clone().requires_grad_(True) if k != "priors" else v) for k, v in out.items()}
        losses = crit(live, targets)
        assert crit.last_stats.numel() == 16                                  # the fused path's vector
        tr = Trainer.__new__(Trainer)
        tr.criterion, tr.world, tr.pg = crit, 2, None
        g = tr._globalise(live, losses, targets)
        iouc = live_iou_calibration(crit, live, targets)
        assert torch.allclose(iouc, crit.last_stats[11], rtol=1e-5)
        for i in (0, 1, 2, 4, 5, 6):
            assert torch.allclose(g[i], 2.0 * losses[i], rtol=1e-6), i
        assert torch.allclose(g[3], 2.0 * (losses[3] - iouc) + iouc, rtol=1e-5)
    finally:
        dist.destroy_process_group()


def test_head_allreduce_hook_fires_once_after_both_passes_of_an_ssl_step():
    """a step with the SSL pass has two backbone autograd nodes; the SSL branch's node runs FIRST in the backward,
    before the main pass's head has accumulated its gradients.  The hook that launches the head buffers' all-reduce must fire
    once, when the LAST backbone backward starts — by then the head parameter's gradient holds both passes' contributions."""
    from opental_b200 import backbone as bb_mod
    net = bb_mod.I3DBackbone()
    calls = []
    head_w = torch.nn.Parameter(torch.tensor([2.0, -1.0]))

    def fake_forward_planes(x, saved):
        saved["x"] = x
        f = x.sum() * torch.ones(1, 1, 1, 1, 2)
        return {"Mixed_4f": f.clone(), "Mixed_5c": f.clone()}

    net.forward_planes = fake_forward_planes
    net.backward_planes = lambda saved, g4, g5: calls.append(("bb_bwd", None))
    net._ensure_flat = lambda dev: None
    net._anchor = torch.zeros(1, requires_grad=True)
    net.on_backward_start = lambda: calls.append(("hook", None if head_w.grad is None else head_w.grad.clone()))

    def head(feat):                       # stands for the pyramid + loss: linear in the head parameter
        return (feat["Mixed_5c"].reshape(-1)[:2] * head_w).sum()

    x_main, x_ssl = torch.full((1,), 3.0), torch.full((1,), 5.0)
    cost = head(net(x_main)) + 0.5 * head(net(x_ssl))
    assert net._pending_bwd == 2
    cost.backward()
    hooks = [c for c in calls if c[0] == "hook"]
    assert len(hooks) == 1 and [c[0] for c in calls].count("bb_bwd") == 2
    assert calls.index(hooks[0]) == 1                                  # after the first (SSL) backbone backward, before the last
    assert torch.allclose(hooks[0][1], torch.full((2,), 3.0 + 0.5 * 5.0))          # both passes' head gradients are in
    assert net._pending_bwd == 0
    # single-pass step: fires at the one and only backbone backward
    calls.clear(); head_w.grad = None
    head(net(x_main)).backward()
    assert [c[0] for c in calls] == ["hook", "bb_bwd"]


def test_launch_head_allreduce_is_idempotent_within_a_step():
    from opental_b200.engine import Trainer
    tr = Trainer.__new__(Trainer)
    launched = []
    tr.reducer = type("R", (), {"launch": lambda self, which=None: launched.append(list(which))})()
    tr.groups = [None, None, None]
    tr._n_head_buckets = 2                      # exchange buckets: the head's two buffers first, then the backbone's two parts
    tr._head_launched = tr._deep_launched = tr._warming_up = tr._defer_comm = False
    tr._early, tr.early_update = False, True    # not inside a step: the hooks launch no optimizer work
    tr._launch_head_allreduce(); tr._launch_head_allreduce()
    assert launched == [[0, 1]]
    # the deep-stage hook launches the head buckets if nobody has, then Mixed_4b..5c, once
    launched.clear(); tr._head_launched = False
    tr._on_backbone_deep_done(); tr._on_backbone_deep_done()
    assert launched == [[0, 1], [2]]


def test_early_optimizer_launches_cover_every_parameter_exactly_once(monkeypatch):
    """Trainer._early_update / _finish_step bookkeeping (no GPU): whatever subset of the hooks fired during the backward — none
    (a graph whose update follows the replay), the head hook only (one GPU), head + deep (data parallel) or deep first — the Adam
    launches of one step partition every group exactly once, the step counter advances once, and buckets are waited for before
    their ranges are updated."""
    import types
    from opental_b200 import engine
    from opental_b200.engine import Trainer
    events = []
    monkeypatch.setattr(engine.ops, "adam_step", lambda w, g, m, v, **kw: events.append(("adam", w.data_ptr(), w.numel(), kw["lr"])))

    def make(world):
        tr = Trainer.__new__(Trainer)
        tr.world, tr.device = world, torch.device("cpu")
        tr.groups = [(torch.zeros(100), torch.zeros(100)), (torch.zeros(40), torch.zeros(40)), (torch.zeros(7), torch.zeros(7))]
        tr.state = [dict(m=torch.zeros_like(w), v=torch.zeros_like(w)) for w, _ in tr.groups]
        tr._bb_split, tr._n_head_buckets = 60, 2
        tr.lr, tr.backbone_lr_scale, tr.betas, tr.eps, tr.wd = 1e-3, 0.1, (0.9, 0.999), 1e-8, 0.0
        tr._step_dev = torch.zeros(1, dtype=torch.int32)
        tr._head_launched = tr._deep_launched = tr._warming_up = tr._defer_comm = False
        tr._early, tr.early_update, tr._updated, tr._counted, tr._opt_stream = True, True, [], False, None
        tr.reducer = types.SimpleNamespace(grad_scale=1.0 / world,
                                           launch=lambda which=None: events.append(("launch", list(which))),
                                           wait=lambda which=None: events.append(("wait", None if which is None else list(which))))
        return tr

    def covered(tr):
        out = {}
        for e in events:
            if e[0] == "adam":
                gi = next(i for i, (w, _) in enumerate(tr.groups) if w.data_ptr() <= e[1] < w.data_ptr() + 4 * w.numel())
                lo = (e[1] - tr.groups[gi][0].data_ptr()) // 4
                out.setdefault(gi, []).append((lo, lo + e[2], e[3]))
        return {gi: sorted(v) for gi, v in out.items()}

    for world, hooks in ((1, ()), (1, ("head",)), (2, ("head", "deep")), (2, ("deep",)), (2, ())):
        events.clear()
        tr = make(world)
        for h in hooks:
            (tr._on_backbone_backward if h == "head" else tr._on_backbone_deep_done)()
        tr._early = False
        tr._finish_step()
        cov = covered(tr)
        assert [(lo, hi) for lo, hi, _ in cov[0]] in ([(0, 100)], [(0, 60), (60, 100)]), (world, hooks, cov)
        assert [(lo, hi) for lo, hi, _ in cov[1]] == [(0, 40)] and [(lo, hi) for lo, hi, _ in cov[2]] == [(0, 7)]
        assert all(abs(lr - 1e-4) < 1e-12 for _, _, lr in cov[0]) and cov[1][0][2] == 1e-3          # the backbone's own rate
        assert int(tr._step_dev) == 1 and tr._updated == [] and tr._counted is False
        if "deep" in hooks:                        # Mixed_4b..5c: updated early, after the wait for its bucket
            i_wait = events.index(("wait", [2]))
            i_adam = next(i for i, e in enumerate(events) if e[0] == "adam" and e[2] == 40 and e[1] == tr.groups[0][0].data_ptr() + 240)
            assert i_wait < i_adam
        if world > 1:                              # every bucket launched exactly once
            assert sorted(b for e in events if e[0] == "launch" for b in e[1]) == [0, 1, 2, 3]
    # outside a step (a bare forward_backward) the hooks leave the weights alone
    events.clear()
    tr = make(1); tr._early = False
    tr._on_backbone_backward()
    assert not any(e[0] == "adam" for e in events)
