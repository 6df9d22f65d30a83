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
