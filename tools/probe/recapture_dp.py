"""Probe: capture the data-parallel step graph twice on one Trainer (different batch sizes) — the sweep's pattern."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from opental_b200 import engine
from opental_b200.multisegment_loss import pad_targets
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
net, crit = engine.build_opental(device=dev, frame_num=128, epoch=11)
tr = engine.Trainer(net, crit); tr.broadcast_parameters(0)
keep = []
MODE = os.environ.get("PROBE_MODE", "destroy")      # destroy | keep | nograph_first
if MODE == "nograph_first":
    tr.graph_update = False
for B in (1, 2, 4):
    clips = torch.randint(0, 256, (B, 128, 112, 112, 3), dtype=torch.uint8, device=dev)
    tg = [engine.synthetic_targets(i, rank) for i in range(B)]
    sc = torch.stack([engine.synthetic_scores(t, frames=128) for t in tg]).to(dev)
    tp, tv = (t.to(dev) for t in pad_targets(tg, device="cpu"))
    try:
        tr.capture(clips, (tp, tv), sc)
        for _ in range(3):
            cost, *_ = tr.step(clips, (tp, tv), sc)
        torch.cuda.synchronize()
        sys.stdout.write(f"{rank} B {B} ok {float(cost):.6f} {torch.cuda.memory_allocated() >> 20} MiB graph_updates {tr._graph_updates}\n"); sys.stdout.flush()
    except Exception:
        sys.stdout.write(f"{rank} B {B} FAILED\n"); sys.stdout.flush()
        traceback.print_exc()
        break
    finally:
        if MODE == "keep":
            keep.append((tr._graph, tr._static, tr._graph_out))
        tr._graph = tr._graph_out = tr._static = None
        tr._graph_cache.clear()
if world > 1:
    chk = torch.stack([w.double().sum() for w, _ in tr.groups])
    both = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    sys.stdout.write(f"{rank} params in sync: {all(bool((b == both[0]).all()) for b in both)}\n"); sys.stdout.flush()
engine.shutdown_distributed([tr])
