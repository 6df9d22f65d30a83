"""End-to-end training on a synthetic dataset: the loader-side pieces (window index, score maps, cut-paste frame maps, random
crop / mirror offsets) feeding `train_loop.fit` exactly as a THUMOS14 run would (AFSD/thumos14/train.py), on one or more
GPUs, through the product loader (opental_b200/dataset.py + loader.py).  A development / validation tool (the benchmark is bench.py):

    python tools/train_synthetic.py --videos 6 --epochs 12 --batch 4 --ibm-start 3 --out gpurun_out/train_synth
    torchrun --nproc-per-node 2 tools/train_synthetic.py ...

Checks while it runs: losses finite, the step graph is re-captured only when the batch flavour / IBM switch changes,
checkpoints appear after epoch 10 and `--resume E` continues from them, ranks stay bit-identical."""
import argparse
import itertools
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from opental_b200 import dataset, engine, loader, train_loop  # noqa: E402


def synthetic_dataset(n_videos: int, seed: int = 0):
    """uint8 videos [count,112,112,3] (AFSD/common/video2npy.py:61-74) with annotation tables in sampled frames."""
    r = random.Random(seed)
    infos, annos, data = {}, {}, {}
    for v in range(n_videos):
        name = f"video_{v:04d}"
        count = r.choice([300, 420, 517, 640])
        infos[name] = dict(fps=30.0, sample_fps=10.0, count=count * 3, sample_count=count)
        segs, t = [], r.uniform(5, 40)
        while t + 40 < count:
            length = r.uniform(30, 110)
            segs.append([t, min(t + length, count - 2.0), r.randint(1, 15)])
            t += length + r.uniform(25, 90)
        annos[name] = segs or [[10.0, 80.0, 1]]
        g = torch.Generator().manual_seed(1000 + v)
        data[name] = torch.randint(0, 256, (count, 112, 112, 3), generator=g, dtype=torch.uint8).numpy()
    return infos, annos, data


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=6)
    ap.add_argument("--epochs", type=int, default=12)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--ibm-start", type=int, default=3)
    ap.add_argument("--resume", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--loader-threads", type=int, default=4)
    ap.add_argument("--steps-per-epoch", type=int, default=0, help="stop every epoch after this many steps (smoke runs)")
    ap.add_argument("--device", default="cuda", help="'cuda' (the product has no CPU path; other values are for the test harness)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "train_synth"))
    args = ap.parse_args(argv)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device("cuda", local) if args.device == "cuda" else torch.device(args.device)
    if dev.type == "cuda":
        torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    clip_length, crop = 256, 96
    infos, annos, data = synthetic_dataset(args.videos)
    ds = dataset.ThumosWindows(data, infos, annos, clip_length=clip_length, crop_size=crop, stride=30, training=True)
    if rank == 0:
        print(f"{len(ds)} windows from {args.videos} videos; cut-paste thresholds {sorted(ds.th.values())}")

    torch.manual_seed(0)
    net, crit = engine.build_opental(device=dev, epoch=1)
    crit.cls_loss.ibm_start = args.ibm_start
    tr = engine.Trainer(net, crit, lr=1e-5, weight_decay=1e-3, ssl_weight=0.001)
    tr.broadcast_parameters(0)
    # RandomCrop + RandomHorizontalFlip happen in the ingest kernel; a static buffer, so a captured graph sees every update
    net.backbone.crop_offsets = torch.zeros(args.batch, 3, dtype=torch.int32, device=dev)
    captures = []
    capture = tr.capture
    tr.capture = lambda *a, **k: (captures.append((crit.cls_loss.epoch, sorted(k))), capture(*a, **k))[1]

    def make_batches(epoch):
        # the product's loader: window index -> loader threads -> pinned ring -> copy stream (opental_b200/loader.py)
        pf = loader.Prefetcher(ds, args.batch, epoch, rank=rank, world=world, seed=0, device=dev, workers=args.loader_threads,
                               crop_offsets=net.backbone.crop_offsets)
        return itertools.islice(iter(pf), args.steps_per_epoch) if args.steps_per_epoch > 0 else pf

    ck, st = os.path.join(args.out, "checkpoint"), os.path.join(args.out, "train_state")
    hist = train_loop.fit(tr, make_batches, max_epoch=args.epochs, resume=args.resume, checkpoint_path=ck, train_state_path=st,
                          use_graph=not args.no_graph)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    ok = all(np.isfinite(h["cost"]) for h in hist if h.get("steps"))
    if world > 1:
        chk = torch.stack([w.double().sum() for w, _ in tr.groups])
        hi, lo = chk.clone(), chk.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        ok = ok and bool((hi == lo).all())
    if rank == 0:
        print(f"captures (epoch, kwargs): {captures}")
        print(f"checkpoints: {sorted(os.listdir(ck)) if os.path.isdir(ck) else []}")
        print("TRAIN_SYNTHETIC", "OK" if ok else "FAILED")
    if world > 1:
        engine.shutdown_distributed([tr])
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
