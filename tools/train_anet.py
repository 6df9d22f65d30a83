"""ActivityNet-1.3 training with the reference's command line  (AFSD/anet/train.py:290-350), on one or more B200s:

    python tools/train_anet.py configs/anet_opental.yaml --open_set --split=0 --lw=1 --cw=1 --piou=0.6 [--batch_size 8]
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_anet.py <same arguments>

Same yaml and flags (opental_b200/config.py), the reference's json / npy files (opental_b200/anet_dataset.py), the backbone at
0.1 x the head's learning rate (train.py:303-310 -> Trainer(backbone_lr_scale=0.1)), score maps (action, start, end) holding
class ids (App. D7), the self-supervised second pass through the cut-paste frame map when `--ssl` > 0 and the first sample of
the batch could be augmented (anet/train.py:222-226; pinned by tests/golden/model_anet_ssl.*).  Checkpoints: model `state_dict` interoperates; the optimizer entry is written in
the single-group layout of the THUMOS14 script (the reference's ActivityNet script uses two parameter groups).
Not exercised by the GPU test-suite (needs the dataset); parts: tests/test_anet_dataset_cpu.py, tests/test_model_anet_gpu.py."""
import itertools
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from opental_b200 import anet_dataset as AD, config as C, dataset as D, engine, train_loop  # noqa: E402
from opental_b200.bdnet import BDNet  # noqa: E402
from opental_b200.loader import Prefetcher  # noqa: E402
from opental_b200.multisegment_loss import MultiSegmentLossANet  # noqa: E402


def main(argv=None) -> int:
    parser = C.build_parser()
    parser.add_argument("--no_graph", action="store_true")
    parser.add_argument("--log_json", type=str, default=None)
    parser.add_argument("--loader_threads", type=int, default=4)
    parser.add_argument("--steps_per_epoch", type=int, default=0, help="stop every epoch after this many steps (smoke runs)")
    parser.add_argument("--device", type=str, default="cuda", help="'cuda' (the product has no CPU path; other values are for the test harness)")
    args = parser.parse_args(argv)
    cfg = C.get_config(argv, parser)
    tr_cfg, ds_cfg, model = cfg["training"], cfg["dataset"]["training"], cfg["model"]
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device(args.device, local) if args.device == "cuda" else torch.device(args.device)
    if dev.type == "cuda":
        torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    seed = tr_cfg["random_seed"]
    torch.manual_seed(seed), np.random.seed(seed), random.seed(seed)

    clip_length = ds_cfg["clip_length"]
    net = BDNet.from_config(cfg, use_edl=model.get("use_edl", False), variant="anet", frame_num=clip_length).to(dev)
    net.train()
    kw = C.loss_arguments(cfg)
    crit = MultiSegmentLossANet(kw["num_classes"], kw["overlap_thresh"], 1.0, cls_loss_type=kw["cls_loss_type"],
                                edl_config=kw["edl_config"], os_head=kw["os_head"], clip_length=clip_length).to(dev)
    trainer = engine.Trainer(net, crit, lr=tr_cfg["learning_rate"], weight_decay=tr_cfg["weight_decay"], lw=tr_cfg["lw"],
                             cw=tr_cfg["cw"], ctw=tr_cfg["ctw"], actw=tr_cfg["actw"], ssl_weight=tr_cfg["ssl"], backbone_lr_scale=0.1)
    trainer.broadcast_parameters(0)
    ds = AD.AnetWindows(ds_cfg["video_info_path"], ds_cfg["video_mp4_path"], clip_length, ds_cfg["crop_size"], training=True,
                        binary_class=cfg["dataset"]["num_classes"] == 2)
    batch = tr_cfg["batch_size"]
    if rank == 0:
        print(f"{len(ds)} videos; {len(ds) // (batch * world)} steps per epoch at batch {batch} x {world} GPUs")
    net.backbone.crop_size = ds_cfg["crop_size"]
    net.backbone.crop_offsets = torch.zeros(batch, 3, dtype=torch.int32, device=dev)

    def make_batches(epoch):
        # loader threads -> pinned ring -> copy stream (opental_b200/loader.py); the ingest kernel reads the crop / mirror
        # decisions from the static tensor below, so a captured step graph sees every update
        pf = Prefetcher(ds, batch, epoch, rank=rank, world=world, seed=seed, device=dev, workers=args.loader_threads,
                        crop_offsets=net.backbone.crop_offsets, ssl=tr_cfg["ssl"] > 0)
        return itertools.islice(iter(pf), args.steps_per_epoch) if args.steps_per_epoch > 0 else pf

    ck = tr_cfg["checkpoint_path"]
    st = os.path.join(ck, "training")
    if rank == 0:
        os.makedirs(st, exist_ok=True)
    hist = train_loop.fit(trainer, make_batches, max_epoch=tr_cfg["max_epoch"], resume=tr_cfg["resume"], checkpoint_path=ck,
                          train_state_path=st, use_graph=not args.no_graph, log=lambda line: print(line, flush=True))
    if dev.type == "cuda":
        torch.cuda.synchronize()
    if rank == 0 and args.log_json:
        with open(args.log_json, "w") as fh:
            json.dump(hist, fh)
    if world > 1:
        engine.shutdown_distributed([trainer])
    return 0 if all(np.isfinite(h["cost"]) for h in hist if h.get("steps")) else 1


if __name__ == "__main__":
    sys.exit(main())
