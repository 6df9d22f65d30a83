"""THUMOS14 training with the reference's command line  (AFSD/thumos14/train.py:307-380), on one or more B200s:

    python tools/train_thumos.py configs/thumos14_opental_final.yaml --open_set --split=0 --lw=1 --cw=10 --ctw=1 --ssl=0.001 --piou=0.5
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_thumos.py <same arguments> --batch_size 8

Same yaml, same flags (opental_b200/config.py), same data files (opental_b200/dataset.py), same checkpoint layout
(opental_b200/checkpoint.py); `--batch_size` is per GPU.  Differences from the reference script, all deliberate:
  * the classification loss follows `edl_loss: true` (the script's own line 31 clobbers it to focal, SURVEY App. D2;
    `--script_compat` reproduces the script as written);
  * `nn.DataParallel` on one GPU becomes one process per GPU with an NCCL gradient all-reduce (`engine.Trainer`);
  * the loader ships uint8 windows; crop / mirror / normalise / cut-paste run in the ingest kernel;
  * no tensorboard (the per-epoch summary line is printed, `--log_json` appends the epoch means to a file).
Not exercised by the GPU test-suite (it needs the dataset); its parts are: config / dataset / train_loop CPU tests and
tools/train_synthetic.py."""
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

from opental_b200 import config as C, dataset as D, engine, train_loop  # noqa: E402
from opental_b200.bdnet import BDNet  # noqa: E402
from opental_b200.loader import Prefetcher  # noqa: E402
from opental_b200.multisegment_loss import MultiSegmentLoss  # noqa: E402


def main(argv=None) -> int:
    parser = C.build_parser()
    parser.add_argument("--script_compat", action="store_true", help="cls_loss_type exactly as thumos14/train.py:27-31 computes it")
    parser.add_argument("--no_graph", action="store_true", help="eager steps instead of CUDA-graph replay")
    parser.add_argument("--log_json", type=str, default=None)
    parser.add_argument("--loader_threads", type=int, default=4)
    parser.add_argument("--steps_per_epoch", type=int, default=0, help="stop every epoch after this many steps (smoke runs)")
    parser.add_argument("--device", type=str, default="cuda", help="'cuda' (the product has no CPU path; other values are for the test harness)")
    args = parser.parse_args(argv)
    cfg = C.get_config(argv, parser)
    tr_cfg, ds_cfg = cfg["training"], cfg["dataset"]["training"]
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device(args.device, local) if args.device == "cuda" else torch.device(args.device)
    if dev.type == "cuda":
        torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    seed = tr_cfg["random_seed"]
    torch.manual_seed(seed), np.random.seed(seed), random.seed(seed)                     # set_seed (train.py:62-70)

    model = cfg["model"]
    net = BDNet.from_config(cfg, use_edl=model.get("use_edl", False), use_rpl=model.get("use_rpl", False)).to(dev)
    net.train()
    kw = C.loss_arguments(cfg, script_compat=args.script_compat)
    crit = MultiSegmentLoss(kw.pop("num_classes"), kw.pop("overlap_thresh"), kw.pop("negpos_ratio"), **kw,
                            clip_length=ds_cfg["clip_length"]).to(dev)
    trainer = engine.Trainer(net, crit, lr=tr_cfg["learning_rate"], weight_decay=tr_cfg["weight_decay"], lw=tr_cfg["lw"],
                             cw=tr_cfg["cw"], ctw=tr_cfg["ctw"], actw=tr_cfg["actw"], ssl_weight=tr_cfg["ssl"])
    trainer.broadcast_parameters(0)

    infos = D.get_video_info(ds_cfg["video_info_path"])
    annos = D.get_video_anno(infos, ds_cfg["video_anno_path"], cfg["dataset"]["class_info_path"])
    data = D.load_video_data(infos, ds_cfg["video_data_path"])
    ds = D.ThumosWindows(data, infos, annos, clip_length=ds_cfg["clip_length"], crop_size=ds_cfg["crop_size"],
                         stride=ds_cfg["clip_stride"], training=True)
    batch = tr_cfg["batch_size"]
    if rank == 0:
        print(f"{len(ds)} windows of {len(annos)} videos; {len(ds) // (batch * world)} steps per epoch at batch {batch} x {world} GPUs; "
              f"loss {crit.cls_loss_type}; lr {tr_cfg['learning_rate']}, wd {tr_cfg['weight_decay']}, max_epoch {tr_cfg['max_epoch']}")
    # the ingest kernel reads the crop / mirror decisions from a static device tensor, so a captured graph sees every update
    net.backbone.crop_offsets = torch.zeros(batch, 3, dtype=torch.int32, device=dev)
    ssl_on = tr_cfg["ssl"] > 0

    def make_batches(epoch):
        # loader threads -> pinned ring -> copy stream (opental_b200/loader.py); the ingest kernel reads the crop / mirror
        # decisions from the static tensor below, so a captured step graph sees every update
        pf = Prefetcher(ds, batch, epoch, rank=rank, world=world, seed=seed, device=dev, workers=args.loader_threads,
                        crop_offsets=net.backbone.crop_offsets, ssl=ssl_on)
        return itertools.islice(iter(pf), args.steps_per_epoch) if args.steps_per_epoch > 0 else pf

    ck = tr_cfg["checkpoint_path"]
    st = os.path.join(ck, "training")                                                    # train.py:37
    if rank == 0:
        os.makedirs(st, exist_ok=True)

    def log(line):
        print(line, flush=True)

    hist = train_loop.fit(trainer, make_batches, max_epoch=tr_cfg["max_epoch"], resume=tr_cfg["resume"], checkpoint_path=ck,
                          train_state_path=st, use_graph=not args.no_graph, log=log)
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
