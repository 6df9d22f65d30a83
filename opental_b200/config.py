"""The reference's command line and yaml configuration  (AFSD/common/config.py:5-98), as a function of argv.

Same flags, same defaults, same overrides and the same `{id}` substitution of the open-set split into the paths, so the
scripts of `experiments/` can be pointed at `tools/train_thumos.py` unchanged.  The reference evaluates this once per process
at import time (`config = get_config()`, :101) and every module reads the global; here the dict is passed explicitly
(`BDNet.from_config`, `tools/train_thumos.py`).  Pinned to the reference's own function on a synthetic yaml
(oracle/make_golden.py --config, tests/golden/config_cases.json)."""
from __future__ import annotations

import argparse

import yaml


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    p.add_argument("config_file", type=str, default="configs/default.yaml", nargs="?")
    for name, typ in (("batch_size", int), ("learning_rate", float), ("weight_decay", float), ("max_epoch", int),
                      ("checkpoint_path", str), ("seed", int), ("focal_loss", bool), ("nms_thresh", float), ("nms_sigma", float),
                      ("top_k", int), ("output_json", str)):
        p.add_argument("--" + name, type=typ)
    for name, typ, default in (("lw", float, 1.0), ("cw", float, 10.0), ("ctw", float, 1.0), ("actw", float, 1.0), ("ssl", float, 0.1),
                               ("piou", float, 0), ("resume", int, 0), ("ngpu", int, 1)):
        p.add_argument("--" + name, type=typ, default=default)
    p.add_argument("--fusion", action="store_true")
    p.add_argument("--open_set", action="store_true")
    p.add_argument("--split", type=int, choices=[0, 1, 2, 3, 4], default=0)
    p.add_argument("--ood_scoring", type=str, default="confidence",
                   choices=["uncertainty", "confidence", "uncertainty_actionness", "a_by_inv_u", "u_by_inv_a", "half_au"])
    p.add_argument("--exp_tag", type=str, default=None)
    return p


def get_config(argv: list[str] | None = None, parser: argparse.ArgumentParser | None = None) -> dict:
    """argv (without the program name; None = sys.argv[1:]) -> the configuration dict of config.py:40-98.  `parser` lets a
    script add its own flags; flags this function does not know are left in the namespace (`data['args']` is not added)."""
    args = (parser or build_parser()).parse_args(argv)
    with open(args.config_file, "r", encoding="utf-8") as fh:
        data = yaml.load(fh.read(), Loader=yaml.FullLoader)
    tr, te = data["training"], data["testing"]
    tr["learning_rate"], tr["weight_decay"] = float(tr["learning_rate"]), float(tr["weight_decay"])
    if args.batch_size is not None:
        tr["batch_size"] = int(args.batch_size)
    if args.learning_rate is not None:
        tr["learning_rate"] = float(args.learning_rate)
    if args.weight_decay is not None:
        tr["weight_decay"] = float(args.weight_decay)
    if args.max_epoch is not None:
        tr["max_epoch"] = int(args.max_epoch)
    if args.checkpoint_path is not None:
        tr["checkpoint_path"] = te["checkpoint_path"] = args.checkpoint_path
    if args.seed is not None:
        tr["random_seed"] = args.seed
    if args.focal_loss is not None:
        tr["focal_loss"] = args.focal_loss
    for k in ("lw", "cw", "ctw", "actw", "ssl", "piou", "resume"):
        tr[k] = getattr(args, k)
    data["ngpu"] = args.ngpu
    te["fusion"], te["split"], te["ood_scoring"] = args.fusion, args.split, args.ood_scoring
    for k in ("nms_thresh", "nms_sigma", "top_k", "output_json", "exp_tag"):
        if getattr(args, k) is not None:
            te[k] = getattr(args, k)
    data["open_set"] = args.open_set
    if args.open_set:
        ds = data["dataset"]
        ds["class_info_path"] = ds["class_info_path"].format(id=args.split)
        for part in ("training", "testing"):
            ds[part]["video_anno_path"] = ds[part]["video_anno_path"].format(id=args.split)
            vip = ds[part]["video_info_path"]
            ds[part]["video_info_path"] = vip.format(id=args.split) if "split_" in vip else vip
        tr["checkpoint_path"] = tr["checkpoint_path"].format(id=args.split)
        te["checkpoint_path"] = te["checkpoint_path"].format(id=args.split)
        te["output_path"] = te["output_path"].format(id=args.split)
    return data


def loss_arguments(config: dict, script_compat: bool = False) -> dict:
    """Keyword arguments of `MultiSegmentLoss(num_cls, piou, 1.0, ...)` as train.py:22-35, 329-331 derives them.
    Reference quirk: thumos14/train.py assigns `cls_loss_type` twice — 'edl' at :27, then
    `'rpl' if rpl_loss else 'focal'` at :31 — so the shipped THUMOS14 script trains the focal loss even with `edl_loss: true`
    (anet/train.py:24 is correct).  Default here is the intended precedence rpl > edl > focal; `script_compat=True` reproduces
    the script as written."""
    tr, model = config["training"], config.get("model", {})
    os_head = bool(model.get("os_head", False))
    num_classes = config["dataset"]["num_classes"]
    cls_loss_type = "rpl" if tr.get("rpl_loss", False) else ("edl" if tr.get("edl_loss", False) and not script_compat else "focal")
    return dict(num_classes=num_classes - 1 if os_head else num_classes, overlap_thresh=tr["piou"], negpos_ratio=1.0,
                cls_loss_type=cls_loss_type, edl_config=tr.get("edl_config"), rpl_config=tr.get("rpl_config"), os_head=os_head,
                act_config=tr.get("act_config"))
