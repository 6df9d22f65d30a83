"""The training driver around `Trainer.step`  (the semantics of AFSD/thumos14/train.py:203-300, 360-380).

What the reference's script does per epoch and per iteration, minus tensorboard / tqdm:
  * epoch i: `CPD_Loss.cls_loss.epoch = i`, `.total_epoch = max_epoch` (train.py:376-379) — the IBM re-weighting switches
    on at `ibm_start`, so a captured step graph is re-captured when that flag flips;
  * iteration: main pass on (clips, targets, scores); the self-supervised second pass only when `flags[0]` — the FIRST
    sample of the batch could be cut-paste augmented (train.py:237-242) — weighted by `config['training']['ssl']`;
  * running means of the loss terms over the epoch (train.py:270-289), one summary line per epoch (:297-304);
  * `save_model` after every epoch > 10 (train.py:291-293) in the reference's two-file layout (opental_b200/checkpoint.py).
Loss values stay on the device during the epoch (one stack + one host copy at the end): nothing here synchronises the step.

A batch is a dict: clips, targets, scores and optionally flags (list of bool), ssl_targets and ONE of ssl_clips (the loader's
augmented fp32 clips) or ssl_frame_map (int32 [B,T], opental_b200.augment.cut_paste — with uint8 clips)."""
from __future__ import annotations

from typing import Callable, Iterable

import torch

TERMS = ("cost", "loc", "conf", "prop_loc", "prop_conf", "center", "start", "end", "act", "prop_act")


def _ssl_active(batch: dict) -> bool:
    flags = batch.get("flags")
    has = batch.get("ssl_targets") is not None and (batch.get("ssl_clips") is not None or batch.get("ssl_frame_map") is not None)
    return bool(has and (flags is None or bool(flags[0])))


def run_one_epoch(trainer, batches: Iterable[dict], epoch: int, *, use_graph: bool = True,
                  on_step: Callable[[int, dict], None] | None = None) -> dict:
    """One pass over `batches`; returns the epoch means of TERMS (+ 'grad_norm', 'steps', 'ssl_steps')."""
    rows, norms, n_ssl = [], [], 0
    for it, b in enumerate(batches):
        ssl = _ssl_active(b)
        kw = {}
        if ssl:
            kw = dict(ssl_targets=b["ssl_targets"], ssl_clips=b.get("ssl_clips"), ssl_frame_map=b.get("ssl_frame_map"))
            kw = {k: v for k, v in kw.items() if v is not None}
            n_ssl += 1
        if use_graph and not trainer.select_graph(ssl, b["targets"]):
            # capture: no graph of this flavour yet (one is kept per flavour), the batch flavour changed (with / without the SSL pass), a clip has more
            # ground-truth segments than the graph has slots, or the IBM switch flipped with the epoch.  Flavour changes are rare when most windows can be augmented.
            trainer.capture(b["clips"], b["targets"], b["scores"], **kw)
        cost, losses, ls, le = trainer.step(b["clips"], b["targets"], b["scores"], **kw)
        zero = torch.zeros_like(cost)
        l = [x if x is not None else zero for x in losses]
        rows.append(torch.stack([cost, l[0], l[1], l[2], l[3], l[4], ls, le, l[5], l[6]]))
        norms.append(trainer.grad_norm())
        if on_step is not None:
            on_step(it, dict(cost=cost, losses=losses, loss_start=ls, loss_end=le, ssl=ssl))
    if not rows:
        return dict(steps=0, ssl_steps=0)
    mean = torch.stack(rows).mean(0).tolist()            # the epoch's only device -> host copy
    out = dict(zip(TERMS, mean))
    out.update(grad_norm=float(torch.stack(norms).mean()), steps=len(rows), ssl_steps=n_ssl, epoch=epoch)
    return out


def summary_line(epoch: int, m: dict, prefix: str = "Train") -> str:
    """The reference's per-epoch log line (train.py:297-304)."""
    return ("Epoch-{} {} Loss: Total - {:.5f}, loc - {:.5f}, conf - {:.5f}, prop_loc - {:.5f}, prop_conf - {:.5f}, "
            "IoU - {:.5f}, start - {:.5f}, end - {:.5f}").format(epoch, prefix, m["cost"], m["loc"], m["conf"], m["prop_loc"],
                                                                   m["prop_conf"], m["center"], m["start"], m["end"])


def fit(trainer, make_batches: Callable[[int], Iterable[dict]], *, max_epoch: int, resume: int = 0,
        checkpoint_path: str | None = None, train_state_path: str | None = None, save_after_epoch: int = 10,
        use_graph: bool = True, log: Callable[[str], None] = print) -> list[dict]:
    """`__main__` of train.py:370-380: resume, then epochs start..max_epoch.  make_batches(epoch) yields the epoch's batches
    (rank-sharded by the caller: engine.shard_indices).  Checkpoints are written by rank 0 only."""
    import torch.distributed as dist
    start = 1
    if resume > 0:
        start = trainer.resume(resume, checkpoint_path, train_state_path)
    history = []
    crit = trainer.criterion.cls_loss
    for epoch in range(start, max_epoch + 1):
        if hasattr(crit, "epoch"):
            crit.epoch, crit.total_epoch = epoch, max_epoch
        m = run_one_epoch(trainer, make_batches(epoch), epoch, use_graph=use_graph)
        history.append(m)
        rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
        if m.get("steps") and rank0:
            log(summary_line(epoch, m))
        if epoch > save_after_epoch and checkpoint_path and rank0:
            trainer.save_checkpoint(epoch, checkpoint_path, train_state_path or checkpoint_path)
    return history
