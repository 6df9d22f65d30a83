"""Checkpoint / resume in the reference's on-disk format (AFSD/thumos14/train.py:106-129).

The reference writes two files per epoch: `checkpoint-<epoch>.ckpt` = `model.module.state_dict()` and a training-state
file `checkpoint_<epoch>.ckpt` = `{'optimizer': torch.optim.Adam(net.parameters()).state_dict(), 'state': rng states}`.
Our Trainer keeps its Adam moments in three flat fp32 buffers that mirror the flat parameter buffers, and every
`nn.Parameter` of the model is a (possibly strided) view into one of those parameter buffers — so the per-parameter
moments are the SAME views taken on the moment buffers.  `adam_state_dict` / `load_adam_state_dict` translate between
the two forms; parameter order and `requires_grad` flags are the reference's (tests/test_api_cpu.py), so the index-keyed
`state` / `param_groups` of `torch.optim.Adam` line up and the files interoperate in both directions — in the layout the
Trainer is configured for: one group (thumos14/train.py:321) or, with `backbone_lr_scale != 1`, the ActivityNet script's two groups
numbered backbone-first (anet/train.py:304-311).

Extras the reference loses on resume ride along under their own keys and are ignored by the reference's
`resume_training`: the IBM `weight_accum` buffer, the loss epoch counters and the Trainer's step counter."""
from __future__ import annotations

import os
import random

import numpy as np
import torch


def _locate(p: torch.Tensor, groups) -> tuple[int, int]:
    """(group index, element offset) of parameter `p` inside the flat weight buffers."""
    a = p.data_ptr()
    for gi, (w, _) in enumerate(groups):
        lo = w.data_ptr()
        if lo <= a < lo + 4 * w.numel():
            return gi, (a - lo) // 4
    raise ValueError("parameter does not live in any flat buffer")


def moment_view(p: torch.Tensor, groups, flat_moments) -> torch.Tensor:
    """The view of the flat moment buffer that corresponds to parameter `p` (same shape / strides as p.data)."""
    gi, off = _locate(p, groups)
    return torch.as_strided(flat_moments[gi], tuple(p.shape), tuple(p.stride()), off)


def _ordered(params, param_groups):
    """The optimizer's parameter order: torch.optim.Adam numbers parameters group by group.  `param_groups` = None (one group,
    `params` as given: thumos14/train.py:321) or [(parameters, lr), ...] (the ActivityNet script's two groups, backbone at
    0.1 x the rate first, then coarse_pyramid_detection: anet/train.py:304-311)."""
    if param_groups is None:
        return list(params), None
    ordered = [p for ps, _ in param_groups for p in ps]
    if {id(p) for p in ordered} != {id(p) for p in params} or len(ordered) != len(params):
        raise ValueError("param_groups must partition the model's parameters")
    return ordered, [len(ps) for ps, _ in param_groups]


def adam_state_dict(params, groups, state, *, step: int, lr, betas, eps, weight_decay, param_groups=None) -> dict:
    """`torch.optim.Adam(...).state_dict()`-compatible dict (train.py:115).  params: list(net.parameters()); `param_groups`:
    see _ordered (each group's own learning rate is written, `lr` is the single group's)."""
    ordered, sizes = _ordered(params, param_groups)
    st = {}
    if step > 0:
        for i, p in enumerate(ordered):
            if not p.requires_grad:
                continue                                     # torch keeps no state for parameters that never had a grad
            st[i] = dict(step=torch.tensor(float(step)),
                         exp_avg=moment_view(p, groups, [s["m"] for s in state]).detach().clone().contiguous(),
                         exp_avg_sq=moment_view(p, groups, [s["v"] for s in state]).detach().clone().contiguous())

    def group(rate, idx):
        return dict(lr=rate, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                    capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False, params=idx)
    if sizes is None:
        return dict(state=st, param_groups=[group(lr, list(range(len(ordered))))])
    out, lo = [], 0
    for n, (_, rate) in zip(sizes, param_groups):
        out.append(group(rate, list(range(lo, lo + n))))
        lo += n
    return dict(state=st, param_groups=out)


def load_adam_state_dict(sd: dict, params, groups, state, param_groups=None) -> int:
    """Copy a torch.optim.Adam state_dict (ours or the reference's; `step` an int as in torch 1.9 or a tensor) into the
    flat moment buffers.  Returns the step count.  The file's group layout must be the Trainer's (one group, or the
    ActivityNet script's two): the index -> parameter mapping depends on it."""
    ordered, sizes = _ordered(params, param_groups)
    have = [len(g["params"]) for g in sd["param_groups"]]
    if have != (sizes if sizes is not None else [len(ordered)]):
        raise ValueError(f"optimizer state has parameter groups of {have} parameters, this Trainer has "
                         f"{sizes if sizes is not None else [len(ordered)]} (a single-rate THUMOS14 state and a two-rate "
                         "ActivityNet state are not interchangeable)")
    idx = [i for g in sd["param_groups"] for i in g["params"]]
    step = 0
    for s in state:
        s["m"].zero_()
        s["v"].zero_()
    for k, ent in sd["state"].items():
        p = ordered[idx.index(k)]
        if tuple(ent["exp_avg"].shape) != tuple(p.shape):
            raise ValueError(f"optimizer state {k}: shape {tuple(ent['exp_avg'].shape)} vs parameter {tuple(p.shape)}")
        moment_view(p, groups, [s["m"] for s in state]).copy_(ent["exp_avg"])
        moment_view(p, groups, [s["v"] for s in state]).copy_(ent["exp_avg_sq"])
        step = max(step, int(ent["step"]))
    return step


def get_rng_states() -> list:
    """train.py:79-86."""
    states = [random.getstate(), np.random.get_state(), torch.get_rng_state()]
    if torch.cuda.is_available():
        states.append(torch.cuda.get_rng_state())
    return states


def set_rng_states(states) -> None:
    """train.py:89-94."""
    random.setstate(states[0])
    np.random.set_state(states[1])
    torch.set_rng_state(states[2])
    if torch.cuda.is_available() and len(states) > 3:
        torch.cuda.set_rng_state(states[3])


def _update_latest(src: str, dst: str) -> None:
    """train.py:97-103: `dst` becomes a symlink to `src`."""
    if os.path.lexists(dst):
        os.remove(dst)
    os.symlink(os.path.abspath(src), dst)


def save(trainer, epoch: int, checkpoint_path: str, train_state_path: str) -> tuple[str, str]:
    """train.py:106-118 (`save_model`).  Call on rank 0."""
    os.makedirs(checkpoint_path, exist_ok=True)
    os.makedirs(train_state_path, exist_ok=True)
    model_file = os.path.join(checkpoint_path, f"checkpoint-{epoch}.ckpt")
    torch.save({k: v.detach().cpu().contiguous() for k, v in trainer.net.state_dict().items()}, model_file)
    _update_latest(model_file, os.path.join(checkpoint_path, "checkpoint-latest.ckpt"))
    state_file = os.path.join(train_state_path, f"checkpoint_{epoch}.ckpt")
    torch.save({"optimizer": trainer.optimizer_state_dict(), "state": get_rng_states(),
                "criterion": {k: v.detach().cpu() for k, v in trainer.criterion.state_dict().items()},
                "loss_epoch": (trainer.criterion.cls_loss.epoch, trainer.criterion.cls_loss.total_epoch)
                if hasattr(trainer.criterion.cls_loss, "epoch") else None,
                "step_count": trainer.step_count}, state_file)
    _update_latest(state_file, os.path.join(train_state_path, "checkpoint_latest.ckpt"))
    return model_file, state_file


def resume(trainer, resume_epoch: int, checkpoint_path: str, train_state_path: str) -> int:
    """train.py:121-131 (`resume_training`): returns the epoch to start from."""
    start_epoch = 1
    if resume_epoch > 0:
        start_epoch += resume_epoch
        sd = torch.load(os.path.join(checkpoint_path, f"checkpoint-{resume_epoch}.ckpt"), map_location="cpu")
        trainer.net.load_state_dict(sd)
        st = torch.load(os.path.join(train_state_path, f"checkpoint_{resume_epoch}.ckpt"), map_location="cpu", weights_only=False)
        trainer.load_optimizer_state_dict(st["optimizer"])
        if "criterion" in st:
            trainer.criterion.load_state_dict(st["criterion"], strict=False)
        if st.get("loss_epoch") is not None:
            trainer.criterion.cls_loss.epoch, trainer.criterion.cls_loss.total_epoch = st["loss_epoch"]
        set_rng_states(st["state"])
    return start_epoch
