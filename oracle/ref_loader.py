"""Import the *reference's own* Python code (read-only, from /root/reference) on CPU  —  TEST INFRASTRUCTURE.

Only usable in the build container (the GPU box has no /root/reference); used by oracle/make_golden.py to pin the
oracle restatement and to generate tests/golden/*.  Recipe: SURVEY.md §8(c).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("OPENTAL_REFERENCE", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "AFSD")):
    # the GPU box has no /root/reference: the hot-path modules placed by oracle/build_ref.py:build_py()
    REF_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_src")


def kernel_emulation_forward(inp: torch.Tensor, seg: torch.Tensor) -> torch.Tensor:
    """Scalar emulation of BoundaryPoolingForward's index math (boundary_max_pooling_kernel.cu:18-46)."""
    B, C, T = inp.shape
    K = seg.shape[1]
    x = inp.detach().contiguous().numpy().reshape(-1)
    s = seg.detach().contiguous().numpy().reshape(-1)
    out = np.zeros(B * C * K, dtype=x.dtype)
    for index in range(B * C * K):
        k = index % K
        c = (index // K) % C
        n = index // K // C
        st = c // (C // 2)
        si = n * K * 4 + k * 4 + st * 2
        l = min(max(0, int(s[si])), T - 1)
        r = min(max(0, int(s[si + 1])), T - 1)
        base = n * C * T + c * T
        m = x[base + l]
        for i in range(l + 1, r + 1):
            if x[base + i] > m:
                m = x[base + i]
        out[index] = m
    return torch.from_numpy(out.reshape(B, C, K))


def kernel_emulation_backward(gout: torch.Tensor, inp: torch.Tensor, seg: torch.Tensor, compat: bool = True) -> torch.Tensor:
    """Scalar emulation of BoundaryPoolingBackward incl. the launcher's tscale = grad_output.size(2)
    (boundary_max_pooling_kernel.cu:49-82, :114-145)."""
    B, C, T = inp.shape
    K = seg.shape[1]
    ts = K if compat else T
    x = inp.detach().contiguous().numpy().reshape(-1)
    s = seg.detach().contiguous().numpy().reshape(-1)
    g = gout.detach().contiguous().numpy().reshape(-1)
    gin = np.zeros(B * C * T, dtype=x.dtype)
    for index in range(B * C * K):
        k = index % K
        c = (index // K) % C
        n = index // K // C
        st = c // (C // 2)
        si = n * K * 4 + k * 4 + st * 2
        l = min(max(0, int(s[si])), ts - 1)
        r = min(max(0, int(s[si + 1])), ts - 1)
        base = n * C * ts + c * ts
        m, am = x[base + l], l
        for i in range(l + 1, r + 1):
            if x[base + i] > m:
                m, am = x[base + i], i
        gin[base + am] += g[index]
    return torch.from_numpy(gin.reshape(B, C, T))


def _fast_stub(oracle_mod, compat: bool):
    """A stand-in `boundary_max_pooling_cuda` for running the reference model on CPU at full size: forward =
    vectorised oracle max (bit-exact: pure max), backward = oracle restatement.  The scalar emulation above pins
    both on small cases in make_golden.py."""
    mod = types.ModuleType("boundary_max_pooling_cuda")

    def forward(inp, seg):
        out, _ = oracle_mod._bmp_max_argmax(inp, oracle_mod._bmp_windows(seg, inp.shape[2]))
        return out

    def backward(gout, inp, seg):
        with torch.enable_grad():
            x = inp.detach().clone().requires_grad_(True)
            y = oracle_mod.boundary_max_pooling(x, seg, compat)
            (gx,) = torch.autograd.grad(y, x, gout)
        return gx

    mod.forward, mod.backward = forward, backward
    return mod


def load_reference(config="configs/thumos14_opental_final.yaml", extra_args=("--open_set", "--split=0", "--lw=1", "--cw=10",
                   "--ctw=1", "--ssl=0.001", "--piou=0.5"), compat=True, flavour="thumos14"):
    """Returns a namespace with the reference's BDNet module, MultiSegmentLoss class and config dict.
    flavour: 'thumos14' or 'anet' (AFSD/<flavour>/BDNet.py, multisegment_loss.py).  The reference evaluates its config
    once per process at import (AFSD/common/config.py:101): load ONE flavour per process."""
    import importlib

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import opental_oracle as oracle_mod

    if not os.path.isdir(os.path.join(REF_ROOT, "AFSD")):
        raise RuntimeError(f"{REF_ROOT} not present: run oracle/build_ref.py in the build container first")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    argv = sys.argv
    sys.argv = ["ref", os.path.join(REF_ROOT, config), *extra_args]
    sys.modules["boundary_max_pooling_cuda"] = _fast_stub(oracle_mod, compat)
    cuda_orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self        # cls_loss.py:114 calls .cuda() unconditionally
    try:
        cwd = os.getcwd()
        bdnet = importlib.import_module(f"AFSD.{flavour}.BDNet")
        msl = importlib.import_module(f"AFSD.{flavour}.multisegment_loss")
        cfg = importlib.import_module("AFSD.common.config").config
        os.chdir(cwd)
    finally:
        sys.argv = argv
    ns = types.SimpleNamespace(BDNet=bdnet.BDNet, bdnet_module=bdnet, MultiSegmentLoss=msl.MultiSegmentLoss,
                               config=cfg, restore_cuda=lambda: setattr(torch.Tensor, "cuda", cuda_orig))
    return ns


def reference_training_cost(ns, net, crit, x, targets, scores, lw=1.0, cw=10.0, ctw=1.0, actw=1.0):
    """Cost of one non-SSL training step of the reference's own modules: `forward_one_epoch` + the weighting of
    `run_one_epoch` (thumos14/train.py:164-235), restated because importing train.py needs tensorboardX and creates
    directories at import.  Returns (cost, 7 losses, loss_start, loss_end)."""
    import torch.nn.functional as F

    def bce(start, end, sc):
        s = torch.tanh(start).mean(-1)
        e = torch.tanh(end).mean(-1)
        return (F.binary_cross_entropy(s.view(-1), sc[:, 0].contiguous().view(-1)),
                F.binary_cross_entropy(e.view(-1), sc[:, 1].contiguous().view(-1)))

    out = net(x)
    l, c, pl, pc, ct, act, pact = crit(out, [t.clone() for t in targets])
    ls, le = bce(out["start"], out["end"], scores)
    sc4 = F.interpolate(scores, scale_factor=1.0 / 4)
    a, b = bce(out["start_loc_prop"], out["end_loc_prop"], sc4)
    c2, d = bce(out["start_conf_prop"], out["end_conf_prop"], sc4)
    ls = ls + 0.1 * (a + c2)
    le = le + 0.1 * (b + d)
    cost = lw * l + cw * c + lw * pl + cw * pc + ctw * ct + ls + le
    if act is not None:
        cost = cost + actw * act + actw * pact
    return cost, (l, c, pl, pc, ct, act, pact), ls, le
