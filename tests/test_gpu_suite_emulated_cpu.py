"""CPU: selected `-m gpu` tests run UNCHANGED against the C-ABI emulation (tests/abi_emu.py).

The GPU test modules move their data with `.cuda()`; here `.cuda()` is the identity, `engine.build_opental*` build on the CPU
with the loss in its torch formulation, and every `_lib.call` lands in the emulator.  So the same assertions — golden vectors
of the reference, same tolerances — check the product's host code on every CPU run: per-endpoint backbone features, the
self-supervised (triplet) pass, the closed-set and the ActivityNet flavours, the head with forced windows, inference post-processing
(the frame-map form of the SSL pass and checkpoint / resume are covered by tests/test_model_emulated_cpu.py).
The kernels are what the real `-m gpu` run checks."""
import functools
import importlib
import math

import pytest
import torch

import abi_emu


@pytest.fixture
def emulated(monkeypatch):
    from opental_b200 import engine
    emu = abi_emu.install(monkeypatch)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)

    def on_cpu(build):
        @functools.wraps(build)
        def wrapper(*args, **kw):
            kw["device"] = "cpu"
            net, crit = build(*args, **kw)
            if hasattr(crit, "fused"):
                crit.fused = False              # the loss in its torch formulation: the single-CTA kernel is GPU-only
            return net, crit
        return wrapper

    monkeypatch.setattr(engine, "build_opental", on_cpu(engine.build_opental))
    monkeypatch.setattr(engine, "build_opental_anet", on_cpu(engine.build_opental_anet))
    return emu


def gpu_test(module: str, name: str):
    fn = getattr(importlib.import_module(module), name)
    return getattr(fn, "__wrapped__", fn)


def golden_pair(golden_dir, stem):
    import json
    import os

    import numpy as np
    arrays = np.load(os.path.join(golden_dir, stem + ".npz"))
    with open(os.path.join(golden_dir, stem + ".json")) as fh:
        return arrays, json.load(fh)


def test_backbone_endpoints(emulated, golden_dir):
    gpu_test("test_model_gpu", "test_backbone_endpoints_match_reference_golden")(golden_pair(golden_dir, "model_thumos_opental"))


def test_ssl_triplet_pass(emulated, golden_dir):
    gpu_test("test_model_gpu", "test_ssl_triplet_pass_matches_reference_golden")(golden_dir)


def test_closed_set_model(emulated, golden_dir):
    gpu_test("test_model_closed_gpu", "test_closed_set_forward_focal_loss_backward_match_reference_golden")(golden_dir, "biased", math.log(32.0))




def test_activitynet_model(emulated, golden_dir):
    gpu_test("test_model_anet_gpu", "test_anet_forward_loss_backward_match_reference_golden")(golden_dir)


def test_head_with_forced_windows(emulated):
    gpu_test("test_head_gpu", "test_head_forward_backward_matches_oracle_with_forced_windows")(2)


def test_inference_post_processing(emulated, golden_dir):
    import numpy as np
    import os
    mod = importlib.import_module("test_infer_gpu")
    golden = np.load(os.path.join(golden_dir, "infer_cases.npz"))
    mod.test_decode_scores_matches_reference_golden.__wrapped__(golden) if hasattr(mod.test_decode_scores_matches_reference_golden, "__wrapped__") \
        else mod.test_decode_scores_matches_reference_golden(golden)
    mod.test_decode_scores_batch_matches_oracle()
    mod.test_softnms_many_classes_matches_oracle()


def test_activitynet_ssl_triplet_pass(emulated, golden_dir):
    gpu_test("test_model_anet_gpu", "test_anet_ssl_triplet_pass_matches_reference_golden")(golden_dir)


def test_staged_uint8_conv1a_gpu_tests_run_against_the_emulation(emulated):
    """tests/test_conv1a_u8_gpu.py is the first thing the next GPU call runs: its own logic (shapes, helper calls, tolerances) is
    checked here against the emulated entry points, so that a GPU minute is never spent on a broken test."""
    mod = importlib.import_module("test_conv1a_u8_gpu")
    mod.test_raw_ingest_is_exact()
    for shape in [(1, 16, 24, 24), (2, 12, 16, 32), (1, 8, 8, 16)]:
        mod.test_conv1a_fwd_u8(shape)
    mod.test_border_class_sums()
    for shape in [(2, 12, 16, 16), (1, 8, 16, 16)]:
        mod.test_conv1a_wgrad_u8(shape)
