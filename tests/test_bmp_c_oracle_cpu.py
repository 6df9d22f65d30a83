"""CPU: the plain-C restatement of BoundaryMaxPooling (oracle/bmp_oracle.c, gcc) against the golden cases generated from
the reference kernel's semantics (tests/golden/bmp_cases.npz) and against the torch restatement on random inputs with
degenerate windows (r < l, negative, beyond T, fractional boundaries)."""
import ctypes
import os
import shutil
import sys

import numpy as np
import pytest
import torch

import opental_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")


@pytest.fixture(scope="module")
def clib():
    import build_c
    lib = ctypes.CDLL(build_c.build())
    P, I = ctypes.c_void_p, ctypes.c_int
    lib.bmp_oracle_forward.argtypes = [P, P, P, I, I, I, I]
    lib.bmp_oracle_backward.argtypes = [P, P, P, P, I, I, I, I, I]
    lib.bmp_oracle_forward.restype = lib.bmp_oracle_backward.restype = None
    return lib


def c_forward(lib, inp, seg):
    B, C, T = inp.shape
    K = seg.shape[1]
    inp, seg = np.ascontiguousarray(inp, np.float32), np.ascontiguousarray(seg, np.float32)
    out = np.empty((B, C, K), np.float32)
    lib.bmp_oracle_forward(inp.ctypes.data, seg.ctypes.data, out.ctypes.data, B, C, T, K)
    return out


def c_backward(lib, gout, inp, seg, compat):
    B, C, T = inp.shape
    K = seg.shape[1]
    gout, inp, seg = (np.ascontiguousarray(a, np.float32) for a in (gout, inp, seg))
    gin = np.empty((B, C, T), np.float32)
    lib.bmp_oracle_backward(gout.ctypes.data, inp.ctypes.data, seg.ctypes.data, gin.ctypes.data, B, C, T, K, int(compat))
    return gin


@pytest.mark.parametrize("name", ["level", "frame", "ssl", "tiny", "wide"])
def test_c_oracle_matches_golden(clib, golden_dir, name):
    g = np.load(os.path.join(golden_dir, "bmp_cases.npz"))
    inp, seg, gout = g[f"{name}.inp"], g[f"{name}.seg"], g[f"{name}.gout"]
    assert np.array_equal(c_forward(clib, inp, seg), g[f"{name}.fwd"])                      # pure max: bit exact
    assert np.allclose(c_backward(clib, gout, inp, seg, False), g[f"{name}.bwd_fixed"], atol=1e-6)
    if f"{name}.bwd_compat" in g:
        assert np.allclose(c_backward(clib, gout, inp, seg, True), g[f"{name}.bwd_compat"], atol=1e-6)


@pytest.mark.parametrize("seed", range(6))
def test_c_oracle_matches_torch_oracle_on_degenerate_windows(clib, seed):
    gen = torch.Generator().manual_seed(seed)
    B, C, T, K = 2, 6, 17 + seed, 9
    inp = torch.randn(B, C, T, generator=gen)
    inp[:, :, 3] = inp[:, :, 4]                                    # exact ties: the first maximum wins
    seg = (torch.rand(B, K, 4, generator=gen) * (T + 8) - 4)       # negative, beyond T, fractional, r < l
    gout = torch.randn(B, C, K, generator=gen)
    x = inp.clone().requires_grad_(True)
    y = O.boundary_max_pooling(x, seg, False)
    (gx,) = torch.autograd.grad(y, x, gout)
    assert np.array_equal(c_forward(clib, inp.numpy(), seg.numpy()), y.detach().numpy())
    assert np.allclose(c_backward(clib, gout.numpy(), inp.numpy(), seg.numpy(), False), gx.numpy(), atol=1e-6)
    if K <= T:
        (gc,) = torch.autograd.grad(O.boundary_max_pooling(x, seg, True), x, gout)
        assert np.allclose(c_backward(clib, gout.numpy(), inp.numpy(), seg.numpy(), True), gc.numpy(), atol=1e-6)
