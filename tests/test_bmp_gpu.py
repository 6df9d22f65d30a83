"""GPU parity: BoundaryMaxPooling through the C ABI vs (a) the golden vectors from the reference kernel emulation,
(b) the CPU oracle on seeded inputs at the real call shapes, (c) the reference's OWN CUDA kernel compiled
unmodified into oracle/_ref, plus size-independent properties and the reference's error behaviour."""
import os

import numpy as np
import pytest
import torch

import opental_oracle as O

pytestmark = pytest.mark.gpu


def make_case(B, C, T, K, seed, fractional=False):
    g = torch.Generator().manual_seed(seed)
    inp = torch.randn(B, C, T, generator=g)
    c = torch.rand(B, K, 1, generator=g) * T
    seg = torch.cat([c - torch.rand(B, K, 1, generator=g) * 0.2 * T - 1, c + torch.rand(B, K, 1, generator=g) * 0.1 * T,
                     c - torch.rand(B, K, 1, generator=g) * 0.1 * T, c + torch.rand(B, K, 1, generator=g) * 0.2 * T + 1], -1)
    seg = seg / 1.3 if fractional else seg.round()
    gout = torch.randn(B, C, K, generator=g)
    return inp, seg, gout


@pytest.fixture(scope="module")
def ops():
    from opental_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("name", ["level", "frame", "ssl", "tiny", "wide"])
def test_golden_cases(ops, golden_dir, name):
    g = np.load(os.path.join(golden_dir, "bmp_cases.npz"))
    inp, seg, gout = (torch.from_numpy(g[f"{name}.{k}"]).cuda() for k in ("inp", "seg", "gout"))
    assert torch.equal(ops.bmp_forward(inp, seg).cpu(), torch.from_numpy(g[f"{name}.fwd"]))       # bit exact
    fixed = ops.bmp_backward(gout, inp, seg, False).cpu()
    assert torch.allclose(fixed, torch.from_numpy(g[f"{name}.bwd_fixed"]), atol=1e-6)
    if f"{name}.bwd_compat" in g:
        compat = ops.bmp_backward(gout, inp, seg, True).cpu()
        assert torch.allclose(compat, torch.from_numpy(g[f"{name}.bwd_compat"]), atol=1e-6)


# the real call shapes of one THUMOS14 forward: level calls T == K, frame calls T=256, K=t
SHAPES = [(2, 1024, 64, 64), (2, 1024, 2, 2), (2, 512, 256, 64), (2, 512, 256, 2), (1, 1024, 96, 96), (1, 512, 768, 48),
          (3, 10, 37, 5)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("fractional", [False, True])
def test_vs_oracle(ops, shape, fractional):
    B, C, T, K = shape
    inp, seg, gout = make_case(B, C, T, K, seed=sum(shape), fractional=fractional)
    x = inp.clone().requires_grad_(True)
    y = O.boundary_max_pooling(x, seg, False)
    (gx,) = torch.autograd.grad(y, x, gout)
    out = ops.bmp_forward(inp.cuda(), seg.cuda())
    assert torch.equal(out.cpu(), y.detach())
    gi = ops.bmp_backward(gout.cuda(), inp.cuda(), seg.cuda(), False)
    assert torch.allclose(gi.cpu(), gx, atol=1e-5, rtol=1e-5)
    if K <= T:
        (gxc,) = torch.autograd.grad(O.boundary_max_pooling(x, seg, True), x, gout)
        gic = ops.bmp_backward(gout.cuda(), inp.cuda(), seg.cuda(), True)
        assert torch.allclose(gic.cpu(), gxc, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("shape", [(8, 1024, 64, 64), (8, 512, 256, 64), (8, 512, 256, 8)])
def test_vs_compiled_reference_kernel(ops, shape):
    """The reference's own kernel (oracle/_ref, built from /root/reference sources): forward bit-equal, backward
    equal up to fp32 atomic-order noise, including the tscale quirk for T != K."""
    import build_ref
    ref = build_ref.load_module()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    B, C, T, K = shape
    inp, seg, gout = (t.cuda() for t in make_case(B, C, T, K, seed=7))
    assert torch.equal(ops.bmp_forward(inp, seg), ref.forward(inp, seg))
    assert torch.allclose(ops.bmp_backward(gout, inp, seg, True), ref.backward(gout, inp, seg), atol=1e-5, rtol=1e-5)


def test_properties_full_size(ops):
    """Size-independent properties at BASELINE batch size (B=8): idempotence on constant rows, monotonicity,
    window containment (out is an element of the row), gradient mass conservation, determinism."""
    B, C, T, K = 8, 1024, 64, 64
    inp, seg, gout = (t.cuda() for t in make_case(B, C, T, K, seed=11))
    out = ops.bmp_forward(inp, seg)
    assert torch.equal(ops.bmp_forward(torch.full_like(inp, 3.25), seg), torch.full_like(out, 3.25))
    assert bool((ops.bmp_forward(inp + 1.0, seg) >= out).all())
    assert bool((out <= inp.amax(dim=2, keepdim=True)).all()) and bool((out >= inp.amin(dim=2, keepdim=True)).all())
    gi = ops.bmp_backward(gout, inp, seg, False)
    assert torch.allclose(gi.sum(dim=2), gout.sum(dim=2), atol=1e-3)
    assert torch.equal(gi, ops.bmp_backward(gout, inp, seg, False))      # deterministic (no float atomics)
    wide = seg.clone(); wide[..., 0] = -5; wide[..., 1] = T + 5; wide[..., 2] = -5; wide[..., 3] = T + 5
    assert torch.equal(ops.bmp_forward(inp, wide), inp.amax(dim=2, keepdim=True).expand(B, C, K))


def test_float64_and_nan(ops):
    inp, seg, gout = make_case(2, 8, 16, 4, seed=5)
    y = O.boundary_max_pooling(inp.double(), seg.double(), False)
    assert torch.equal(ops.bmp_forward(inp.double().cuda(), seg.double().cuda()).cpu(), y)
    x = inp.clone(); x[0, 0, 3] = float("nan")
    s = seg.clone(); s[0, 0] = torch.tensor([1.0, 6.0, 1.0, 6.0])
    out = ops.bmp_forward(x.cuda(), s.cuda()).cpu()
    assert not torch.isnan(out[0, 0, 0])        # NaN never wins the strict '>' unless it is the first element


@pytest.mark.parametrize("shape", [(2, 1024, 64, 64), (2, 512, 256, 64), (3, 10, 37, 5)])
def test_half_precision_dispatch(ops, shape):
    """The reference dispatches AT_DISPATCH_FLOATING_TYPES_AND_HALF (boundary_max_pooling_kernel.cu:96,128): float16 input with
    float16 segments (integers below 2048 are exact).  Forward: the maximum is an element of the row — bit-exact against the oracle
    run on the same half values.  Backward: ours sums a frame's contributions in fp32 and rounds once (half ulp of the result);
    the reference's own kernel rounds after every atomicAdd in a run-dependent order, so it is compared in norm."""
    B, C, T, K = shape
    inp, seg, gout = (t.half() for t in make_case(B, C, T, K, seed=3 + sum(shape)))
    x = inp.float().requires_grad_(True)
    y = O.boundary_max_pooling(x, seg.float(), False)
    (gx,) = torch.autograd.grad(y, x, gout.float())
    out = ops.bmp_forward(inp.cuda(), seg.cuda())
    assert out.dtype == torch.float16 and torch.equal(out.cpu(), y.detach().half())
    gi = ops.bmp_backward(gout.cuda(), inp.cuda(), seg.cuda(), False)
    assert gi.dtype == torch.float16
    assert torch.allclose(gi.float().cpu(), gx, rtol=1e-3, atol=1e-6)           # one rounding to half: 2^-11 relative
    assert torch.equal(gi, ops.bmp_backward(gout.cuda(), inp.cuda(), seg.cuda(), False))      # deterministic
    import build_ref
    ref = build_ref.load_module()
    if ref is not None and K <= T:
        assert torch.equal(out, ref.forward(inp.cuda(), seg.cuda()))
        want = ref.backward(gout.cuda(), inp.cuda(), seg.cuda()).float()
        got = ops.bmp_backward(gout.cuda(), inp.cuda(), seg.cuda(), True).float()
        assert float((got - want).norm() / want.norm().clamp(min=1e-6)) < 5e-3 and float((got - want).abs().max()) < 0.25


def test_error_behaviour(ops):
    inp, seg, gout = (t.cuda() for t in make_case(2, 8, 16, 4, seed=5))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.bmp_forward(inp.cpu(), seg)
    with pytest.raises(RuntimeError, match="contiguous"):
        ops.bmp_forward(inp.transpose(1, 2), seg)
    with pytest.raises(RuntimeError):
        ops.bmp_forward(inp, seg[:1])                      # D11: segments batch must match
    with pytest.raises(RuntimeError):
        ops.bmp_forward(inp[:, :7].contiguous(), seg)      # odd channel count
    with pytest.raises(RuntimeError):
        ops.bmp_backward(gout, inp[:, :, :3].contiguous(), seg, True)   # compat with K > T


def test_reference_python_api(ops):
    """Autograd glue identical to boundary_pooling_op.py: module without ctor args, (grad_input, None)."""
    from opental_b200.prop_pooling import BoundaryMaxPooling, BoundaryMaxPoolingFunction
    inp, seg, gout = (t.cuda() for t in make_case(2, 8, 16, 16, seed=9))
    x = inp.clone().requires_grad_(True)
    y = BoundaryMaxPooling()(x, seg)
    y.backward(gout.transpose(1, 2).contiguous().transpose(1, 2))      # non-contiguous grad is accepted
    xo = inp.cpu().clone().requires_grad_(True)
    (gx,) = torch.autograd.grad(O.boundary_max_pooling(xo, seg.cpu(), False), xo, gout.cpu())
    assert torch.allclose(x.grad.cpu(), gx, atol=1e-5)
    assert BoundaryMaxPoolingFunction.compat_tscale_bug is False
    import opental_b200
    opental_b200.install_shim()
    import boundary_max_pooling_cuda as shim
    assert torch.equal(shim.forward(inp, seg), y.detach())
