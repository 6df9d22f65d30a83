"""GPU parity of the fused GroupNorm(32, C) + ReLU kernel (opental_b200/csrc/gn.cu, through the C ABI) against the oracle
formulation `relu(group_norm(x, 32, w, b, eps=1e-5))` (oracle/opental_oracle.py:gn_relu — torch CPU; here evaluated in
fp64 so the comparison measures OUR rounding, not the oracle's).  Tolerance 2e-6 absolute / 1e-5 relative forward and
input gradient, 1e-5 relative on the parameter gradients (sums over B*T terms)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [(8, 512, 64), (2, 1024, 32), (8, 512, 256), (3, 512, 2), (1, 1024, 4), (2, 64, 7), (1, 512, 1024)]


@pytest.mark.parametrize("B,C,T", SHAPES)
@pytest.mark.parametrize("relu", [True, False])
def test_groupnorm_relu_matches_fp64(B, C, T, relu):
    from opental_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + C + T)
    x = torch.randn(B, C, T, generator=g) * 2 + 0.5
    w = 1 + 0.1 * torch.randn(C, generator=g)
    b = 0.1 * torch.randn(C, generator=g)
    gy = torch.randn(B, C, T, generator=g)
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, w, b))
    ref = F.group_norm(xr, 32, wr, br, eps=1e-5)
    if relu:
        ref = ref.relu()
    gxr, gwr, gbr = torch.autograd.grad(ref, (xr, wr, br), gy.double())
    xd, wd, bd = (t.cuda().requires_grad_(True) for t in (x, w, b))
    y = ops.groupnorm_relu(xd, wd, bd, 32, 1e-5, relu)
    gx, gw, gb = torch.autograd.grad(y, (xd, wd, bd), gy.cuda())
    assert torch.allclose(y.detach().cpu().double(), ref.detach(), atol=4e-6, rtol=1e-5)
    # the ReLU mask may differ for outputs within rounding of zero; their gradient contribution is bounded by |gy|*tiny
    assert torch.allclose(gx.cpu().double(), gxr, atol=2e-5, rtol=1e-4), float((gx.cpu().double() - gxr).abs().max())
    assert torch.allclose(gw.cpu().double(), gwr, atol=1e-4 * float(gwr.abs().max()), rtol=1e-4)
    assert torch.allclose(gb.cpu().double(), gbr, atol=1e-4 * float(gbr.abs().max()), rtol=1e-4)


def test_groupnorm_module_keeps_reference_state_dict_keys():
    from opental_b200.bdnet import GroupNormReLU, Unit1D, _unit_gn
    seq = _unit_gn(Unit1D(512, 512, 3), 512)
    assert sorted(seq.state_dict()) == ["0.conv1d.bias", "0.conv1d.weight", "1.bias", "1.weight"]
    assert isinstance(seq[1], GroupNormReLU) and isinstance(seq[1], torch.nn.GroupNorm)
