"""GPU parity: tcgen05 implicit-GEMM convolution (through the C ABI) vs the CPU oracle's Unit3D / Unit1D
(conv -> folded frozen BN -> ReLU) in fp32, tolerance 1e-4 relative (max-norm) in bf16x3 mode — one order of
magnitude inside the 1e-3 budget of BASELINE.json — and 2e-2 in single-pass bf16 mode."""
import pytest
import torch
import torch.nn.functional as F

import opental_oracle as O

pytestmark = pytest.mark.gpu

TOL_X3 = 1e-4
TOL_BF16 = 2e-2


def run_case(N, Cin, Cout, T, H, W, k, nsplit, seed=0, relu=True, cstride_in=None, in_off=0, out_total=None, out_off=0):
    from opental_b200 import ops
    g = torch.Generator().manual_seed(seed)
    Cx = cstride_in or Cin
    xfull = torch.randn(N, Cx, T, H, W, generator=g)
    w = torch.randn(Cout, Cin, *k, generator=g) * (2.0 / (Cin * k[0] * k[1] * k[2])) ** 0.5
    sd = {"u.conv3d.weight": w, "u.bn.weight": 1 + 0.1 * torch.randn(Cout, generator=g),
          "u.bn.bias": 0.1 * torch.randn(Cout, generator=g), "u.bn.running_mean": 0.1 * torch.randn(Cout, generator=g),
          "u.bn.running_var": 1 + 0.1 * torch.randn(Cout, generator=g).abs()}
    x = xfull[:, in_off:in_off + Cin]
    if relu:
        ref = O.unit3d_bn_relu(x, sd, "u.", k)
    else:
        ref = F.conv3d(O._pad3d(x, k, (1, 1, 1)), w)
    inv = torch.rsqrt(sd["u.bn.running_var"] + O.BN_EPS)
    scale = (sd["u.bn.weight"] * inv) if relu else None
    shift = (sd["u.bn.bias"] - sd["u.bn.running_mean"] * sd["u.bn.weight"] * inv) if relu else None
    xp = ops.split_bf16(xfull.permute(0, 2, 3, 4, 1).contiguous().cuda(), with_lo=nsplit == 3)
    wp = ops.pack_conv_weight(w.cuda(), with_lo=nsplit == 3)
    pads = tuple(O.same_pad(s, kk, 1)[0] for s, kk in zip((T, H, W), k))
    Ct = out_total or Cout
    hi = torch.full((N, T, H, W, Ct), 7.0, dtype=torch.bfloat16, device="cuda")
    out = ops.Planes(hi, torch.zeros_like(hi) if nsplit == 3 else None)
    f32 = torch.full((N, T, H, W, Ct), 7.0, device="cuda")
    ops.conv_igemm(xp, wp, kernel=k, pad_front=pads, scale=None if scale is None else scale.cuda(),
                   shift=None if shift is None else shift.cuda(), relu=relu, in_slice=(in_off, Cin), out=out,
                   out_slice=(out_off, Cout), out_f32=f32)
    torch.cuda.synchronize()
    refl = ref.permute(0, 2, 3, 4, 1)
    got_p = out.float().cpu()
    got_f = f32.cpu()
    sl = slice(out_off, out_off + Cout)
    denom = refl.abs().max()
    e_p = float((got_p[..., sl] - refl).abs().max() / denom)
    e_f = float((got_f[..., sl] - refl).abs().max() / denom)
    # channels outside the slice must be untouched
    mask = torch.ones(Ct, dtype=torch.bool); mask[sl] = False
    assert bool((got_f[..., mask] == 7.0).all()) and bool((got_p[..., mask] == 7.0).all())
    return e_p, e_f


CASES = [
    # N, Cin, Cout, T, H, W, kernel
    (1, 64, 64, 4, 8, 8, (1, 1, 1)),          # Conv3d_2b-like pointwise
    (2, 96, 208, 8, 12, 12, (3, 3, 3)),       # Mixed_4b.b1b channels, ragged N tile (208)
    (1, 24, 64, 5, 6, 6, (3, 3, 3)),          # Cin < 64 (zero-filled K chunk), odd T, partial tiles
    (2, 512, 512, 64, 1, 1, (3, 1, 1)),       # Unit1D k3 tower conv, 2 N-blocks
    (1, 832, 384, 16, 3, 3, (1, 1, 1)),       # Mixed_5c.b0, 13 K chunks, 3x3 spatial
    (1, 64, 192, 16, 24, 24, (3, 3, 3)),      # Conv3d_2c slab
    (1, 16, 32, 8, 12, 12, (3, 3, 3)),        # Mixed_3b.b2b: narrow channels
    (3, 512, 8, 32, 1, 1, (3, 1, 1)),         # a head conv padded to 8 output channels
]


@pytest.mark.parametrize("case", CASES)
def test_conv_bf16x3_vs_oracle(case):
    N, Cin, Cout, T, H, W, k = case
    e_p, e_f = run_case(N, Cin, Cout, T, H, W, k, nsplit=3)
    assert e_f < TOL_X3 and e_p < TOL_X3, (e_p, e_f)


@pytest.mark.parametrize("case", CASES[:3])
def test_conv_bf16_vs_oracle(case):
    N, Cin, Cout, T, H, W, k = case
    e_p, e_f = run_case(N, Cin, Cout, T, H, W, k, nsplit=1)
    assert e_f < TOL_BF16 and e_p < TOL_BF16, (e_p, e_f)


def test_conv_channel_slices_concat_buffer():
    """Reads a channel slice of a wider input and writes a slice of a wider (concat) output."""
    e_p, e_f = run_case(1, 96, 128, 8, 12, 12, (3, 3, 3), nsplit=3, cstride_in=256, in_off=64, out_total=256, out_off=64)
    assert e_f < TOL_X3 and e_p < TOL_X3


def test_conv_linear_no_epilogue():
    e_p, e_f = run_case(1, 64, 64, 4, 6, 6, (3, 3, 3), nsplit=3, relu=False)
    assert e_f < TOL_X3 and e_p < TOL_X3


def test_conv_linearity_full_size():
    """Size-independent property at the real Conv3d_2c size: conv(a*x) == a*conv(x) without epilogue (exact for a
    power of two), and the output of a zero input is zero."""
    from opental_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 128, 24, 24, 64, generator=g).cuda()
    w = (torch.randn(192, 64, 3, 3, 3, generator=g) * 0.03).cuda()
    wp = ops.pack_conv_weight(w)
    y1 = ops.conv_igemm(ops.split_bf16(x), wp, kernel=(3, 3, 3), pad_front=(1, 1, 1)).float()
    y2 = ops.conv_igemm(ops.split_bf16(x * 4.0), wp, kernel=(3, 3, 3), pad_front=(1, 1, 1)).float()
    assert torch.equal(y2, y1 * 4.0)
    y0 = ops.conv_igemm(ops.split_bf16(torch.zeros_like(x)), wp, kernel=(3, 3, 3), pad_front=(1, 1, 1)).float()
    assert float(y0.abs().max()) == 0.0


def test_conv_bad_arguments():
    from opental_b200 import ops
    x = ops.split_bf16(torch.randn(1, 2, 4, 4, 12).cuda())      # 12 channels: not a multiple of 8
    w = ops.pack_conv_weight(torch.randn(8, 12, 1, 1, 1).cuda())
    with pytest.raises(RuntimeError, match="multiples of 8"):
        ops.conv_igemm(x, w, kernel=(1, 1, 1), pad_front=(0, 0, 0))
