"""CPU: the host-side algebra of the raw-uint8 Conv3d_1a path (ops.conv1a_u8_scale_shift / conv1a_u8_weight_grad /
border classes) against torch's conv3d on the normalised, zero-padded clip — i.e. against the reference's Unit3D
(AFSD/common/i3d_backbone.py:51-87) fed by the loader's normalisation (thumos_dataset.py:261-263).  The tensor-core kernels
are stood in for by fp32 conv3d on the raw pixel values (what otal_conv1a_fwd_u8 / otal_conv1a_wgrad_u8 accumulate); the GPU
tests (tests/test_conv1a_u8_gpu.py, opt-in) check the kernels themselves."""
import pytest
import torch
import torch.nn.functional as F

from opental_b200 import ops

PAD = (2, 3, 2, 3, 2, 3)       # TF "same" padding of k=7, s=2 on even extents: front 2, back 3 (i3d_backbone.py:45-69)


def make(T=12, H=16, W=20, cout=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    u = torch.randint(0, 256, (2, 3, T, H, W), generator=g).double()
    x = (u / 255.0) * 2.0 - 1.0
    w = torch.randn(cout, 3, 7, 7, 7, generator=g, dtype=torch.double) * (2.0 / 1029) ** 0.5
    scale = 1 + 0.1 * torch.randn(cout, generator=g, dtype=torch.double)
    shift = 0.1 * torch.randn(cout, generator=g, dtype=torch.double)
    return u, x, w, scale, shift


@pytest.mark.parametrize("n", [3, 4, 5, 6, 48, 128])
def test_border_classes_are_the_in_bounds_patterns(n):
    """class -> mask must equal the brute-force in-bounds pattern of input index 2*o + d - 2 in [0, 2n)."""
    m = ops.border_class_masks()
    cls = ops.border_classes(n)
    for o in range(n):
        want = torch.tensor([1.0 if 0 <= 2 * o + d - 2 < 2 * n else 0.0 for d in range(7)])
        assert torch.equal(m[cls[o]], want), (n, o)


@pytest.mark.parametrize("shape", [(12, 16, 20), (6, 6, 6), (8, 10, 6)])
def test_forward_identity(shape):
    u, x, w, scale, shift = make(*shape)
    ref = F.relu(F.conv3d(F.pad(x, PAD), w, stride=2) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    sc, tab = ops.conv1a_u8_scale_shift(w, scale, shift)
    acc = F.conv3d(F.pad(u, PAD), w, stride=2)                     # what the kernel accumulates: raw pixels, zero outside
    To, Ho, Wo = acc.shape[2:]
    ct, ch, cw = ops.border_classes(To), ops.border_classes(Ho), ops.border_classes(Wo)
    sh = tab[ct][:, ch][:, :, cw].permute(3, 0, 1, 2)              # [Cout,To,Ho,Wo]: the epilogue's table lookup
    got = F.relu(acc * sc.view(1, -1, 1, 1, 1) + sh[None])
    assert tuple(tab.shape) == (4, 4, 4, w.shape[0])
    assert float((got - ref).abs().max()) < 1e-12 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("shape", [(12, 16, 20), (6, 6, 6)])
def test_weight_gradient_identity(shape):
    u, x, w, _, _ = make(*shape, seed=1)
    g = torch.Generator().manual_seed(5)
    wr = w.clone().requires_grad_(True)
    out = F.conv3d(F.pad(x, PAD), wr, stride=2)
    d = torch.randn(out.shape, generator=g, dtype=torch.double)
    (ref,) = torch.autograd.grad(out, wr, d)
    wu = torch.zeros_like(w).requires_grad_(True)
    (raw,) = torch.autograd.grad(F.conv3d(F.pad(u, PAD), wu, stride=2), wu, d)       # sum_p D[p] u[p + tap]
    # the kernel's folded layout [49 (dt,dh), Cout, 8 W taps x 4 slots]
    cout = w.shape[0]
    folded = torch.zeros(7, 7, cout, 8, ops.CLIP_CPAD, dtype=torch.double)
    folded[:, :, :, :7, :3] = raw.permute(2, 3, 0, 4, 1)
    folded[:, :, :, 7, :] = 123.0                                                    # junk in the slots that carry no weight
    folded[:, :, :, :, 3] = -7.0
    To, Ho, Wo = out.shape[2:]
    ct, ch, cw = ops.border_classes(To), ops.border_classes(Ho), ops.border_classes(Wo)
    sums = torch.zeros(4, 4, 4, cout, dtype=torch.double)
    dd = d.sum(0).permute(1, 2, 3, 0)                                                # [To,Ho,Wo,Cout]
    for a in range(4):
        for b in range(4):
            for c in range(4):
                sel = (ct == a)[:, None, None] & (ch == b)[None, :, None] & (cw == c)[None, None, :]
                sums[a, b, c] = dd[sel].sum(0)
    got = ops.conv1a_u8_weight_grad(folded.reshape(49, cout, 32), sums, 3)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 1e-11 * float(ref.abs().max())
    # default form: R read from the ones slot (channel slot 3 = in-image indicator) of the folded gradient itself
    ones = torch.zeros(1, 1, *u.shape[2:], dtype=torch.double); ones[...] = 1.0
    wo = torch.zeros(cout, 1, 7, 7, 7, dtype=torch.double, requires_grad=True)
    (r_ind,) = torch.autograd.grad(F.conv3d(F.pad(ones.expand(u.shape[0], 1, *u.shape[2:]), PAD), wo, stride=2), wo, d)
    folded[:, :, :, :7, 3] = r_ind[:, 0].permute(1, 2, 0, 3)
    got1 = ops.conv1a_u8_weight_grad(folded.reshape(49, cout, 32), None, 3)
    assert float((got1 - ref).abs().max()) < 1e-11 * float(ref.abs().max())


def test_pixel_values_are_exact_in_bf16():
    v = torch.arange(256, dtype=torch.float32)
    assert torch.equal(v.bfloat16().float(), v)


def test_raw_uint8_path_is_the_default_and_switchable(monkeypatch):
    from opental_b200.backbone import I3DBackbone
    monkeypatch.delenv("OTAL_U8_CONV1A", raising=False)
    assert I3DBackbone().u8_conv1a is True
    monkeypatch.setenv("OTAL_U8_CONV1A", "0")
    assert I3DBackbone().u8_conv1a is False
