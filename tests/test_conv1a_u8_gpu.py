"""GPU: the raw-uint8 Conv3d_1a path (the default for uint8 input since round 2) — otal_clip_ingest_u8_raw, otal_conv1a_fwd_u8,
otal_conv1a_wgrad_u8, otal_border_class_sums — vs the CPU oracle's Unit3D on the normalised clip and torch autograd.  Same
tolerance as the bf16x3 form it replaces (1e-4 relative, max-norm)."""
import os

import pytest
import torch
import torch.nn.functional as F

import opental_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def frames(N, T, Hs, Ws, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (N, T, Hs, Ws, 3), generator=g, dtype=torch.uint8)


def loader_clip(px, crop, offs=None):
    """[N,3,T,crop,crop] fp32: the data loader's crop / mirror / normalise (thumos_dataset.py:254-263)."""
    from opental_b200 import dataset as D
    N, T, Hs, Ws, _ = px.shape
    out = []
    for n in range(N):
        c = ((Hs - crop) // 2, (Ws - crop) // 2, 0) if offs is None else tuple(int(v) for v in offs[n])
        out.append(D.host_clip(px[n].numpy(), c, crop))
    return torch.stack(out)


def test_raw_ingest_is_exact():
    from opental_b200 import ops
    px = frames(2, 6, 20, 24, 0)
    offs = torch.tensor([[1, 3, 0], [2, 0, 1]], dtype=torch.int32)
    got = ops.clip_ingest_u8(px.cuda(), 16, offs.cuda(), raw=True)
    assert got.lo is None and tuple(got.hi.shape) == (2, 6, 16, 24, 4)
    want = torch.zeros(2, 6, 16, 24, 4)
    for n in range(2):
        i, j, flip = (int(v) for v in offs[n])
        x = px[n, :, i:i + 16, j:j + 16, :].float()
        want[n, :, :, 2:18, :3] = x.flip(2) if flip else x
        want[n, :, :, 2:18, 3] = 1.0            # the ones slot: in-image indicator (ops.conv1a_u8_weight_grad)
    assert torch.equal(got.hi.float().cpu(), want)


@pytest.mark.parametrize("shape", [(1, 16, 24, 24), (2, 12, 16, 32), (1, 8, 8, 16), (1, 6, 96, 96), (1, 8, 64, 40)])
def test_conv1a_fwd_u8(shape):
    from opental_b200 import ops
    N, T, H, W = shape
    px = frames(N, T, H, W, 2)
    g = torch.Generator().manual_seed(2)
    sd = {"u.conv3d.weight": torch.randn(64, 3, 7, 7, 7, generator=g) * (2.0 / (3 * 343)) ** 0.5,
          "u.bn.weight": 1 + 0.1 * torch.randn(64, generator=g), "u.bn.bias": 0.1 * torch.randn(64, generator=g),
          "u.bn.running_mean": 0.1 * torch.randn(64, generator=g), "u.bn.running_var": 1 + 0.1 * torch.randn(64, generator=g).abs()}
    x = loader_clip(px, W) if H == W else ((px.permute(0, 4, 1, 2, 3).float() / 255.0) * 2.0 - 1.0)
    ref = O.unit3d_bn_relu(x, sd, "u.", (7, 7, 7), (2, 2, 2))
    inv = torch.rsqrt(sd["u.bn.running_var"] + O.BN_EPS)
    scale, shift = (sd["u.bn.weight"] * inv).cuda(), (sd["u.bn.bias"] - sd["u.bn.running_mean"] * sd["u.bn.weight"] * inv).cuda()
    w = sd["u.conv3d.weight"].cuda()
    sc, tab = ops.conv1a_u8_scale_shift(w, scale, shift)
    if H == W:
        a = ops.clip_ingest_u8(px.cuda(), W, raw=True)
    else:       # non-square: build the raw plane by hand (the ingest kernel crops squares)
        hi = torch.zeros(N, T, H, W + 8, 4, dtype=torch.bfloat16).cuda()
        hi[:, :, :, 2:W + 2, :3] = px.cuda().to(torch.bfloat16)
        a = ops.Planes(hi, None)
    y = ops.conv1a_fwd(a, ops.pack_conv1a_weight(w), W, scale=sc, shift=tab, relu=True, u8=True)
    assert rel(y.float().permute(0, 4, 1, 2, 3).cpu(), ref) < TOL
    # the resident-halo kernel (the default when the packed [hi | lo] weights are passed): same operator, same tolerance, and
    # the same tensor-core products as the generic kernel in a different order (partial 16 x 8 tiles are clipped by TMA)
    wp = ops.pack_conv1a_weight(w)
    yh = ops.conv1a_fwd(a, wp, W, scale=sc, shift=tab, relu=True, u8=True, w_cat=ops.pack_conv1a_weight_cat(wp))
    assert rel(yh.float().permute(0, 4, 1, 2, 3).cpu(), ref) < TOL
    assert rel(yh.float().cpu(), y.float().cpu()) < 2e-5
    # and it agrees with the validated bf16x3 form on the normalised planes
    y3 = ops.conv1a_fwd(ops.clip_ingest(x.cuda()), ops.pack_conv1a_weight(w), W, scale=scale, shift=shift, relu=True)
    assert rel(y.float().cpu(), y3.float().cpu()) < TOL


def test_border_class_sums():
    from opental_b200 import ops
    g = torch.Generator().manual_seed(3)
    d = torch.randn(2, 7, 9, 11, 64, generator=g)
    got = ops.border_class_sums(ops.split_bf16(d.cuda())).cpu()
    ct, ch, cw = ops.border_classes(7), ops.border_classes(9), ops.border_classes(11)
    dd = d.double().sum(0)
    for a in range(4):
        for b in range(4):
            for c in range(4):
                sel = (ct == a)[:, None, None] & (ch == b)[None, :, None] & (cw == c)[None, None, :]
                want = dd[sel].sum(0)
                assert float((got[a, b, c].double() - want).abs().max()) < 1e-4 * max(1.0, float(want.abs().max())), (a, b, c)


@pytest.mark.parametrize("shape", [(2, 12, 16, 16), (1, 8, 16, 16), (1, 6, 96, 96), (1, 8, 24, 24)])
def test_conv1a_wgrad_u8(shape):
    from opental_b200 import ops
    N, T, H, W = shape
    px = frames(N, T, H, W, 5)
    x = loader_clip(px, W)
    g = torch.Generator().manual_seed(5)
    w = (torch.randn(64, 3, 7, 7, 7, generator=g) * 0.03).requires_grad_(True)
    y = F.conv3d(O._pad3d(x, (7, 7, 7), (2, 2, 2)), w, stride=2)
    gy = torch.randn(y.shape, generator=g)
    (gw_ref,) = torch.autograd.grad(y, w, gy)
    d = ops.split_bf16(gy.permute(0, 2, 3, 4, 1).contiguous().cuda())
    dw = torch.zeros(49, 64, 32).cuda()
    ops.conv1a_wgrad(ops.clip_ingest_u8(px.cuda(), W, raw=True), d, dw, W, u8=True)
    got = ops.conv1a_u8_weight_grad(dw, ops.border_class_sums(d), 3)          # R from the border-class sums (round-1 form)
    assert rel(got.cpu(), gw_ref) < TOL
    got1 = ops.conv1a_u8_weight_grad(dw, None, 3)                             # R from the ones slot of the raw plane (default)
    assert rel(got1.cpu(), gw_ref) < TOL
    # both kernels: the resident-halo one (the default) and the generic one must agree (same products, different order)
    dw2 = torch.zeros(49, 64, 32).cuda()
    ops.CONV1A_WGRAD_HALO = not ops.CONV1A_WGRAD_HALO
    try:
        ops.conv1a_wgrad(ops.clip_ingest_u8(px.cuda(), W, raw=True), d, dw2, W, u8=True)
    finally:
        ops.CONV1A_WGRAD_HALO = not ops.CONV1A_WGRAD_HALO
    assert rel(dw2.cpu(), dw.cpu()) < 2e-5


def test_model_step_matches_the_bf16x3_path():
    """One training step of the whole model from uint8 frames with and without the staged path: same losses and gradients
    to the bf16x3 tolerance (the two paths differ only in Conv3d_1a's arithmetic)."""
    from opental_b200 import engine
    torch.manual_seed(0)
    net, crit = engine.build_opental(epoch=1)          # before ibm_start: the loss holds no state that the first pass would change
    px = torch.stack([engine.synthetic_clip_u8(i) for i in range(2)]).cuda()
    tg = [engine.synthetic_targets(i).cuda() for i in range(2)]
    sc = torch.stack([engine.synthetic_scores(t.cpu()) for t in tg]).cuda()
    tr = engine.Trainer(net, crit)
    res = {}
    for flag in (False, True):
        net.backbone.u8_conv1a = flag
        tr.zero_grad()
        cost, losses, ls, le = tr.forward_backward(px, tg, sc)
        res[flag] = (float(cost), [float(v) for v in losses], net.backbone.convs["Conv3d_1a_7x7"].unit.conv3d.weight.grad.clone(),
                     tr.groups[0][1].clone())
    net.backbone.u8_conv1a = False
    assert abs(res[True][0] - res[False][0]) <= 1e-3 * abs(res[False][0])
    for a, b in zip(res[True][1], res[False][1]):
        assert abs(a - b) <= 1e-3 * max(1.0, abs(b))
    assert rel(res[True][2], res[False][2]) < 5e-2          # gradients are discontinuous in the activations (ReLU / arg-max flips)
    assert rel(res[True][3], res[False][3]) < 5e-2
