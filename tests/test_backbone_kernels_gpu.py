"""GPU parity of the backbone training kernels (through the C ABI) vs the CPU oracle's Unit3D / MaxPool3dSamePadding
and torch autograd of the same fp32 ops: strided conv, folded Conv3d_1a, dgrad, wgrad, max-pool fwd/bwd, ReLU/BN
backward split, fused Adam.  Tolerance 1e-4 relative (max-norm) in bf16x3 mode — inside BASELINE's 1e-3 budget."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import opental_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def to_planes(x_ncdhw):
    from opental_b200 import ops
    return ops.split_bf16(x_ncdhw.permute(0, 2, 3, 4, 1).contiguous().cuda())


def from_ndhwc(t):
    return t.permute(0, 4, 1, 2, 3).cpu()


def bn_sd(Cout, g):
    return {"u.bn.weight": 1 + 0.1 * torch.randn(Cout, generator=g), "u.bn.bias": 0.1 * torch.randn(Cout, generator=g),
            "u.bn.running_mean": 0.1 * torch.randn(Cout, generator=g),
            "u.bn.running_var": 1 + 0.1 * torch.randn(Cout, generator=g).abs()}


def fold_bn(sd):
    inv = torch.rsqrt(sd["u.bn.running_var"] + O.BN_EPS)
    return (sd["u.bn.weight"] * inv).cuda(), (sd["u.bn.bias"] - sd["u.bn.running_mean"] * sd["u.bn.weight"] * inv).cuda()


@pytest.mark.parametrize("case", [
    # N, Cin, Cout, T, H, W, k, s
    (2, 512, 512, 32, 1, 1, (3, 1, 1), (2, 1, 1)),      # pyramid Unit1D k3 s2 (layers.py:198-210 pads (0,1))
    (1, 16, 32, 9, 10, 12, (3, 3, 3), (2, 2, 2)),       # generic 3-D stride 2, odd extent
    (1, 8, 16, 8, 8, 8, (1, 3, 3), (1, 2, 2)),
])
def test_strided_conv(case):
    from opental_b200 import ops
    N, Cin, Cout, T, H, W, k, s = case
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, Cin, T, H, W, generator=g)
    w = torch.randn(Cout, Cin, *k, generator=g) * (2.0 / (Cin * k[0] * k[1] * k[2])) ** 0.5
    ref = F.conv3d(O._pad3d(x, k, s), w, stride=s)
    pads = tuple(O.same_pad(sz, kk, ss)[0] for sz, kk, ss in zip((T, H, W), k, s))
    y = ops.conv_igemm(to_planes(x), ops.pack_conv_weight(w.cuda()), kernel=k, pad_front=pads, stride=s)
    assert tuple(y.hi.shape) == (N, *ref.shape[2:], Cout)
    assert rel(from_ndhwc(y.float()), ref) < TOL


@pytest.mark.parametrize("shape", [(1, 16, 24, 24), (2, 12, 16, 32)])
def test_conv1a_folded(shape):
    """Conv3d_1a_7x7 (k7 s2, 3 channels) through clip_ingest + the folded-window kernel vs the oracle Unit3D."""
    from opental_b200 import ops
    N, T, H, W = shape
    g = torch.Generator().manual_seed(2)
    x = torch.rand(N, 3, T, H, W, generator=g) * 2 - 1
    sd = {"u.conv3d.weight": torch.randn(64, 3, 7, 7, 7, generator=g) * (2.0 / (3 * 343)) ** 0.5, **bn_sd(64, g)}
    ref = O.unit3d_bn_relu(x, sd, "u.", (7, 7, 7), (2, 2, 2))
    scale, shift = fold_bn(sd)
    xp = ops.clip_ingest(x.cuda())
    wp = ops.pack_conv1a_weight(sd["u.conv3d.weight"].cuda())
    y = ops.conv1a_fwd(xp, wp, W, scale=scale, shift=shift, relu=True)
    assert rel(from_ndhwc(y.float()), ref) < TOL


CONV_BWD_CASES = [
    # N, Cin, Cout, T, H, W, k
    (1, 64, 192, 8, 12, 12, (3, 3, 3)),      # Conv3d_2c slab
    (2, 96, 208, 4, 6, 6, (3, 3, 3)),        # Mixed_4b.b1b: N = 96 (dgrad), ragged
    (1, 480, 192, 4, 6, 6, (1, 1, 1)),       # Mixed_4b.b0: dgrad N = 480 -> 2 N blocks
    (1, 16, 48, 5, 6, 6, (3, 3, 3)),         # narrow channels, odd T
    (2, 512, 512, 64, 1, 1, (3, 1, 1)),      # Unit1D k3
]


@pytest.mark.parametrize("case", CONV_BWD_CASES)
def test_conv_dgrad_wgrad(case):
    from opental_b200 import ops
    N, Cin, Cout, T, H, W, k = case
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, Cin, T, H, W, generator=g).requires_grad_(True)
    w = (torch.randn(Cout, Cin, *k, generator=g) * (2.0 / (Cin * k[0] * k[1] * k[2])) ** 0.5).requires_grad_(True)
    gy = torch.randn(N, Cout, T, H, W, generator=g)
    y = F.conv3d(O._pad3d(x, k, (1, 1, 1)), w)
    gx_ref, gw_ref = torch.autograd.grad(y, (x, w), gy)
    pads = tuple(O.same_pad(sz, kk, 1)[0] for sz, kk in zip((T, H, W), k))
    wp = ops.pack_conv_weight(w.detach().cuda())
    dp = to_planes(gy)
    # dgrad: same kernel on the output gradient, forward weights read transposed with flipped taps
    gx = torch.full((N, T, H, W, Cin), 3.0, device="cuda")
    ops.conv_igemm(dp, wp, kernel=k, pad_front=tuple(kk - 1 - p for kk, p in zip(k, pads)), out_f32=gx,
                   want_planes=False, dgrad=True)
    assert rel(from_ndhwc(gx), gx_ref) < TOL
    # accumulate mode adds to what is there
    ops.conv_igemm(dp, wp, kernel=k, pad_front=tuple(kk - 1 - p for kk, p in zip(k, pads)), out_f32=gx,
                   want_planes=False, dgrad=True, accumulate=True)
    assert rel(from_ndhwc(gx), 2 * gx_ref) < TOL
    # wgrad
    dw = torch.zeros(k[0] * k[1] * k[2], Cout, Cin, device="cuda")
    ops.conv_wgrad(to_planes(x.detach()), dp, dw, kernel=k, pad_front=pads)
    gw = dw.view(*k, Cout, Cin).permute(3, 4, 0, 1, 2).cpu()
    assert rel(gw, gw_ref) < TOL


def test_conv_bwd_channel_slices():
    """dgrad / wgrad on channel slices of wider buffers (inception concat layout)."""
    from opental_b200 import ops
    g = torch.Generator().manual_seed(4)
    N, T, H, W, k = 1, 4, 6, 6, (3, 3, 3)
    Cx, xo, Cin, Cd, do, Cout = 160, 32, 96, 256, 64, 128
    xf = torch.randn(N, Cx, T, H, W, generator=g)
    df = torch.randn(N, Cd, T, H, W, generator=g)
    x = xf[:, xo:xo + Cin].clone().requires_grad_(True)
    w = (torch.randn(Cout, Cin, *k, generator=g) * 0.05).requires_grad_(True)
    y = F.conv3d(O._pad3d(x, k, (1, 1, 1)), w)
    gx_ref, gw_ref = torch.autograd.grad(y, (x, w), df[:, do:do + Cout])
    wp = ops.pack_conv_weight(w.detach().cuda())
    dp, xp = to_planes(df), to_planes(xf)
    gx = torch.full((N, T, H, W, Cx), 5.0, device="cuda")
    ops.conv_igemm(dp, wp, kernel=k, pad_front=(1, 1, 1), in_slice=(do, Cout), out_f32=gx, out_slice=(xo, Cin),
                   want_planes=False, dgrad=True)
    got = from_ndhwc(gx)
    assert rel(got[:, xo:xo + Cin], gx_ref) < TOL
    mask = torch.ones(Cx, dtype=torch.bool); mask[xo:xo + Cin] = False
    assert bool((got[:, mask] == 5.0).all())
    dw = torch.zeros(27, Cout, Cin, device="cuda")
    ops.conv_wgrad(xp, dp, dw, kernel=k, pad_front=(1, 1, 1), in_slice=(xo, Cin), d_slice=(do, Cout))
    assert rel(dw.view(*k, Cout, Cin).permute(3, 4, 0, 1, 2).cpu(), gw_ref) < TOL


def test_conv1a_wgrad():
    from opental_b200 import ops
    g = torch.Generator().manual_seed(5)
    N, T, H, W = 2, 12, 16, 16
    x = torch.rand(N, 3, T, H, W, generator=g) * 2 - 1
    w = (torch.randn(64, 3, 7, 7, 7, generator=g) * 0.03).requires_grad_(True)
    y = F.conv3d(O._pad3d(x, (7, 7, 7), (2, 2, 2)), w, stride=2)
    gy = torch.randn(y.shape, generator=g)
    (gw_ref,) = torch.autograd.grad(y, w, gy)
    dw = torch.zeros(49, 64, 32, device="cuda")
    ops.conv1a_wgrad(ops.clip_ingest(x.cuda()), to_planes(gy), dw, W)
    assert rel(ops.unpack_conv1a_wgrad(dw).cpu(), gw_ref) < TOL
    # the folded slots that carry no weight (8th W tap, channels 3..7) are not part of the parameter
    assert dw.shape == (49, 64, 32)


POOLS = [
    # C, T, H, W, k, s
    (64, 6, 16, 16, (1, 3, 3), (1, 2, 2)),     # MaxPool3d_2a / 3a
    (48, 8, 12, 12, (3, 3, 3), (2, 2, 2)),     # MaxPool3d_4a
    (64, 8, 6, 6, (2, 2, 2), (2, 2, 2)),       # MaxPool3d_5a
    (32, 5, 6, 6, (3, 3, 3), (1, 1, 1)),       # inception b3a
    (832, 4, 3, 3, (2, 2, 2), (2, 2, 2)),      # odd extent with k2 s2 (pad (0,1))
]


@pytest.mark.parametrize("case", POOLS)
def test_maxpool_fwd_bwd(case):
    from opental_b200 import ops
    C, T, H, W, k, s = case
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, C, T, H, W, generator=g).relu()          # post-ReLU like every pool input in I3D
    xp = to_planes(x)
    xr = from_ndhwc(xp.float()).requires_grad_(True)             # the represented values (hi + lo)
    ref = O.maxpool3d_same(xr, k, s)
    pads = tuple(O.same_pad(sz, kk, ss)[0] for sz, kk, ss in zip((T, H, W), k, s))
    y = ops.maxpool_fwd(xp, kernel=k, stride=s, pad_front=pads)
    assert torch.equal(from_ndhwc(y.float()), ref.detach())      # pure selection: bit exact
    gy = torch.randn(ref.shape, generator=g)
    (gx_ref,) = torch.autograd.grad(ref, xr, gy)
    gin = torch.zeros(2, T, H, W, C, device="cuda")
    ops.maxpool_bwd(xp, gy.permute(0, 2, 3, 4, 1).contiguous().cuda(), gin, kernel=k, stride=s, pad_front=pads)
    got = from_ndhwc(gin)
    # gradient routed to zero-valued inputs is killed by the ReLU mask upstream; compare where the input is positive
    pos = xr.detach() > 0
    assert torch.allclose(got[pos], gx_ref[pos], atol=1e-5, rtol=1e-5)
    # (an all-zero window may hand its gradient to the zero padding instead of a zero input: same effect)
    # recorded arg-max path: same output, and a pure-scatter backward that does not read x
    y2, arg = ops.maxpool_fwd(xp, kernel=k, stride=s, pad_front=pads, save_argmax=True)
    assert torch.equal(y2.hi, y.hi) and torch.equal(y2.lo, y.lo) and arg.dtype == torch.uint8
    gin2 = torch.zeros(2, T, H, W, C, device="cuda")
    ops.maxpool_bwd(xp, gy.permute(0, 2, 3, 4, 1).contiguous().cuda(), gin2, kernel=k, stride=s, pad_front=pads, argmax=arg)
    assert torch.allclose(from_ndhwc(gin2)[pos], gx_ref[pos], atol=1e-5, rtol=1e-5)
    assert abs(float(gin2.sum()) - float(gin.sum())) <= 1e-3 * float(gin.abs().sum())


def test_maxpool_signed_inputs_exact():
    """The integer-key comparison is exact for any sign (the I3D pools only ever see post-ReLU inputs)."""
    from opental_b200 import ops
    g = torch.Generator().manual_seed(16)
    x = torch.randn(1, 16, 4, 6, 6, generator=g) * 3
    xp = to_planes(x)
    xr = from_ndhwc(xp.float())
    for k, s in (((3, 3, 3), (1, 1, 1)), ((2, 2, 2), (2, 2, 2))):
        pads = tuple(O.same_pad(sz, kk, ss)[0] for sz, kk, ss in zip((4, 6, 6), k, s))
        y = ops.maxpool_fwd(xp, kernel=k, stride=s, pad_front=pads)
        assert torch.equal(from_ndhwc(y.float()), O.maxpool3d_same(xr, k, s))


def test_relu_bn_bwd_split():
    from opental_b200 import ops
    g = torch.Generator().manual_seed(7)
    N, T, H, W, C = 2, 3, 4, 5, 48
    gy = torch.randn(N, T, H, W, 64, generator=g).cuda()
    y = torch.randn(N, T, H, W, 56, generator=g).relu().cuda()
    scale = (1 + 0.1 * torch.randn(C, generator=g)).cuda()
    yp = ops.split_bf16(y)
    d = ops.relu_bn_bwd_split(gy, yp, scale, g_slice=(8, C), y_slice=(8, C))
    ref = gy[..., 8:8 + C] * (yp.float()[..., 8:8 + C] > 0) * scale
    assert rel(d.float().cpu(), ref.cpu()) < 1e-5


def test_fused_adam_matches_torch():
    from opental_b200 import ops
    g = torch.Generator().manual_seed(8)
    p0 = torch.randn(10007, generator=g)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p_ref], lr=1e-3, weight_decay=1e-3)
    p = p0.clone().cuda(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step in range(1, 4):
        grad = torch.randn(10007, generator=g)
        p_ref.grad = grad.clone()
        opt.step()
        ops.adam_step(p, (grad * 4).cuda(), m, v, lr=1e-3, weight_decay=1e-3, grad_scale=0.25, step=step)
    assert torch.allclose(p.cpu(), p_ref.detach(), atol=1e-6, rtol=1e-5)


HEAD_CONVS = [
    # B, Cin, Cout, T, k, stride
    (2, 512, 512, 64, 3, 1),       # tower conv
    (2, 512, 2, 16, 3, 1),         # loc_head: Cout padded to 8
    (3, 512, 15, 8, 1, 1),         # prop_conf_head
    (2, 512, 1, 2, 3, 1),          # actionness head on the shortest level
    (2, 512, 512, 32, 3, 2),       # pyramid Unit1D k3 s2 (pads (0,1)), strided wgrad + upsampled dgrad
    (1, 512, 512, 4, 3, 2),
    (2, 2048, 512, 16, 1, 1),      # proposal_conv
    (1, 512, 1024, 64, 1, 1),      # lr_conv: 4 N blocks
]


@pytest.mark.parametrize("case", HEAD_CONVS)
def test_head_conv1d_fwd_bwd(case):
    """Unit1D on the tensor-core kernels (headconv.py) vs torch's fp32 conv1d + autograd (the oracle's unit1d)."""
    from opental_b200.headconv import HeadConvStore, head_conv
    B, Cin, Cout, T, k, s = case
    g = torch.Generator().manual_seed(11)
    conv = torch.nn.Conv1d(Cin, Cout, k, s).cuda()
    with torch.no_grad():
        conv.weight.copy_(torch.randn(Cout, Cin, k, generator=g) * (2.0 / (Cin * k)) ** 0.5)
        conv.bias.copy_(0.1 * torch.randn(Cout, generator=g))
    w_ref = conv.weight.detach().cpu().clone().requires_grad_(True)
    b_ref = conv.bias.detach().cpu().clone().requires_grad_(True)
    x = torch.randn(B, Cin, T, generator=g)
    x_ref = x.clone().requires_grad_(True)
    y_ref = O.unit1d(x_ref, {"u.conv1d.weight": w_ref, "u.conv1d.bias": b_ref}, "u.", stride=s)
    gy = torch.randn(y_ref.shape, generator=g)
    gx_ref, gw_ref, gb_ref = torch.autograd.grad(y_ref, (x_ref, w_ref, b_ref), gy)

    store = HeadConvStore()
    rec = store.register(conv.weight, conv.bias, "conv1d")
    store.prepare(torch.device("cuda", torch.cuda.current_device()))
    store.bind_grads()
    xg = x.cuda().requires_grad_(True)
    y = head_conv(xg, conv.weight, conv.bias, store, rec, s)
    assert tuple(y.shape) == tuple(y_ref.shape)
    assert rel(y.detach().cpu(), y_ref.detach()) < TOL
    y.backward(gy.cuda())
    assert rel(xg.grad.cpu(), gx_ref) < TOL
    assert rel(conv.weight.grad.cpu(), gw_ref) < TOL
    assert rel(conv.bias.grad.cpu(), gb_ref) < TOL
    # parameters still have the reference's shapes and alias the packed store
    assert tuple(conv.weight.shape) == (Cout, Cin, k)
    assert torch.equal(conv.weight.detach().cpu(), w_ref.detach())


def test_head_valid3d_conv():
    """pyramids.0: Conv3d [512,832,1,6,6] 'spatial_valid' on the backbone's channels-last feature map."""
    from opental_b200.headconv import HeadConvStore, head_conv
    g = torch.Generator().manual_seed(12)
    B, C, T, Hh = 2, 832, 8, 6
    conv = torch.nn.Conv3d(C, 512, (1, Hh, Hh)).cuda()
    with torch.no_grad():
        conv.weight.copy_(torch.randn(512, C, 1, Hh, Hh, generator=g) * (2.0 / (C * Hh * Hh)) ** 0.5)
        conv.bias.copy_(0.1 * torch.randn(512, generator=g))
    w_ref = conv.weight.detach().cpu().clone().requires_grad_(True)
    b_ref = conv.bias.detach().cpu().clone().requires_grad_(True)
    x = torch.randn(B, C, T, Hh, Hh, generator=g).relu()
    x_ref = x.clone().requires_grad_(True)
    y_ref = O.head_unit3d_valid(x_ref, {"u.conv3d.weight": w_ref, "u.conv3d.bias": b_ref}, "u.").squeeze(-1).squeeze(-1)
    gy = torch.randn(y_ref.shape, generator=g)
    gx_ref, gw_ref, gb_ref = torch.autograd.grad(y_ref, (x_ref, w_ref, b_ref), gy)
    store = HeadConvStore()
    rec = store.register(conv.weight, conv.bias, "valid3d")
    store.prepare(torch.device("cuda", torch.cuda.current_device()))
    store.bind_grads()
    xl = x.permute(0, 2, 3, 4, 1).contiguous().cuda().requires_grad_(True)        # NDHWC storage
    y = head_conv(xl, conv.weight, conv.bias, store, rec, 1)
    # reduction length 29 952: bf16x3 drops the lo*lo products (2^-16 each, random-walk ~1e-4 of the max) and the
    # fp32 CPU reference carries its own summation error of the same order; 3e-4 is still 3x inside the budget
    assert rel(y.detach().cpu(), y_ref.detach()) < 3e-4
    y.backward(gy.cuda())
    assert rel(xl.grad.permute(0, 4, 1, 2, 3).cpu(), gx_ref) < TOL
    assert rel(conv.weight.grad.cpu(), gw_ref) < TOL
    assert rel(conv.bias.grad.cpu(), gb_ref) < TOL


def test_clip_ingest_u8_bit_exact():
    """uint8 frames -> planes in one kernel == the data loader's crop / flip / (x/255)*2-1 followed by the fp32 ingest."""
    from opental_b200 import ops
    g = torch.Generator().manual_seed(31)
    N, T, Hs, Ws, crop = 3, 6, 28, 30, 24
    px = torch.randint(0, 256, (N, T, Hs, Ws, 3), generator=g, dtype=torch.uint8)

    def loader(px, oh, ow, flip):
        x = px[:, oh:oh + crop, ow:ow + crop, :]
        if flip:
            x = x.flip(2)
        return (x.permute(3, 0, 1, 2).float() / 255.0) * 2.0 - 1.0          # thumos_dataset.py:261-263

    # centre crop, no mirroring (the default)
    ref = ops.clip_ingest(torch.stack([loader(px[n], (Hs - crop) // 2, (Ws - crop) // 2, False) for n in range(N)]).cuda())
    got = ops.clip_ingest_u8(px.cuda(), crop)
    assert torch.equal(got.hi, ref.hi) and torch.equal(got.lo, ref.lo)
    # per-sample random crop + mirror flag
    offs = torch.tensor([[0, 0, 0], [4, 6, 1], [2, 3, 1]], dtype=torch.int32)
    ref = ops.clip_ingest(torch.stack([loader(px[n], int(offs[n, 0]), int(offs[n, 1]), bool(offs[n, 2])) for n in range(N)]).cuda())
    got = ops.clip_ingest_u8(px.cuda(), crop, offs.cuda())
    assert torch.equal(got.hi, ref.hi) and torch.equal(got.lo, ref.lo)


def test_clip_ingest_u8_frame_map_is_a_temporal_gather():
    """SSL cut-paste (thumos_dataset.py:187-229) as a frame map in the ingest kernel == the same ingest followed by an
    index_select along time; together with crop offsets / mirroring; bit-exact."""
    import random
    from opental_b200 import augment, ops
    g = torch.Generator().manual_seed(32)
    N, T, Hs, Ws, crop = 3, 256, 20, 22, 16
    px = torch.randint(0, 256, (N, T, Hs, Ws, 3), generator=g, dtype=torch.uint8).cuda()
    offs = torch.tensor([[0, 0, 0], [4, 6, 1], [2, 3, 1]], dtype=torch.int32).cuda()
    annos = [[[30, 120, 3], [170, 230, 7]], [[10, 20, 1]], [[60.5, 140.25, 2]]]
    maps = [augment.cut_paste(a, 8, T, 1, rng=random.Random(i))[0] for i, a in enumerate(annos)]
    assert (maps[1] == np.arange(T)).all() and (maps[0] != np.arange(T)).any() and (maps[2] != np.arange(T)).any()
    fmap = torch.from_numpy(np.stack(maps)).cuda()
    plain = ops.clip_ingest_u8(px, crop, offs)
    got = ops.clip_ingest_u8(px, crop, offs, frame_map=fmap)
    idx = fmap.long().view(N, T, 1, 1, 1).expand_as(plain.hi)
    assert torch.equal(got.hi, plain.hi.gather(1, idx)) and torch.equal(got.lo, plain.lo.gather(1, idx))
    # out-of-range entries are clamped (memory safety), never read outside the clip
    bad = fmap.clone()
    bad[0, 0], bad[0, 1] = -5, T + 7
    got = ops.clip_ingest_u8(px, crop, offs, frame_map=bad)
    assert torch.equal(got.hi[0, 0], plain.hi[0, 0]) and torch.equal(got.hi[0, 1], plain.hi[0, T - 1])


def test_backbone_accepts_uint8_frames():
    from opental_b200 import engine
    net, _ = engine.build_opental(epoch=1)
    px = engine.synthetic_clip_u8(0, frames=64).unsqueeze(0).cuda()
    x = engine.normalise_clip(px[0].cpu()).unsqueeze(0).cuda()
    with torch.no_grad():
        b = net.backbone(x)
        net.backbone.u8_conv1a = False       # normalised planes from the ingest kernel: the same bits as the fp32 clip's planes
        a = net.backbone(px)
        assert torch.equal(a["Mixed_5c"], b["Mixed_5c"]) and torch.equal(a["Mixed_4f"], b["Mixed_4f"])
        net.backbone.u8_conv1a = True        # default: Conv3d_1a on the raw pixel values (different rounding, same accuracy class)
        c = net.backbone(px)
    for k in ("Mixed_4f", "Mixed_5c"):
        assert float((c[k] - b[k]).abs().max() / b[k].abs().max()) < 1e-4, k
