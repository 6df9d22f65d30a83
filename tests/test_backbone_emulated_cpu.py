"""CPU: the product's backbone HOST code (opental_b200/backbone.py forward + explicit backward schedule, descriptor building in
opental_b200/ops.py) run against an emulation of the C ABI (tests/abi_emu.py) and compared with the oracle's I3D + torch
autograd.  The kernels themselves are checked on the GPU (tests/test_backbone_kernels_gpu.py, tests/test_model_gpu.py); this
test pins the bookkeeping around them — channel slices of the inception concat buffers, pads, the K-concatenated 1x1 data
gradient, the fused stage-pool backward, arg-max plumbing — on every CPU run, and exercises the STAGED raw-uint8 Conv3d_1a
path end to end through the same schedule."""
import pytest
import torch

import abi_emu
import opental_oracle as O


def build(monkeypatch):
    from opental_b200.backbone import I3DBackbone
    emu = abi_emu.install(monkeypatch)
    net = I3DBackbone()
    sd = {k[len("backbone."):]: v for k, v in O.synthetic_state_dict(O.OracleConfig()).items() if k.startswith("backbone.")}
    net.load_state_dict(sd)
    net.train()
    return net, emu


def oracle_run(x, g4, g5):
    sd = {k: v.clone() for k, v in O.synthetic_state_dict(O.OracleConfig()).items() if k.startswith("backbone.")}
    ws = {k: v.requires_grad_(True) for k, v in sd.items() if k.endswith("conv3d.weight")}
    f = O.i3d_features(x, sd)
    (f["Mixed_4f"] * g4).sum().add((f["Mixed_5c"] * g5).sum()).backward()
    return f, {k: v.grad for k, v in ws.items()}


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def check(net, out, want, grads):
    """Forward: 2e-4 (the GPU bar).  Weight gradients: 5e-2 in L2 — the bar of the whole-model GPU tests: max-pool routing and ReLU masks make the gradient discontinuous, and the
    oracle's own early-layer gradients move by ~1e-2 when its input is perturbed by 1e-5 (hi + lo planes carry ~2^-17)."""
    assert rel(out["Mixed_4f"], want["Mixed_4f"].detach()) < 2e-4 and rel(out["Mixed_5c"], want["Mixed_5c"].detach()) < 2e-4
    params = dict(net.named_parameters())
    l2, mx = {}, {}
    for k, g in grads.items():
        got = params[k[len("backbone."):]].grad
        assert got is not None and got.shape == g.shape, k
        l2[k] = float((got - g).norm() / g.norm().clamp(min=1e-30))
        mx[k] = rel(got, g)
    assert len(l2) == 57
    bad = {k: v for k, v in l2.items() if v > 5e-2}
    assert not bad, bad
    # behind the last stage pool a layer is either exact or hit by one of those discrete events (an arg-max / mask flip in the
    # block's own 3x3x3 pool or ReLU): several of the 12 must be exact, whatever the input
    deep = {k: v for k, v in mx.items() if "Mixed_5" in k}
    assert len(deep) == 12 and sum(v < 1e-4 for v in deep.values()) >= 4, deep


def test_backbone_schedule_matches_oracle(monkeypatch):
    net, emu = build(monkeypatch)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 3, 32, 64, 64, generator=g) * 2 - 1
    g4, g5 = torch.randn(1, 832, 8, 4, 4, generator=g), torch.randn(1, 1024, 4, 2, 2, generator=g)
    want, grads = oracle_run(x, g4, g5)
    out = net(x)
    assert tuple(out["Mixed_4f"].shape) == (1, 832, 8, 4, 4) and tuple(out["Mixed_5c"].shape) == (1, 1024, 4, 2, 2)
    net.flat_parameters()[1].zero_()
    (out["Mixed_4f"] * g4).sum().add((out["Mixed_5c"] * g5).sum()).backward()
    check(net, out, want, grads)
    # the schedule's launch counts: 57 convs (1 folded + 56), 13 pools; b1a + b2a of the 9 inception blocks share one forward
    # and one weight-gradient launch (fuse_b12a, the default): 56 - 9 = 47
    assert emu.calls["otal_conv1a_fwd"] == 1 and emu.calls["otal_conv1a_wgrad"] == 1 and emu.calls["otal_maxpool_fwd"] == 13
    assert emu.calls["otal_conv_wgrad"] == (47 if net.fuse_b12a else 56)
    assert "otal_conv1a_fwd_u8" not in emu.calls                       # the staged path is off by default


@pytest.mark.parametrize("flip", [0, 1])
def test_staged_uint8_conv1a_path_through_the_schedule(monkeypatch, flip):
    """uint8 frames -> raw-pixel plane -> otal_conv1a_fwd_u8 with the border-class shift table -> ... -> otal_conv1a_wgrad_u8 +
    class sums -> the reference's weight gradient: the host algebra wired into backbone.py, against the oracle fed with the
    loader's normalised clip."""
    from opental_b200 import dataset as D
    net, emu = build(monkeypatch)
    net.u8_conv1a = True
    net.crop_size = 64
    g = torch.Generator().manual_seed(4)
    px = torch.randint(0, 256, (1, 32, 72, 72, 3), generator=g, dtype=torch.uint8)
    net.crop_offsets = torch.tensor([[3, 5, flip]], dtype=torch.int32)
    x = D.host_clip(px[0].numpy(), (3, 5, flip), 64).unsqueeze(0)
    g4, g5 = torch.randn(1, 832, 8, 4, 4, generator=g), torch.randn(1, 1024, 4, 2, 2, generator=g)
    want, grads = oracle_run(x, g4, g5)
    out = net(px)
    net.flat_parameters()[1].zero_()
    (out["Mixed_4f"] * g4).sum().add((out["Mixed_5c"] * g5).sum()).backward()
    check(net, out, want, grads)
    assert emu.calls["otal_clip_ingest_u8_raw"] == 1 and emu.calls["otal_conv1a_fwd_u8_halo"] == 1
    assert emu.calls["otal_conv1a_wgrad_u8_halo"] == 1 and "otal_border_class_sums" not in emu.calls     # R comes from the ones slot
    assert "otal_conv1a_fwd" not in emu.calls and "otal_clip_ingest_u8" not in emu.calls


def test_uint8_frames_normalised_ingest_path(monkeypatch):
    """uint8 frames through the normalising ingest kernel (u8_conv1a off: hi + lo planes of the normalised clip) equal the
    loader's fp32 clip; the default raw-pixel path is test_staged_uint8_conv1a_path_through_the_schedule."""
    from opental_b200 import dataset as D
    net, emu = build(monkeypatch)
    net.crop_size = 64
    net.u8_conv1a = False
    g = torch.Generator().manual_seed(5)
    px = torch.randint(0, 256, (1, 32, 72, 72, 3), generator=g, dtype=torch.uint8)
    x = D.host_clip(px[0].numpy(), (4, 4, 0), 64).unsqueeze(0)
    with torch.no_grad():
        a, b = net(px), net(x)
    assert rel(a["Mixed_5c"], b["Mixed_5c"]) < 1e-6 and emu.calls["otal_clip_ingest_u8"] == 1 and emu.calls["otal_clip_ingest"] == 1


def test_staged_fused_bottleneck_convs_through_the_schedule(monkeypatch):
    """OTAL_FUSE_B12A: b1a + b2a of every inception block as one forward and one weight-gradient launch — same results, 18 launches
    fewer (9 forward, 9 weight gradient)."""
    net, emu = build(monkeypatch)
    net.fuse_b12a = True
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 3, 32, 64, 64, generator=g) * 2 - 1
    g4, g5 = torch.randn(1, 832, 8, 4, 4, generator=g), torch.randn(1, 1024, 4, 2, 2, generator=g)
    want, grads = oracle_run(x, g4, g5)
    out = net(x)
    net.flat_parameters()[1].zero_()
    (out["Mixed_4f"] * g4).sum().add((out["Mixed_5c"] * g5).sum()).backward()
    check(net, out, want, grads)
    assert emu.calls["otal_conv_wgrad"] == 56 - 9
    # forward launches: 56 generic convs without the switch, 47 with it
    for flag, want_launches in ((False, 56), (True, 47)):
        net.fuse_b12a = flag
        before = emu.calls["otal_conv_igemm_fwd"]
        with torch.no_grad():
            net(x)
        assert emu.calls["otal_conv_igemm_fwd"] - before == want_launches
