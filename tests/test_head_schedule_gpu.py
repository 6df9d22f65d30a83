"""GPU: the kernels of the explicit head schedule (opental_b200/head_schedule.py) against torch, and the schedule itself against the
per-module autograd formulation of the same head (which tests/test_head_gpu.py and the whole-model goldens pin to the oracle and
the reference).  Tolerances: 1e-5 relative for the fp32 glue / GroupNorm kernels, bf16 hi+lo round trip (2^-16) for planes, 1e-4
relative for results that went through tensor-core convolutions in a different accumulation order."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def planes_f32(p):
    return (p.hi.float() + (p.lo.float() if p.lo is not None else 0)).squeeze(2).squeeze(2)          # [B,T,C]


@pytest.mark.parametrize("segments", [None, ((1, 64), (66, 32), (99, 16), (116, 8), (125, 4), (130, 2))])
def test_groupnorm_ex_matches_torch(segments):
    from opental_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, C, T = 3, 1024, 136
    x = torch.randn(B, C, T, generator=g).cuda()
    gamma, beta = (1 + 0.1 * torch.randn(C, generator=g)).cuda(), (0.1 * torch.randn(C, generator=g)).cuda()
    segs = segments or ((0, T),)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = torch.zeros_like(x)
    for o, l in segs:
        ref[:, :, o:o + l] = F.relu(F.group_norm(xr[:, :, o:o + l], 32, gr, br, 1e-5))
    cat = ops.Planes(torch.zeros(B, T, 1, 1, C + 64, dtype=torch.bfloat16).cuda(), torch.zeros(B, T, 1, 1, C + 64, dtype=torch.bfloat16).cuda())
    y, pl, yt, stats = ops.groupnorm_relu_fwd_ex(x, gamma, beta, segments=segments, want_y=True, planes=cat, planes_coff=32, yt_range=(1, 64))
    assert rel(y, ref.detach()) < 1e-5
    got = planes_f32(cat)
    assert rel(got[:, :, 32:32 + C], ref.detach().permute(0, 2, 1)) < 2e-5
    assert float(got[:, :, :32].abs().max()) == 0 and float(got[:, :, 32 + C:].abs().max()) == 0
    assert torch.equal(yt, y[:, :, 1:65].permute(0, 2, 1))
    # backward: gy is a channel slice of a wider tensor, plus channels-last gradients of the two halves for columns [1, 65)
    wide = torch.randn(B, C + 48, T, generator=g).cuda()
    ga, gb = torch.randn(B, 64, C // 2, generator=g).cuda(), torch.randn(B, 64, C // 2, generator=g).cuda()
    gy = wide[:, 16:16 + C].clone()
    gy[:, :C // 2, 1:65] += ga.permute(0, 2, 1)
    gy[:, C // 2:, 1:65] += gb.permute(0, 2, 1)
    ref.backward(gy)
    dgamma, dbeta, dbias = torch.zeros(C).cuda(), torch.zeros(C).cuda(), torch.zeros(C).cuda()
    dpl, gx = ops.groupnorm_relu_bwd_ex(wide, x, gamma, beta, stats, dgamma=dgamma, dbeta=dbeta, dbias=dbias, segments=segments,
                                        gy_coff=16, want_gx=True, gy2=(ga, gb), gy2_off=1)
    assert rel(gx, xr.grad) < 2e-5 and rel(planes_f32(dpl), xr.grad.permute(0, 2, 1)) < 4e-5
    assert rel(dgamma, gr.grad) < 2e-5 and rel(dbeta, br.grad) < 2e-5 and rel(dbias, xr.grad.sum(dim=(0, 2))) < 1e-4
    # gradient only through the channels-last halves (no [B,C,T] gradient at all)
    dgamma.zero_(); dbeta.zero_()
    xr.grad = None
    ref2 = torch.zeros_like(x)
    for o, l in segs:
        ref2[:, :, o:o + l] = F.relu(F.group_norm(xr[:, :, o:o + l], 32, gamma, beta, 1e-5))
    only = torch.zeros_like(x)
    only[:, :C // 2, 1:65] = ga.permute(0, 2, 1)
    ref2.backward(only)
    dpl2, gx2 = ops.groupnorm_relu_bwd_ex(None, x, gamma, beta, stats, dgamma=dgamma, dbeta=dbeta, dbias=None, segments=segments,
                                          want_gx=True, gy2=(ga, None), gy2_off=1)
    assert rel(gx2, xr.grad) < 2e-5


def test_rows_combine_and_head_gather():
    from opental_b200 import ops
    g = torch.Generator().manual_seed(1)
    B, C = 2, 64
    srcs = [torch.randn(B, C, t, generator=g).cuda() for t in (8, 4, 16)]
    table = torch.tensor([[[0, j % 8], [1, j // 3], [-1, 0]] if j % 5 else [[-1, 0], [2, j], [2, 15 - j]] for j in range(12)], dtype=torch.int32).cuda()
    want = torch.zeros(B, C, 12).cuda()
    for j in range(12):
        for s, col in table[j].tolist():
            if s >= 0:
                want[:, :, j] += srcs[s][:, :, col]
    dst, pl = ops.rows_combine(srcs, table, want_f32=True, want_planes=True)
    assert torch.allclose(dst, want, atol=1e-6) and rel(planes_f32(pl), want.permute(0, 2, 1)) < 2e-5
    # head gather: ScaleExp head (2 of 8 channels) and a plain head (15 of 16), 10 priors in 14 columns, 3 levels
    S, P = 14, 10
    sep_idx = torch.tensor([1, 2, 3, 4, 5, 6, 8, 9, 10, 12], dtype=torch.int32).cuda()
    prior_of_col = torch.full((S,), -1, dtype=torch.int32)
    prior_of_col[sep_idx.cpu().long()] = torch.arange(P, dtype=torch.int32)
    level = torch.tensor([0] * 6 + [1] * 3 + [2], dtype=torch.int32).cuda()
    mult = torch.tensor([4.0] * 6 + [8.0] * 3 + [16.0]).cuda()
    scales = [torch.tensor([v], requires_grad=True, device="cuda") for v in (1.0, 0.7, 1.3)]
    raws = [torch.randn(B, 8, S, generator=g).cuda().requires_grad_(True), torch.randn(B, 16, S, generator=g).cuda().requires_grad_(True)]
    biases = [torch.randn(2, generator=g).cuda().requires_grad_(True), torch.randn(15, generator=g).cuda().requires_grad_(True)]
    idx = sep_idx.long()
    z0 = raws[0][:, :2, idx].permute(0, 2, 1) + biases[0]
    want0 = torch.exp(z0 * torch.cat(scales)[level.long()].view(1, P, 1)) * mult.view(1, P, 1)
    want1 = raws[1][:, :15, idx].permute(0, 2, 1) + biases[1]
    outs = ops.head_gather_fwd([r.detach() for r in raws], [2, 15], [1, 0], [b.detach() for b in biases], sep_idx, level, mult,
                               [s.detach() for s in scales])
    assert rel(outs[0], want0.detach()) < 1e-5 and rel(outs[1], want1.detach()) < 1e-6
    g0, g1 = torch.randn(B, P, 2, generator=g).cuda(), torch.randn(B, P, 15, generator=g).cuda()
    (want0 * g0).sum().add((want1 * g1).sum()).backward()
    dbias = [torch.zeros(2).cuda(), torch.zeros(15).cuda()]
    dscale = [torch.zeros(1).cuda() for _ in range(3)]
    dps = ops.head_gather_bwd([r.detach() for r in raws], [2, 15], [1, 0], [b.detach() for b in biases], dbias, [g0, g1], outs, sep_idx,
                              prior_of_col.cuda(), level, mult, [s.detach() for s in scales], dscale)
    for dp, r in zip(dps, raws):
        assert rel(planes_f32(dp), r.grad.permute(0, 2, 1)) < 4e-5
    for a, b in zip(dbias, biases):
        assert rel(a, b.grad) < 1e-5
    for a, b in zip(dscale, scales):
        assert rel(a, b.grad) < 1e-5
    # one head without upstream gradient: zero planes
    dps = ops.head_gather_bwd([r.detach() for r in raws], [2, 15], [1, 0], [b.detach() for b in biases], [None, None], [None, g1], outs,
                              sep_idx, prior_of_col.cuda(), level, mult, [s.detach() for s in scales], dscale)
    assert float(planes_f32(dps[0]).abs().max()) == 0


@pytest.mark.parametrize("variant,B", [("thumos", 2), ("anet", 1)])
def test_schedule_matches_the_autograd_formulation(variant, B):
    """The same CoarsePyramid, same weights, same features: explicit schedule vs per-module autograd — outputs, feature
    gradients and every parameter gradient."""
    from opental_b200.bdnet import CoarsePyramid
    torch.manual_seed(3)
    frames = 256 if variant == "thumos" else 768
    cp = CoarsePyramid([832, 1024], 15, frames, os_head=True, variant=variant).cuda()
    for m in cp.modules():
        if isinstance(m, (torch.nn.Conv1d, torch.nn.Conv3d)):
            torch.nn.init.normal_(m.weight, std=(2.0 / m.weight[0].numel()) ** 0.5)
            torch.nn.init.normal_(m.bias, std=0.02)
        if isinstance(m, torch.nn.GroupNorm):
            torch.nn.init.normal_(m.weight, mean=1.0, std=0.1)
            torch.nn.init.normal_(m.bias, std=0.1)
    t5 = frames // 8
    x1 = torch.randn(B, 832, 2 * t5, 6, 6).cuda().relu() if variant == "thumos" else None
    x2 = torch.randn(B, 1024, t5, 3, 3).cuda().relu()
    res = {}
    for sched in (True, False):
        cp.native_schedule = sched
        cp.zero_grad(set_to_none=True)
        if cp.conv_store is not None and cp.conv_store.dev is not None:
            cp.conv_store.flat_g.zero_()
        a1 = x1.clone().requires_grad_(True) if x1 is not None else None
        a2 = x2.clone().requires_grad_(True)
        out = cp({"Mixed_4f": a1, "Mixed_5c": a2})
        torch.manual_seed(11)
        cost = 0
        for k in ("loc", "conf", "act", "prop_loc", "prop_conf", "prop_act", "center", "start", "end", "start_loc_prop", "end_loc_prop",
                  "start_conf_prop", "end_conf_prop"):
            w = torch.randn(out[k].shape, device="cuda")
            cost = cost + (out[k] * w).sum() * (0.01 if k == "loc" else 1.0)
        cost.backward()
        res[sched] = ({k: v.detach().clone() for k, v in out.items() if v is not None}, None if a1 is None else a1.grad.clone(), a2.grad.clone(),
                      {n: p.grad.detach().clone() for n, p in cp.named_parameters()})
    for k, v in res[False][0].items():
        assert tuple(res[True][0][k].shape) == tuple(v.shape), k
        assert rel(res[True][0][k], v) < 1e-4, (k, rel(res[True][0][k], v))
    if x1 is not None:
        assert rel(res[True][1], res[False][1]) < 2e-4
    assert rel(res[True][2], res[False][2]) < 2e-4
    bad = {n: rel(res[True][3][n], gref) for n, gref in res[False][3].items() if rel(res[True][3][n], gref) > 3e-4}
    assert not bad, bad
