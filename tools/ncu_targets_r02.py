"""Round-2 `ncu --set full` targets: one launch of every kernel that is new or changed this round, at its batch-8 THUMOS14 shape.

    ncu --set full --clock-control none --import-source on -k regex:'conv1a|conv_igemm|conv_wgrad|maxpool|gn_relu|msl_|rows_combine|head_gather' \
        -c 40 -o gpurun_out/r02_kernels python tools/ncu_targets_r02.py

Prints the launch order so the report's IDs can be mapped back to layers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

from opental_b200 import ops  # noqa: E402
from opental_b200.backbone import _pads  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
B = 8
order = []


def planes(shape, relu=True):
    x = torch.randn(*shape, device=dev)
    return ops.split_bf16(x.relu() if relu else x)


# ---- Conv3d_1a on raw uint8 pixels: resident-halo forward and weight gradient (the round-2 kernels)
px = torch.randint(0, 256, (B, 256, 112, 112, 3), dtype=torch.uint8, device=dev)
a8 = ops.clip_ingest_u8(px, 96, raw=True)
w = torch.randn(64, 3, 7, 7, 7, device=dev) * 0.03
wp = ops.pack_conv1a_weight(w)
sc, tab = ops.conv1a_u8_scale_shift(w, torch.ones(64, device=dev), torch.zeros(64, device=dev))
ops.conv1a_fwd(a8, wp, 96, scale=sc, shift=tab, u8=True, w_cat=ops.pack_conv1a_weight_cat(wp))
order.append("conv1a_halo_kernel        Conv3d_1a forward, raw uint8, resident halo (algorithmic 38.8 GF/clip x 8)")
d1 = planes((B, 128, 48, 48, 64), relu=False)
dw1 = torch.zeros(49, 64, 32, device=dev)
ops.conv1a_wgrad(a8, d1, dw1, 96, u8=True)
order.append("conv1a_wgrad_halo_kernel  Conv3d_1a weight gradient, raw uint8, resident halo")
del px, a8, d1


def conv_pair(name, shape, cin, cout, k, dgrad=True):
    x = planes((B, *shape, cin))
    wt = torch.randn(cout, cin, *k, device=dev) * (2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5
    wpk = ops.pack_conv_weight(wt)
    pads = _pads(shape, k)
    ops.conv_igemm(x, wpk, kernel=k, pad_front=pads, scale=torch.ones(cout, device=dev), shift=torch.zeros(cout, device=dev), relu=True)
    order.append(f"conv_igemm_kernel  fwd   {name}")
    if dgrad:
        d = planes((B, *shape, cout), relu=False)
        gx = torch.empty(B, *shape, cin, device=dev)
        ops.conv_igemm(d, wpk, kernel=k, pad_front=tuple(kk - 1 - p for kk, p in zip(k, pads)), out_f32=gx, want_planes=False, dgrad=True)
        order.append(f"conv_igemm_kernel  dgrad {name}")


conv_pair("Conv3d_2c_3x3 64->192 @128x24x24 (largest launch of the step)", (128, 24, 24), 64, 192, (3, 3, 3))
conv_pair("Mixed_3c.b0 256->128 1x1 @128x12x12", (128, 12, 12), 256, 128, (1, 1, 1))
conv_pair("Conv3d_2b_1x1 64->64 @128x24x24", (128, 24, 24), 64, 64, (1, 1, 1))

# ---- stage pool backward fused with ReLU / BN backward (templated gather)
x = planes((B, 128, 48, 48, 64))
pads = _pads((128, 48, 48), (1, 3, 3), (1, 2, 2))
yp, arg = ops.maxpool_fwd(x, kernel=(1, 3, 3), stride=(1, 2, 2), pad_front=pads, save_argmax=True)
order.append("maxpool_fwd_fast<1,3,3,1,2,2>  MaxPool3d_2a @128x48x48x64 (+arg-max)")
g = torch.randn(B, 128, 24, 24, 64, device=dev)
ops.maxpool_bwd_relu_bn_split(x, arg, g, torch.ones(64, device=dev), kernel=(1, 3, 3), stride=(1, 2, 2), pad_front=pads)
order.append("maxpool_bwd_gather_fused_t<1,3,3,1,2,2>  MaxPool3d_2a backward + ReLU/BN backward + split")
del x, yp, arg, g

# ---- explicit head schedule: extended GroupNorm, glue
segs = ((1, 64), (66, 32), (99, 16), (116, 8), (125, 4), (130, 2))
xg = torch.randn(B, 1024, 136, device=dev)
ga, be = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
y, pl, yt, stats = ops.groupnorm_relu_fwd_ex(xg, ga, be, segments=segs, want_y=True, want_planes=True, yt_range=(1, 64))
order.append("gn_relu_fwd_ex_kernel  [B,1024,136] 6 segments: fp32 + planes + channels-last slice")
ops.groupnorm_relu_bwd_ex(torch.randn_like(xg), xg, ga, be, stats, dgamma=torch.zeros(1024, device=dev), dbeta=torch.zeros(1024, device=dev),
                          dbias=torch.zeros(1024, device=dev), segments=segs)
order.append("gn_relu_bwd_ex_kernel  [B,1024,136] 6 segments: planes + in-place parameter gradients")

# ---- fused loss, ActivityNet flavour (B x 189 priors, 150 classes)
import opental_oracle as O  # noqa: E402  (targets / priors helpers only)
from opental_b200.engine import OPENTAL_EDL_CONFIG  # noqa: E402
from opental_b200.multisegment_loss import MultiSegmentLossANet  # noqa: E402
cfg = O.anet_config()
pri = torch.cat(O.level_priors(cfg), 0).cuda()
P = pri.shape[0]
crit = MultiSegmentLossANet(150, 0.5, 1.0, cls_loss_type="edl", edl_config=OPENTAL_EDL_CONFIG, os_head=True).cuda()
crit.cls_loss.epoch = 11
pred = [(torch.rand(B, P, 2, device=dev) * 60 + 1).requires_grad_(True), torch.randn(B, P, 150, device=dev).requires_grad_(True),
        (0.5 * torch.randn(B, P, 2, device=dev)).requires_grad_(True), torch.randn(B, P, 150, device=dev).requires_grad_(True),
        torch.randn(B, P, 1, device=dev).requires_grad_(True), pri, torch.randn(B, P, 1, device=dev).requires_grad_(True),
        torch.randn(B, P, 1, device=dev).requires_grad_(True)]
losses = crit(pred, [O.synthetic_targets(i, num_classes=150).cuda() for i in range(B)])
order.append("msl_forward_kernel  ActivityNet flavour, B*P = %d, K = 150" % (B * P))
sum(losses).backward()
order.append("msl_backward_kernel")
torch.cuda.synchronize()
print("launch order of the profiled kernels:")
for i, o_ in enumerate(order):
    print(f"  {i:2d}  {o_}")
