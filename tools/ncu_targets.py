"""One launch of every hot kernel at its batch-8 THUMOS14 shape, for `ncu --set full` captures (profiles/):

    ncu --set full --clock-control none --import-source on \
        -k regex:'conv_igemm|conv_wgrad|maxpool|bmp_|msl_|gn_relu|relu_bn|clip_ingest|decode_scores|softnms|boundary_bce' \
        -c 48 -o gpurun_out/r01_kernels python tools/ncu_targets.py

Prints the launch order so the report's IDs can be mapped back to layers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from opental_b200 import ops  # noqa: E402
from opental_b200.backbone import _pads  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
B = int(os.environ.get("NCU_BATCH", "8"))
order = []


def planes(shape, relu=True):
    x = torch.randn(*shape, device=dev)
    if relu:
        x = x.relu()
    return ops.split_bf16(x)


def wplanes(cout, cin, k):
    w = torch.randn(cout, cin, *k, device=dev) * (2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5
    return ops.pack_conv_weight(w)


def conv_trio(name, shape, cin, cout, k):
    x = planes((B, *shape, cin))
    w = wplanes(cout, cin, k)
    sc = torch.ones(cout, device=dev); sh = torch.zeros(cout, device=dev)
    pads = _pads(shape, k)
    y = ops.conv_igemm(x, w, kernel=k, pad_front=pads, scale=sc, shift=sh, relu=True)
    order.append(f"conv_igemm_kernel  fwd   {name}")
    d = planes((B, *shape, cout), relu=False)
    gx = torch.empty(B, *shape, cin, device=dev)
    ops.conv_igemm(d, w, kernel=k, pad_front=tuple(kk - 1 - p for kk, p in zip(k, pads)), out_f32=gx, want_planes=False, dgrad=True)
    order.append(f"conv_igemm_kernel  dgrad {name}")
    dw = torch.zeros(k[0] * k[1] * k[2], cout, cin, device=dev)
    ops.conv_wgrad(x, d, dw, kernel=k, pad_front=pads)
    order.append(f"conv_wgrad_kernel  wgrad {name}")
    return y


# Conv3d_1a_7x7 forward + weight gradient
px = torch.randint(0, 256, (B, 256, 112, 112, 3), dtype=torch.uint8, device=dev)
a = ops.clip_ingest_u8(px, 96)
order.append("clip_ingest_u8_kernel  uint8 [B,256,112,112,3] -> W-padded planes [B,256,96,104,4]")
w1 = ops.pack_conv1a_weight(torch.randn(64, 3, 7, 7, 7, device=dev) * 0.03)
y1 = ops.conv1a_fwd(a, w1, 96, scale=torch.ones(64, device=dev), shift=torch.zeros(64, device=dev))
order.append("conv_igemm_kernel  fwd   Conv3d_1a_7x7 (folded)")
d1 = planes((B, 128, 48, 48, 64), relu=False)
dw1 = torch.zeros(49, 64, 32, device=dev)
ops.conv1a_wgrad(a, d1, dw1, 96)
order.append("conv_wgrad_kernel  wgrad Conv3d_1a_7x7 (folded)")
del px, a, d1

conv_trio("Conv3d_2c_3x3 64->192 @128x24x24", (128, 24, 24), 64, 192, (3, 3, 3))
conv_trio("Mixed_3c.b1b 128->192 @128x12x12", (128, 12, 12), 128, 192, (3, 3, 3))
conv_trio("Mixed_3c.b0 256->128 1x1 @128x12x12", (128, 12, 12), 256, 128, (1, 1, 1))
conv_trio("Mixed_4f.b1b 160->320 @64x6x6", (64, 6, 6), 160, 320, (3, 3, 3))

# max pools
x = planes((B, 128, 48, 48, 64))
yp, arg2a = ops.maxpool_fwd(x, kernel=(1, 3, 3), stride=(1, 2, 2), pad_front=_pads((128, 48, 48), (1, 3, 3), (1, 2, 2)), save_argmax=True)
order.append("maxpool_fwd_fast<1,3,3,1,2,2>  MaxPool3d_2a @128x48x48x64 (+arg-max)")
g = torch.randn(B, 128, 24, 24, 64, device=dev)
ops.maxpool_bwd_relu_bn_split(x, arg2a, g, torch.ones(64, device=dev), kernel=(1, 3, 3), stride=(1, 2, 2),
                              pad_front=_pads((128, 48, 48), (1, 3, 3), (1, 2, 2)))
order.append("maxpool_bwd_gather_fused  MaxPool3d_2a backward + ReLU/BN backward + split")
del arg2a
x = planes((B, 128, 12, 12, 256))
_, arg3 = ops.maxpool_fwd(x, kernel=(3, 3, 3), stride=(1, 1, 1), pad_front=(1, 1, 1), save_argmax=True)
order.append("maxpool333_tiled  Mixed_3c.b3a (3,3,3)/1 @128x12x12x256 (+arg-max)")
g = torch.randn(B, 128, 12, 12, 256, device=dev)
gi = torch.zeros_like(g)
ops.maxpool_bwd(x, g, gi, kernel=(3, 3, 3), stride=(1, 1, 1), pad_front=(1, 1, 1), argmax=arg3)
order.append("maxpool_bwd_argmax  Mixed_3c.b3a backward (scatter)")
yq = planes((B, 128, 12, 12, 256))
ops.relu_bn_bwd_split(g, yq, torch.ones(256, device=dev))
order.append("relu_bn_bwd_split  @128x12x12x256")

# BoundaryMaxPooling at the level-batched call shapes
inp = torch.randn(B, 1024, 126, device=dev)
c = torch.rand(B, 126, 1, device=dev) * 126
seg = torch.cat([c - 8, c + 3, c - 3, c + 8], -1).round().contiguous()
o = ops.bmp_forward(inp, seg)
order.append("bmp_forward_kernel  [B,1024,126] x [B,126,4]")
ops.bmp_backward(torch.randn_like(o), inp, seg, False)
order.append("bmp_backward_kernel [B,1024,126] x [B,126,4]")
frame = torch.randn(B, 512, 256, device=dev)
c = torch.rand(B, 126, 1, device=dev) * 256
fs = torch.cat([c - 30, c + 8, c - 8, c + 30], -1).round().contiguous()
o = ops.bmp_forward(frame, fs)
order.append("bmp_forward_kernel  [B,512,256] x [B,126,4] (frame level)")
ops.bmp_backward(torch.randn_like(o), frame, fs, False)
order.append("bmp_backward_kernel [B,512,256] x [B,126,4] (frame level)")

# GroupNorm + ReLU (segmented, sep layout) and the fused loss
segs = ((1, 64), (66, 32), (99, 16), (116, 8), (125, 4), (130, 2))
xg = torch.randn(B, 512, 136, device=dev, requires_grad=True)
wg = torch.ones(512, device=dev, requires_grad=True); bg = torch.zeros(512, device=dev, requires_grad=True)
yg = ops.groupnorm_relu(xg, wg, bg, 32, 1e-5, True, segs)
order.append("gn_relu_fwd_kernel  [B,512,136] 6 segments")
yg.backward(torch.randn_like(yg))
order.append("gn_relu_bwd_kernel  [B,512,136] 6 segments")

from opental_b200.engine import OPENTAL_ACT_CONFIG, OPENTAL_EDL_CONFIG, synthetic_targets  # noqa: E402
from opental_b200.multisegment_loss import MultiSegmentLoss  # noqa: E402
crit = MultiSegmentLoss(15, 0.5, 1.0, cls_loss_type="edl", edl_config=OPENTAL_EDL_CONFIG, os_head=True, act_config=OPENTAL_ACT_CONFIG).cuda()
crit.cls_loss.epoch = 11
P = 126
pri = torch.cat([(torch.arange(t) + 0.5) / t for t in (64, 32, 16, 8, 4, 2)]).view(-1, 1).cuda()
out = dict(loc=(torch.rand(B, P, 2, device=dev) * 30 + 1).requires_grad_(True), conf=torch.randn(B, P, 15, device=dev).requires_grad_(True),
           prop_loc=(0.3 * torch.randn(B, P, 2, device=dev)).requires_grad_(True), prop_conf=torch.randn(B, P, 15, device=dev).requires_grad_(True),
           center=torch.randn(B, P, 1, device=dev).requires_grad_(True), act=torch.randn(B, P, 1, device=dev).requires_grad_(True),
           prop_act=torch.randn(B, P, 1, device=dev).requires_grad_(True), priors=pri)
losses = crit(out, [synthetic_targets(i).cuda() for i in range(B)])
order.append("msl_forward_kernel  B*P = %d" % (B * P))
sum(losses).backward()
order.append("msl_backward_kernel")
# boundary BCE and inference post-processing
from opental_b200.multisegment_loss import calc_bce_loss  # noqa: E402
st = torch.randn(B, 256, 256, device=dev).relu().requires_grad_(True)
en = torch.randn(B, 256, 256, device=dev).relu().requires_grad_(True)
scm = (torch.rand(B, 2, 256, device=dev) > 0.7).float()
ls, le = calc_bce_loss(st, en, scm)
order.append("boundary_bce_fwd_kernel x2  [B,256,256]")
(ls + le).backward()
order.append("boundary_bce_bwd_kernel x2")
seg, sco, un, ac = ops.decode_scores({k: v.detach() for k, v in out.items()}, torch.zeros(B), 256, 10.0)
order.append("decode_scores_kernel  B*P = %d, K = 15" % (B * P))
ops.softnms(seg.reshape(-1, 2), sco.permute(1, 0, 2).reshape(15, -1), sigma=0.5, top_k=200)
order.append("softnms_kernel  15 classes x %d candidates, top_k 200" % (B * P))
torch.cuda.synchronize()
print("launch order of the profiled kernels:")
for i, o_ in enumerate(order):
    print(f"  {i:2d}  {o_}")
