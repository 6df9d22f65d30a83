"""Conv3d_1a at the batch-8 THUMOS14 shape: the bf16x3 form on normalised planes vs the STAGED raw-uint8 form
(otal_clip_ingest_u8_raw / otal_conv1a_fwd_u8 / otal_conv1a_wgrad_u8 + otal_border_class_sums), CUDA-event timed,
with a parity check between the two.  `python tools/conv1a_bench.py [--ncu]` (--ncu: one launch each, no timing loop)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from opental_b200 import ops  # noqa: E402

B = int(os.environ.get("NCU_BATCH", "8"))
dev = "cuda"
torch.manual_seed(0)
px = torch.randint(0, 256, (B, 256, 112, 112, 3), dtype=torch.uint8, device=dev)
w = torch.randn(64, 3, 7, 7, 7, device=dev) * 0.03
wp = ops.pack_conv1a_weight(w)
scale, shift = 1 + 0.1 * torch.randn(64, device=dev), 0.1 * torch.randn(64, device=dev)
d = ops.split_bf16(torch.randn(B, 128, 48, 48, 64, device=dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > L2


def timed(fn, reps):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def x3():
    a = ops.clip_ingest_u8(px, 96)
    y = ops.conv1a_fwd(a, wp, 96, scale=scale, shift=shift)
    dw = torch.zeros(49, 64, 32, device=dev)
    ops.conv1a_wgrad(a, d, dw, 96)
    return y, ops.unpack_conv1a_wgrad(dw)


def u8():
    a = ops.clip_ingest_u8(px, 96, raw=True)
    sc, tab = ops.conv1a_u8_scale_shift(w, scale, shift)
    y = ops.conv1a_fwd(a, wp, 96, scale=sc, shift=tab, u8=True)
    dw = torch.zeros(49, 64, 32, device=dev)
    ops.conv1a_wgrad(a, d, dw, 96, u8=True)
    return y, ops.conv1a_u8_weight_grad(dw)


ya, ga = x3()
yb, gb = u8()
torch.cuda.synchronize()
rel = lambda p, q: float((p - q).abs().max() / q.abs().max())
print(f"parity u8 vs bf16x3: forward {rel(yb.float(), ya.float()):.2e}, weight gradient {rel(gb, ga):.2e}")
if "--ncu" in sys.argv:
    sys.exit(0)
reps = 11
a3 = ops.clip_ingest_u8(px, 96)
a8 = ops.clip_ingest_u8(px, 96, raw=True)
sc8, tab8 = ops.conv1a_u8_scale_shift(w, scale, shift)
wcat = ops.pack_conv1a_weight_cat(wp)
dw = torch.zeros(49, 64, 32, device=dev)
rows = [
    ("ingest  normalised hi+lo", lambda: ops.clip_ingest_u8(px, 96)),
    ("ingest  raw one plane   ", lambda: ops.clip_ingest_u8(px, 96, raw=True)),
    ("fwd     bf16x3          ", lambda: ops.conv1a_fwd(a3, wp, 96, scale=scale, shift=shift)),
    ("fwd     u8              ", lambda: ops.conv1a_fwd(a8, wp, 96, scale=sc8, shift=tab8, u8=True)),
    ("fwd     u8 halo         ", lambda: ops.conv1a_fwd(a8, wp, 96, scale=sc8, shift=tab8, u8=True, w_cat=wcat)),
    ("wgrad   bf16x3          ", lambda: ops.conv1a_wgrad(a3, d, dw, 96)),
    ("wgrad   u8 generic      ", lambda: (setattr(ops, "CONV1A_WGRAD_HALO", False), ops.conv1a_wgrad(a8, d, dw, 96, u8=True))),
    ("wgrad   u8 halo         ", lambda: (setattr(ops, "CONV1A_WGRAD_HALO", True), ops.conv1a_wgrad(a8, d, dw, 96, u8=True))),
    ("class sums of dY        ", lambda: ops.border_class_sums(d)),
    ("host algebra (shift tab)", lambda: ops.conv1a_u8_scale_shift(w, scale, shift)),
]
for name, fn in rows:
    fn(); torch.cuda.synchronize()
    print(f"{name}  {timed(fn, reps):8.3f} ms")
