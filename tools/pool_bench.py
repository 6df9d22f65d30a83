"""Developer micro-benchmark of the max-pool kernels at the batch-8 I3D shapes (CUDA events, L2 flushed between reps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from opental_b200 import ops
from opental_b200.backbone import _pads
B = 8
SHAPES = [((128, 48, 48), 64, (1, 3, 3), (1, 2, 2)), ((128, 24, 24), 192, (1, 3, 3), (1, 2, 2)), ((128, 12, 12), 192, (3, 3, 3), (1, 1, 1)),
          ((128, 12, 12), 256, (3, 3, 3), (1, 1, 1)), ((128, 12, 12), 480, (3, 3, 3), (2, 2, 2)), ((64, 6, 6), 512, (3, 3, 3), (1, 1, 1)),
          ((64, 6, 6), 832, (2, 2, 2), (2, 2, 2)), ((32, 3, 3), 832, (3, 3, 3), (1, 1, 1))]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=5):
    fn(); ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
tot = [0, 0, 0]
for shape, C, k, s in SHAPES:
    x = ops.split_bf16(torch.randn(B, *shape, C, device="cuda").relu())
    pads = _pads(shape, k, s)
    y, arg = ops.maxpool_fwd(x, kernel=k, stride=s, pad_front=pads, save_argmax=True)
    g = torch.randn(*y.hi.shape, device="cuda"); gi = torch.zeros(*x.hi.shape, device="cuda")
    t0 = timeit(lambda: ops.maxpool_fwd(x, kernel=k, stride=s, pad_front=pads, out=y))
    t1 = timeit(lambda: ops.maxpool_fwd(x, kernel=k, stride=s, pad_front=pads, out=y, save_argmax=True))
    t2 = timeit(lambda: ops.maxpool_bwd(x, g, gi, kernel=k, stride=s, pad_front=pads, argmax=arg))
    t3 = timeit(lambda: ops.maxpool_bwd(x, g, gi, kernel=k, stride=s, pad_front=pads))
    byt = (x.hi.numel() + y.hi.numel()) * 4
    extra = ""
    if s != (1, 1, 1):      # stage pools: backward fused with the ReLU / BN backward of the producing layer (gather form)
        sc = torch.rand(C, device="cuda") + 0.5
        t4 = timeit(lambda: ops.maxpool_bwd_relu_bn_split(x, arg, g, sc, kernel=k, stride=s, pad_front=pads))
        fb = x.hi.numel() * (4 + 2) + y.hi.numel() * (4 + 1)         # d planes written, y_hi read; pooled gradient + arg-max read
        extra = f"  bwd fused+relu_bn {t4:.3f} ({fb / t4 / 1e6:.0f} GB/s)"
    print(f"{shape} C{C} k{k} s{s}: fwd {t0:.3f} ms ({byt / t0 / 1e6:.0f} GB/s)  fwd+argmax {t1:.3f}  bwd(argmax) {t2:.3f}  bwd(recompute) {t3:.3f}{extra}")
    tot[0] += t0; tot[1] += t1; tot[2] += t2
print("sum (one of each):", [round(t, 3) for t in tot], "WB", os.environ.get("OTAL_POOL_WB"))
