"""Developer profile of one training step on the GPU box (not a pytest, not the bench):
  * phase times (backbone fwd / head fwd + loss / backward / optimizer) with CUDA events,
  * every C-ABI launch timed with events and aggregated by (entry point, shape label) -> per-layer TFLOP/s,
  * optional: the same step captured into a CUDA graph and replayed.
Usage:  python tools/step_profile.py [--batch 8] [--precision bf16x3] [--graph] > gpurun_out/step_profile.txt"""
import argparse
import collections
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from opental_b200 import _lib, engine  # noqa: E402
from opental_b200.multisegment_loss import pad_targets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--syncs", action="store_true")
    ap.add_argument("--top", type=int, default=70)
    ap.add_argument("--u8", action="store_true", help="uint8 frames as input (the bench's path: raw-pixel Conv3d_1a)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net, crit = engine.build_opental(device=dev, precision=args.precision, epoch=11)
    tr = engine.Trainer(net, crit)
    B = args.batch
    if args.u8:
        clips = torch.stack([engine.synthetic_clip_u8(i) for i in range(B)]).to(dev)
    else:
        clips = torch.stack([engine.normalise_clip(engine.synthetic_clip_u8(i)) for i in range(B)]).to(dev)
    tg = [engine.synthetic_targets(i) for i in range(B)]
    sc = torch.stack([engine.synthetic_scores(t) for t in tg]).to(dev)
    tp, tv = (t.to(dev) for t in pad_targets(tg, device="cpu"))

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    for _ in range(3):
        tr.step(clips, (tp, tv), sc)
    torch.cuda.synchronize()

    # ---- whole step, eager
    t0 = time.perf_counter(); e0 = ev()
    for _ in range(5):
        tr.step(clips, (tp, tv), sc)
    e1 = ev(); torch.cuda.synchronize()
    print(f"eager step: {e0.elapsed_time(e1) / 5:.2f} ms (events), {(time.perf_counter() - t0) * 200:.2f} ms (wall), batch {B}, {args.precision}")
    # ---- how long does the host need to enqueue one step (GPU may lag behind)?
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tr.step(clips, (tp, tv), sc)
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"host enqueue time of one step: {t_enq * 1e3:.2f} ms")

    # ---- phases
    tr.zero_grad()
    torch.cuda.synchronize()
    a = ev()
    feat = net.backbone(clips)
    b = ev()
    out = net.coarse_pyramid_detection(feat)
    pass
    c = ev()
    losses = crit(out, (tp, tv))
    from opental_b200.multisegment_loss import training_cost
    cost, _, _ = training_cost(out, losses, sc)
    d = ev()
    hooks = {}
    net.backbone.on_backward_start = lambda: hooks.setdefault("bb", ev())
    cost.backward()
    net.backbone.on_backward_start = None
    e = ev()
    for (w, g), st in zip(tr.groups, tr.state):
        from opental_b200 import ops
        ops.adam_step(w, g, st["m"], st["v"], lr=1e-5, weight_decay=1e-3, step=10)
    f = ev()
    torch.cuda.synchronize()
    print("phases (sync'd before, so host enqueue overlaps less than in a real step):")
    print(f"  backbone fwd {a.elapsed_time(b):8.2f} ms\n  head fwd     {b.elapsed_time(c):8.2f} ms\n  loss         {c.elapsed_time(d):8.2f} ms")
    print(f"  head+loss bwd{d.elapsed_time(hooks['bb']):8.2f} ms\n  backbone bwd {hooks['bb'].elapsed_time(e):8.2f} ms\n  adam         {e.elapsed_time(f):8.2f} ms")

    # ---- per-launch trace
    _lib.TRACE = []
    tr.step(clips, (tp, tv), sc)
    torch.cuda.synchronize()
    trace, _lib.TRACE = _lib.TRACE, None
    agg = collections.OrderedDict()
    for name, label, s, e_ in trace:
        key = (name, label[0] if label else "")
        r = agg.setdefault(key, [0, 0.0, 0.0])
        r[0] += 1; r[1] += s.elapsed_time(e_); r[2] += label[1] if label else 0.0
    tot = sum(r[1] for r in agg.values())
    print(f"traced {len(trace)} C-ABI launches, {tot:.2f} ms between their events")
    by_entry = collections.defaultdict(lambda: [0, 0.0])
    for (name, _), r in agg.items():
        by_entry[name][0] += r[0]; by_entry[name][1] += r[1]
    for name, (n, ms) in sorted(by_entry.items(), key=lambda kv: -kv[1][1]):
        print(f"  {ms:8.3f} ms {n:5d}  {name}")
    print("top launches by time (ms, count, TFLOP/s algorithmic):")
    for (name, label), r in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        tf = r[2] / (r[1] * 1e-3) / 1e12 if r[1] > 0 and r[2] > 0 else 0.0
        print(f"  {r[1]:8.3f} {r[0]:4d} {tf:7.1f}  {name[5:]:22s} {label}")

    if args.syncs:
        import traceback
        import warnings
        seen = collections.Counter()

        def show(message, category, filename, lineno, file=None, line=None):
            if "synchroniz" in str(message):
                st = [f for f in traceback.extract_stack() if "opental_b200" in f.filename or "bench" in f.filename]
                seen[" <- ".join(f"{os.path.basename(f.filename)}:{f.lineno}" for f in st[-3:])] += 1

        warnings.showwarning = show
        warnings.simplefilter("always")
        torch.cuda.set_sync_debug_mode("warn")
        tr.step(clips, (tp, tv), sc)
        torch.cuda.set_sync_debug_mode("default")
        torch.cuda.synchronize()
        print("host<->device synchronisations inside one eager step:", sum(seen.values()))
        for k, v in seen.most_common(20):
            print(f"  {v:4d}  {k}")

    if args.graph:
        try:
            del cost, losses, out, feat
            tr.capture(clips, (tp, tv), sc)
            for _ in range(2):
                tr.step(clips, (tp, tv), sc)
            torch.cuda.synchronize()
            e0 = ev()
            for _ in range(5):
                cost, losses, ls, le = tr.step(clips, (tp, tv), sc)
            e1 = ev(); torch.cuda.synchronize()
            print(f"CUDA-graph step (replay + Adam): {e0.elapsed_time(e1) / 5:.2f} ms, cost {float(cost):.4f}, "
                  f"losses {[round(float(v), 4) for v in losses]}")
        except Exception as ex:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            print("graph capture failed:", repr(ex)[:1500])


if __name__ == "__main__":
    main()
