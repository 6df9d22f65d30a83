"""Developer tool: kernel timeline of one captured training step (torch.profiler / CUPTI): per stream busy time, idle gaps on the
main stream, and the longest kernels — where a graph replay's wall time goes when the per-kernel sums say it should be shorter.
Usage: python tools/graph_timeline.py > gpurun_out/graph_timeline.txt"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from opental_b200 import engine  # noqa: E402
from opental_b200.multisegment_loss import pad_targets  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
net, crit = engine.build_opental(device=dev, epoch=11)
tr = engine.Trainer(net, crit)
B = 8
clips = torch.stack([engine.synthetic_clip_u8(i) for i in range(B)]).to(dev)
tg = [engine.synthetic_targets(i) for i in range(B)]
sc = torch.stack([engine.synthetic_scores(t) for t in tg]).to(dev)
tp, tv = (t.to(dev) for t in pad_targets(tg, device="cpu"))
tr.capture(clips, (tp, tv), sc)
for _ in range(3):
    tr.step(clips, (tp, tv), sc)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(clips, (tp, tv), sc)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name and "Memcpy" not in e.name and "Memset" not in e.name]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
t1 = max(e.time_range.end for e in evs)
print(f"kernels {len(evs)}, span {(t1 - t0) / 1e3:.3f} ms")
by_stream = collections.defaultdict(list)
for e in evs:
    by_stream[getattr(e, "stream", None) if hasattr(e, "stream") else e.device_index].append(e)
for s, lst in by_stream.items():
    busy = sum(e.time_range.end - e.time_range.start for e in lst)
    print(f"stream {s}: {len(lst)} kernels, busy {busy / 1e3:.3f} ms, first at {(lst[0].time_range.start - t0) / 1e3:.3f}, last ends {(max(e.time_range.end for e in lst) - t0) / 1e3:.3f}")
# union of busy intervals over all streams -> total idle time of the GPU inside the span
ivs = sorted((e.time_range.start, e.time_range.end) for e in evs)
cur_s, cur_e = ivs[0]
idle, gaps = 0.0, []
for s, e in ivs[1:]:
    if s > cur_e:
        idle += s - cur_e
        gaps.append((s - cur_e, cur_e - t0))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
print(f"GPU idle inside the span (no kernel on any stream): {idle / 1e3:.3f} ms in {len(gaps)} gaps; largest:")
for g, at in sorted(gaps, reverse=True)[:8]:
    print(f"   {g:8.1f} us at {at / 1e3:7.3f} ms")
# time sliced into 0.5 ms windows: which kernels dominate each window
print("timeline (0.5 ms windows: busy fraction of the union, top kernel by time):")
w = 500.0
nwin = int((t1 - t0) / w) + 1
for i in range(nwin):
    lo, hi = t0 + i * w, t0 + (i + 1) * w
    agg = collections.defaultdict(float)
    for e in evs:
        o = min(e.time_range.end, hi) - max(e.time_range.start, lo)
        if o > 0:
            agg[e.name[:60]] += o
    if agg:
        k, v = max(agg.items(), key=lambda kv: kv[1])
        print(f"  {i * 0.5:5.1f} ms  sum {sum(agg.values()) / w:4.2f}  {k} ({v / w:4.2f})")
