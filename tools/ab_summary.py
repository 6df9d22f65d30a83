"""Summarise the A/B benches of tools/r02_first_gpu_call.sh (gpurun_out/r02_ab_<switches>_pass<k>.json): ms/step per configuration
and pass, the mean, and the difference to the baseline of the same pass (interleaved passes cancel box-to-box drift).

    python tools/ab_summary.py [gpurun_out] > profiles/r02_ab_summary.txt"""
import glob
import json
import os
import re
import sys

root = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
runs: dict[str, dict[int, dict]] = {}
for path in sorted(glob.glob(os.path.join(root, "r02_ab_*_pass*.json"))):
    m = re.match(r"r02_ab_(.+)_pass(\d+)\.json", os.path.basename(path))
    try:
        line = [ln for ln in open(path).read().splitlines() if ln.startswith("{")][-1]
        runs.setdefault(m.group(1), {})[int(m.group(2))] = json.loads(line)
    except (IndexError, ValueError):
        runs.setdefault(m.group(1), {})[int(m.group(2))] = None
if not runs:
    sys.exit(f"no r02_ab_*.json under {root}")
passes = sorted({p for v in runs.values() for p in v})
base = runs.get("base", {})
print(f"{'configuration':48s} " + " ".join(f"pass{p:>2d} ms" for p in passes) + "     mean   vs base   clips/s   conv_igemm ms  wgrad ms")
for tag in ["base"] + sorted(t for t in runs if t != "base"):
    if tag not in runs:
        continue
    ms = [runs[tag].get(p, None) for p in passes]
    vals = [r["ms_per_step"] if r else float("nan") for r in ms]
    good = [v for v in vals if v == v]
    mean = sum(good) / len(good) if good else float("nan")
    deltas = [r["ms_per_step"] - base[p]["ms_per_step"] for p, r in zip(passes, ms) if r and base.get(p)]
    d = sum(deltas) / len(deltas) if deltas else float("nan")
    last = next((r for r in reversed(ms) if r), None)
    ra = (last or {}).get("roofline_all", {})
    ci = ra.get("conv_igemm_kernel", {}).get("ms_per_step", float("nan"))
    cw = ra.get("conv_wgrad_kernel", {}).get("ms_per_step", float("nan"))
    print(f"{tag:48s} " + " ".join(f"{v:9.3f}" for v in vals) + f" {mean:8.3f} {d:+9.3f} {(last or {}).get('value', float('nan')):9.1f} {ci:13.3f} {cw:9.3f}")
