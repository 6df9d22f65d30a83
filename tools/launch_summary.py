"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches and device time for one
training step (the span between two clip-ingest launches).  Usage: python tools/launch_summary.py launches.csv [step]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit == "ms" else v)
        out.append((row["Kernel Name"], v, row["Grid Size"], row["Block Size"]))
    return out


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"at::native::", "", n)
    return n[:100]


def main():
    rows = load(sys.argv[1])
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    marks = [i for i, r in enumerate(rows) if "clip_ingest" in r[0]]
    print("launches captured:", len(rows), "clip_ingest at", marks)
    if marks:
        start = marks[which]
        end = marks[which + 1] if which + 1 < len(marks) else len(rows)
    else:
        start, end = 0, len(rows)
    seg = rows[start:end]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v, g, b in seg:
        a = agg[short(n)]
        a[0] += 1
        a[1] += v
    tot = sum(r[1] for r in seg)
    print(f"step {which}: {len(seg)} launches, {tot / 1e3:.2f} ms of device time (cold-cache, serialised)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{t / 1e3:9.3f} ms {100 * t / tot:5.1f}% {c:5d}  {k}")


if __name__ == "__main__":
    main()
