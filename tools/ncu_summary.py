"""Condense an ncu report (`ncu -i X.ncu-rep --page raw --csv`) into one line per launch with the metrics the
roofline argument uses.  Usage: ncu -i rep --page raw --csv | python tools/ncu_summary.py [labels.txt]"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
labels = []
if len(sys.argv) > 1:
    labels = [l.strip() for l in open(sys.argv[1]) if l.strip()]
print("# id kernel | " + " | ".join(f"{n} [{units[idx[m]]}]" for m, n in WANT if m in idx) + " | what")
for i, r in enumerate(rows[2:]):
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    vals = [r[idx[m]] for m, _ in WANT if m in idx]
    print(f"{i:2d} {name:28s} | " + " | ".join(f"{v:>10s}" for v in vals) + " | " + (labels[i] if i < len(labels) else ""))
