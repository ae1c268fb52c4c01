"""dev/launch_breakdown.py <launches.csv> -- per-launch durations of the LAST splat iteration in an ncu launch list."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    seq.append((r[ki][:70], v))
idx = [i for i, (k, v) in enumerate(seq) if "preprocess" in k]
last = seq[idx[-1] - 2:] if idx else seq
tot = 0.0
for k, v in last:
    print(f"{v:9.1f} us  {k}")
    tot += v
print(f"{tot:9.1f} us  total")
