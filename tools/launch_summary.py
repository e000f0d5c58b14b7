#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum CSV). usage: tools/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i + 1
        break
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
t, n = collections.defaultdict(float), collections.Counter()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[-44:]
    t[name] += v
    n[name] += 1
tot = sum(t.values())
print(f"{'kernel':46s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>10s}")
for k, v in sorted(t.items(), key=lambda x: -x[1]):
    print(f"{k:46s} {n[k]:8d} {v / 1e3:10.3f} {100 * v / tot:6.1f}% {v / n[k]:10.1f}")
print(f"{'total':46s} {sum(n.values()):8d} {tot / 1e3:10.3f}")
