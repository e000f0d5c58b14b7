#!/usr/bin/env python
"""Per-kernel DRAM bytes and time of one frame from tools/gpu_bytes.sh's CSV, and the tracking kernels' per-launch average
for bench.py's roofline.traffic. usage: tools/bytes_summary.py gpurun_out/bytes.csv [profiles/r01_traffic.json]"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i + 1
        break
ki, mi, vi, ui, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1}
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])  # launches, seconds, read, write
seen = set()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[-34:]
    v = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1)
    a = agg[name]
    if (r[ii], name) not in seen:
        seen.add((r[ii], name))
        a[0] += 1
    if r[mi] == "gpu__time_duration.sum":
        a[1] += v
    elif r[mi] == "dram__bytes_read.sum":
        a[2] += v
    elif r[mi] == "dram__bytes_write.sum":
        a[3] += v
print(f"{'kernel':36s} {'n':>4s} {'ms':>8s} {'rd GB':>8s} {'wr GB':>8s} {'GB/s':>8s}")
tot = [0.0, 0.0, 0.0]
for k, (n, s, rd, wr) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:36s} {n:4d} {s * 1e3:8.3f} {rd / 1e9:8.3f} {wr / 1e9:8.3f} {(rd + wr) / s / 1e9 if s else 0:8.0f}")
    tot[0] += s; tot[1] += rd; tot[2] += wr
print(f"{'total':36s} {'':4s} {tot[0] * 1e3:8.3f} {tot[1] / 1e9:8.3f} {tot[2] / 1e9:8.3f} {(tot[1] + tot[2]) / tot[0] / 1e9:8.0f}")
if len(sys.argv) > 2:
    tr = {k: v for k, v in agg.items() if k.startswith("k_wf_track") or k.startswith("k_wf_tr<")}
    n = sum(v[0] for v in tr.values())
    b = sum(v[2] + v[3] for v in tr.values())
    json.dump({"kernel": "k_wf_track + k_wf_tr, all launches of one headline frame (ncu, serialised, cold caches)",
               "dram_bytes_per_launch": b / max(1, n), "launches": n,
               "per_kernel": {k: {"launches": v[0], "seconds": v[1], "dram_bytes_read": v[2], "dram_bytes_write": v[3]} for k, v in tr.items()},
               "note": "average over the frame's launches, like bench.py's algorithmic_bytes_per_launch. The excess over the algorithmic bytes is the "
                       "128-byte path records the walks are fetched from and written back to; the brick pool and the majorant table are L2/L1-resident."},
              open(sys.argv[2], "w"), indent=1)
