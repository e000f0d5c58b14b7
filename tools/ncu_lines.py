#!/usr/bin/env python
"""Warp-instructions executed and stall samples per SOURCE LINE of a captured kernel (needs -lineinfo + --import-source on).
usage: tools/ncu_lines.py rep.ncu-rep [N]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


fname, hdr = "", None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci = {c: k for k, c in enumerate(hdr)}
        continue
    if hdr and r[0].isdigit():
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [r[1].strip()[:110], 0, 0, 0])
        a[1] += num(r[ci["Instructions Executed"]])
        a[2] += num(r[ci["Thread Instructions Executed"]])
        a[3] += num(r[ci["# Samples"]])
ti = sum(a[1] for a in agg.values()); ts = sum(a[3] for a in agg.values())
print(f"total warp-instr {ti:.3e}, stall samples {ts}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][3])[:top]:
    print(f"{100*a[3]/max(ts,1):5.1f}% stall {100*a[1]/max(ti,1):5.1f}% instr  lanes {a[2]/max(a[1],1):4.1f}  {f}:{ln}  {a[0]}")
