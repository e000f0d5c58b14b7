#!/usr/bin/env python
"""SASS evidence from the shipped library (no GPU needed): per-kernel counts of the instructions that prove the bulk async copies,
their mbarrier, the shared-memory majorant loads, and the hardware special-function units of the FAST medium shading.
usage: python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
lib = os.path.join(ROOT, "narvalengine_b200", "lib", "libnarval_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
sha = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
per, cur = collections.defaultdict(collections.Counter), None
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if cur and m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for key in ("UBLKCP", "SYNCS", "LDS.U16", "MUFU.RSQ", "MUFU.RCP", "MUFU.SIN", "MUFU.COS", "MUFU.SQRT", "REDUX", "ATOMG", "RED.E"):
            if op.startswith(key):
                per[cur][key] += 1
def short(n):
    d = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()
    return d.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("(WfBuf")[0].replace("void ", "")[:60]
print(f"# git {sha}, csrc_sha16 {bench.csrc_sha16()}; cuobjdump -sass narvalengine_b200/lib/libnarval_b200.so (sm_100a): static instruction counts per kernel")
keys = ["_total", "UBLKCP", "SYNCS", "LDS.U16", "MUFU.RSQ", "MUFU.RCP", "MUFU.SQRT", "MUFU.SIN", "MUFU.COS", "REDUX"]
print(f"{'kernel':62s}" + "".join(f"{k.replace('_total','instr'):>10s}" for k in keys))
for fn in sorted(per, key=short):
    s = short(fn)
    if not s.startswith("k_wf_"):
        continue
    print(f"{s:62s}" + "".join(f"{per[fn][k]:10d}" for k in keys))
print("""
UBLKCP = cp.async.bulk global -> shared (the majorant tables of the TRACK_*_SM kernels), SYNCS = their mbarrier, LDS.U16 = the half-float
majorant read from shared memory at every brick crossing. k_wf_scatter<FUSE, LS, FASTSH>: the FASTSH = true variants (production) use the
special-function unit (MUFU.RSQ / RCP / SQRT / SIN / COS) where the FASTSH = false variants (NE_B200_EXACT_SHADING=1, the reference-order
arithmetic) run IEEE division / square-root sequences: compare the instruction totals of the two.""")
