#!/usr/bin/env python
"""Turns the round's gpurun_out/ captures (tools/gpu_final.sh) into the tracked evidence under profiles/, every file stamped
with the git SHA and the hash of the kernel sources (bench.py refuses a traffic / issue figure measured on other sources).
usage: python tools/make_profiles.py [r02]"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

R = sys.argv[1] if len(sys.argv) > 1 else "r02"
OUT, PRO = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
SHA = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
CSRC = bench.csrc_sha16()
STAMP = f"# git {SHA}, csrc_sha16 {CSRC} (kernel sources of the capture); workload: bench.py C2 frame (FastNoise 256^3, 1080p, 64 spp), one B200, --clock-control none\n"


def run(*cmd):
    return subprocess.run(list(cmd), capture_output=True, text=True, cwd=ROOT).stdout


def raw_metrics(rep):
    rows = list(csv.reader(io.StringIO(run("ncu", "-i", rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    m = {k: rows[2][i] for i, k in enumerate(hdr)}
    m["__units__"] = {k: units[i] for i, k in enumerate(hdr)}
    return m


def main():
    issue = {}
    for k in ("k_wf_track", "k_wf_tr", "k_wf_scatter", "k_wf_generate", "k_wf_extend"):
        rep = os.path.join(OUT, f"{R}_{k}.ncu-rep")
        if not os.path.exists(rep):
            continue
        which = "launch #1 (the only one with camera rays)" if k == "k_wf_generate" else "launch #2 (the first incoherent iteration)"
        text = STAMP + f"# ncu --set full --import-source on, {which} of the kernel in the frame (host-driven loop, NE_B200_HOST_LOOP=1)\n"
        text += run(sys.executable, "tools/ncu_summary.py", rep) + "\n== hottest source lines (stall samples, share of warp instructions, lanes per instruction)\n"
        text += run(sys.executable, "tools/ncu_lines.py", rep, "40")
        open(os.path.join(PRO, f"{R}_{k}.txt"), "w").write(text)
        m = raw_metrics(rep)
        f = lambda key: float(m[key].replace(",", ""))  # noqa: E731
        ia, lanes = f("smsp__issue_active.avg.pct_of_peak_sustained_active"), f("smsp__thread_inst_executed_per_inst_executed.ratio")
        issue[k] = {"kernel": m["Kernel Name"][:80], "issue_active_pct": ia, "lanes_per_instruction": lanes, "useful_lane_issue_frac": ia / 100 * lanes / 32,
                    "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"), "registers": int(f("launch__registers_per_thread")),
                    "duration_ms": f("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(m["__units__"]["gpu__time_duration.sum"][:2].rstrip("e"), 1.0)}
    if issue:
        t, r = issue.get("k_wf_track"), issue.get("k_wf_tr")
        json.dump({"git": SHA, "csrc_sha16": CSRC,
                   "roofline_issue": {"bound": "issue x lanes", "unit": "fraction of the SM's lane-issue capacity doing work",
                                      "k_wf_track": t, "k_wf_tr": r,
                                      "note": "the tracking kernels are latency / divergence bound, not HBM bound: issue-slot utilisation x active lanes per "
                                              "instruction of one full-size launch each (ncu --set full)"},
                   "kernels": issue}, open(os.path.join(PRO, f"{R}_issue.json"), "w"), indent=1)
    if os.path.exists(os.path.join(OUT, "launches.csv")):
        open(os.path.join(PRO, f"{R}_launches.txt"), "w").write(STAMP + "# ncu launch list (gpu__time_duration.sum, serialised, cold caches) of bench.py --steps 1 --warmup 0: 3 frames (timed + the 2 of the CUDA-event pass)\n"
                                                                  + run(sys.executable, "tools/launch_summary.py", os.path.join(OUT, "launches.csv")))
    if os.path.exists(os.path.join(OUT, "bytes.csv")):
        tj = os.path.join(PRO, f"{R}_traffic.json")
        txt = run(sys.executable, "tools/bytes_summary.py", os.path.join(OUT, "bytes.csv"), tj)
        open(os.path.join(PRO, f"{R}_dram_bytes.txt"), "w").write(STAMP + "# DRAM bytes and duration per kernel over the same 3 frames (divide by 3 for one frame)\n" + txt)
        d = json.load(open(tj))
        d.update({"git": SHA, "csrc_sha16": CSRC})
        json.dump(d, open(tj, "w"), indent=1)
    for src, dst in (("bench.json", f"{R}_bench_1gpu.json"), ("bench_ref.json", f"{R}_bench_reference_arm.json"), ("bench_ref_faithful.json", f"{R}_bench_reference_arm_faithful_rng.json"),
                     (f"{R}_configs.jsonl", f"{R}_configs.jsonl"), ("pytest_gpu.log", f"{R}_pytest_gpu.log"), ("smoke.log", f"{R}_smoke.log"),
                     ("sanitizer_memcheck.log", f"{R}_sanitizer_memcheck.log"), ("sanitizer_synccheck.log", f"{R}_sanitizer_synccheck.log"),
                     ("sanitizer_initcheck.log", f"{R}_sanitizer_initcheck.log"), ("sanitizer_racecheck.log", f"{R}_sanitizer_racecheck.log"),
                     ("inproc_multi.jsonl", f"{R}_inproc_multi_8gpu.jsonl")):
        p = os.path.join(OUT, src)
        if os.path.exists(p) and os.path.getsize(p) > 0:
            shutil.copy(p, os.path.join(PRO, dst))
    for n in (1, 2, 4, 8):
        p = os.path.join(OUT, f"scale_{n}.json")
        if os.path.exists(p) and os.path.getsize(p) > 0:
            shutil.copy(p, os.path.join(PRO, f"{R}_scale_{n}gpu.json"))
    print("profiles written for", SHA, CSRC)


if __name__ == "__main__":
    main()
