#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics per captured launch + top stall lines.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        print("==", row[hdr.index("Kernel Name")][:90])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:85s} {row[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(src)))
        # one table per launch: "Kernel Name" row, header row, instruction rows
        i = 0
        while i < len(rows):
            if rows[i] and rows[i][0] == "Kernel Name":
                name, hdr = rows[i][1], rows[i + 1]
                j = i + 2
                body = []
                while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                    if len(rows[j]) == len(hdr):
                        body.append(rows[j])
                    j += 1
                report(name, hdr, body, n)
                i = j
            else:
                i += 1


def report(name, hdr, body, n):
    col = {c: k for k, c in enumerate(hdr)}
    S = col["Warp Stall Sampling (All Samples)"]
    tot = sum(float(r[S]) for r in body) or 1.0
    inst = sum(float(r[col["Instructions Executed"]]) for r in body)
    tinst = sum(float(r[col["Thread Instructions Executed"]]) for r in body)
    print(f"-- {name[:80]}: {len(body)} SASS instructions, {inst:.3g} warp-instr, lane efficiency {tinst / max(inst, 1) / 32:.1%}, {tot:.0f} stall samples")
    stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    agg = {c: sum(float(r[col[c]]) for r in body) for c in stall_cols}
    print("   stall reasons: " + ", ".join(f"{c[6:]} {v / tot:.1%}" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    byop = {}
    for r in body:
        op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
        if op.startswith("@"):
            op = r[col["Source"]].split()[1]
        op = op.split(".")[0]
        a = byop.setdefault(op, [0.0, 0.0])
        a[0] += float(r[S])
        a[1] += float(r[col["Instructions Executed"]])
    print("   by opcode (stall share / instr share): " + ", ".join(f"{op} {v[0] / tot:.1%}/{v[1] / max(inst, 1):.1%}" for op, v in sorted(byop.items(), key=lambda kv: -kv[1][0])[:14]))
    print(f"   top {n} instructions by stall samples:")
    for r in sorted(body, key=lambda r: -float(r[S]))[:n]:
        print(f"     {float(r[S]) / tot:6.1%}  exec {float(r[col['Instructions Executed']]):>10.0f}  thr/inst {r[col['Avg. Threads Executed']]:>5s}  {r[col['Source']].strip()[:90]}")


if __name__ == "__main__":
    main()
