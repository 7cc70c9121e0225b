#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key metrics, stall reasons and the hottest SASS lines.

usage: python tools/ncu_summary.py <report.ncu-rep> [kernel-regex] [n_hot_lines]
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum", "smsp__warps_eligible.avg.per_cycle_active",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    kre = sys.argv[2] if len(sys.argv) > 2 else None
    nhot = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    extra = ["--kernel-name", f"regex:{kre}"] if kre else []
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"] + extra))))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("=== kernel:", r[ix["Kernel Name"]][:100])
        for k in KEYS:
            if k in ix:
                print(f"  {k:75s} {r[ix[k]]:>16s} {units[ix[k]]}")
        st = []
        for h in hdr:
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    st.append((float(r[ix[h]]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1
        print("  stall samples:", ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:8]))
    # hot SASS of the first matching launch
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--launch-count", "1"] + extra))))
    h = next((i for i, r in enumerate(src) if "Source" in r and "# Samples" in r), None)
    if h is None:
        return
    sx = {n: i for i, n in enumerate(src[h])}
    lines = []
    ops = collections.Counter()
    for r in src[h + 1:]:
        try:
            smp, ins = int(r[sx["# Samples"]]), int(r[sx["Instructions Executed"]])
        except (ValueError, IndexError):
            continue
        lines.append((smp, ins, r[sx["Source"]].strip()))
        t = r[sx["Source"]].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += ins
    ts = sum(l[0] for l in lines) or 1
    ti = sum(l[1] for l in lines) or 1
    print(f"--- hottest SASS by stall samples (total samples {ts}, warp instr {ti}) ---")
    for i, (smp, ins, s) in enumerate(lines):
        lines[i] = (smp, ins, s, i)
    for smp, ins, s, i in sorted(lines, reverse=True)[:nhot]:
        print(f"  {100 * smp / ts:5.1f}% smp {100 * ins / ti:5.2f}% inst  #{i:5d}  {s[:90]}")
    print("--- executed warp instructions by opcode ---")
    print("  " + ", ".join(f"{o} {100 * n / ti:.1f}%" for o, n in ops.most_common(18)))


if __name__ == "__main__":
    main()
