#!/usr/bin/env python
"""Per-source-line executed warp instructions / stall samples of one kernel from an .ncu-rep (source page).
usage: ncu_lines.py report.ncu-rep kernel_regex [top_n] [launch_substring]"""
import csv, subprocess, sys, io, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
want = sys.argv[4] if len(sys.argv) > 4 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] == "File Path":
        cur = {"file": row[1], "rows": [], "fn": None, "hdr": None}; blocks.append(cur)
    elif row[0] == "Function Name" and cur is not None:
        cur["fn"] = row[1]
    elif row[0] == "Line No" and cur is not None:
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None:
        if row[0] != "": cur["rows"].append(row)
for b in blocks:
    if want and want not in (b["fn"] or ""):
        continue
    h = b["hdr"]; iI = h.index("Instructions Executed"); iS = h.index("# Samples"); iT = h.index("Thread Instructions Executed")
    b["rows"] = [r for r in b["rows"] if len(r) > max(iI, iS, iT) and (r[iI] or "0").isdigit()]
    tot = sum(int(r[iI] or 0) for r in b["rows"]); tots = sum(int(r[iS] or 0) for r in b["rows"])
    print(f"=== {b['fn'][:110]}  [{b['file'].split('/')[-1]}]  warp-inst {tot}  samples {tots}")
    rows = sorted(b["rows"], key=lambda r: -int(r[iI] or 0))[:top]
    for r in rows:
        n = int(r[iI] or 0)
        if n == 0: break
        print(f"  L{r[0]:>5} {100*n/max(tot,1):5.1f}% inst {100*int(r[iS] or 0)/max(tots,1):5.1f}% smp  thr/inst {int(r[iT] or 0)/max(n,1):4.1f}  | {r[1].strip()[:110]}")
