#!/usr/bin/env python
"""Registers / stack / spills per kernel from `make -C polyred_b200/csrc ptxas-info` output (stdin or a file)."""
import re, subprocess, sys
t = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers")
for n, st, ss, sl, r in pat.findall(t):
    d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    d = re.sub(r"\(.*", "", d).replace("void prc::", "")
    print(f"{d[:64]:64s} regs {r:>3} stack {st:>4} spill st/ld {ss}/{sl}")
