#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` output: hottest SASS instructions by executed count and by
stall samples, plus an opcode histogram.  Usage: tools/ncu_src.py report.ncu-rep [topN]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
si, ei, st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
body = [r for r in rows[h + 1:] if len(r) > max(si, ei, st)]
tot = sum(int(r[ei] or 0) for r in body)
tots = sum(int(r[st] or 0) for r in body)
print(f"kernel: {rows[0][1][:100]}\nSASS instructions: {len(body)}  warp-instr executed: {tot}  stall samples: {tots}")
ops = collections.Counter()
for r in body:
    op = r[si].split()[0] if not r[si].startswith("@") else r[si].split()[1]
    ops[op.split(".")[0]] += int(r[ei] or 0)
print("opcode mix:", ", ".join(f"{k} {v / tot:.1%}" for k, v in ops.most_common(14)))
print(f"--- top {top} by stall samples")
for r in sorted(body, key=lambda r: -int(r[st] or 0))[:top]:
    print(f"{int(r[st] or 0) / max(tots, 1):6.1%} exec {int(r[ei] or 0):>10}  {r[si][:110]}")
