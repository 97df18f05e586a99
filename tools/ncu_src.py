#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` output: hottest SASS instructions by executed count and by
stall samples, plus an opcode histogram — one block per kernel in the report.
Usage: tools/ncu_src.py report.ncu-rep [topN]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for n, h in enumerate(heads):
    end = heads[n + 1] if n + 1 < len(heads) else len(rows)
    hdr = rows[h]
    si, ei, st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    # the kernel name line precedes its header
    name = next((r[1] for r in reversed(rows[max(0, h - 3):h]) if len(r) > 1 and "(" in r[1]), "?")
    body = [r for r in rows[h + 1:end] if len(r) > max(si, ei, st) and r[ei].strip().isdigit()]
    tot = sum(int(r[ei] or 0) for r in body)
    tots = sum(int(r[st] or 0) for r in body)
    print(f"kernel: {name[:110]}\nSASS instructions: {len(body)}  warp-instr executed: {tot}  stall samples: {tots}")
    ops = collections.Counter()
    for r in body:
        op = r[si].split()[0] if not r[si].startswith("@") else r[si].split()[1]
        ops[op.split(".")[0]] += int(r[ei] or 0)
    print("opcode mix:", ", ".join(f"{k} {v / max(tot, 1):.1%}" for k, v in ops.most_common(14)))
    print(f"--- top {top} by stall samples")
    for r in sorted(body, key=lambda r: -int(r[st] or 0))[:top]:
        print(f"{int(r[st] or 0) / max(tots, 1):6.1%} exec {int(r[ei] or 0):>10}  {r[si][:110]}")
    print()
