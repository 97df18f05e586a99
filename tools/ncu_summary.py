#!/usr/bin/env python3
"""Key metrics of one or more .ncu-rep files as a markdown table (for profiles/).
Usage: tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...] > profiles/rNN_summary.md"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
]
cols = []
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for k, vals in enumerate(rows[2:]):          # one column per captured launch
        if len(vals) != len(hdr):
            continue
        d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
        name = d.get("Kernel Name", ("?", ""))[0].split("(")[0]
        cols.append((f"{rep.split('/')[-1]} #{k}", name, d))
print("| metric | " + " | ".join(f"{c[0]}<br>`{c[1][:60]}`" for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for key, label in WANT:
    cells = []
    for _, _, d in cols:
        v, u = d.get(key, ("n/a", ""))
        try:
            v = f"{float(v):,.3f}".rstrip("0").rstrip(".")
        except ValueError:
            pass
        cells.append(f"{v} {u}".strip())
    print(f"| {label} (`{key}`) | " + " | ".join(cells) + " |")
