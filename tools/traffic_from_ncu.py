#!/usr/bin/env python3
"""profiles/traffic.json from the ncu CSV logs of tools/traffic_probe.py (Zipf and uniform): DRAM bytes per launch of
the forward kernel, the sort-plan kernels (summed: one plan build) and the segmented reduce — second round only.
    python tools/traffic_from_ncu.py zipf.csv uniform.csv > profiles/traffic.json"""
import collections
import csv
import json
import sys


def load(path):
    per = collections.OrderedDict()
    for r in csv.reader(open(path, errors="ignore")):
        if len(r) > 14 and r[0].isdigit():
            k = int(r[0])
            name = r[4].split("(")[0].replace("void ", "").replace("pb200::", "")
            d = per.setdefault(k, {"name": name})
            d[r[12]] = float(r[14].replace(",", ""))
    return list(per.values())


def summarize(launches):
    fwd = [x for x in launches if x["name"].startswith("tbe_fwd")]
    red = [x for x in launches if x["name"].startswith("segment_reduce")]
    rad = [x for x in launches if x["name"].startswith("radix_")]
    def b(x): return x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0)
    def t(x): return x.get("gpu__time_duration.sum", 0)
    n_plans = max(1, len(red))
    last_plan = rad[len(rad) - len(rad) // n_plans:]
    return {"fwd": b(fwd[-1]), "fwd_ms_under_ncu": t(fwd[-1]) / 1e6,
            "bwd_reduce": b(red[-1]), "bwd_reduce_ms_under_ncu": t(red[-1]) / 1e6,
            "bwd_sort_plan": sum(b(x) for x in last_plan), "bwd_sort_plan_ms_under_ncu": sum(t(x) for x in last_plan) / 1e6,
            "bwd": b(red[-1]) + sum(b(x) for x in last_plan), "sort_plan_launches": len(last_plan)}


z, u = summarize(load(sys.argv[1])), summarize(load(sys.argv[2]))
out = {"_source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none of "
                  "tools/traffic_probe.py 256 <alpha> (256 tables x 1M rows x 128, batch 65536, bag 20): bytes per launch; "
                  "the sort plan is the sum over its launches; durations under ncu are serialised/cold and only indicative"}
for k, v in z.items():
    out[k] = v
for k, v in u.items():
    out[k + "_uniform"] = v
print(json.dumps(out, indent=1))
