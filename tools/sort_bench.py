#!/usr/bin/env python3
"""Times the pieces of the sort-based backward on a cfg2-shaped workload (T tables x 1 M rows x 128, batch
65536, bag 20): the sort plan alone (pb200_tbe_plan_build), the segmented reduce alone (plan ready), the
inline backward (sort + reduce on one stream), forward alone, and the step with the plan built on a side
stream while the forward runs.  Knobs (PB200_SORT_DIGIT / _TILE / _GROUP) are read once per process by the
library, so call it once per setting.

    python tools/sort_bench.py [tables=64] [alpha=1.15] [rows=1000000]
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from param_b200 import ops  # noqa: E402
from param_b200.compute.pt.pytorch_emb import zipf_cdf  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 1.15
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
B, L, D = 65536, 20, 128
dev = torch.device("cuda:0")
arena = ops.TableArena.allocate([rows] * T, D, dev)
ops.fill_uniform_(arena.weights, -1e-3, 1e-3, seed=1)
idx = torch.empty(T * B * L, dtype=torch.int64, device=dev)
cdf = (torch.from_numpy(zipf_cdf(alpha, rows)).to(dev) if alpha > 0 else
       torch.linspace(1.0 / rows, 1.0, rows, dtype=torch.float64, device=dev))
for t in range(T):
    ops.fill_zipf_indices_(idx[t * B * L:(t + 1) * B * L], L, cdf, seed=1000 + t, dedupe=alpha > 0)
off = torch.arange(T * B + 1, dtype=torch.int64, device=dev) * L
out = torch.empty((B, T * D), device=dev)
side = torch.cuda.Stream(device=dev)


def ev(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


res = {}
if os.environ.get("PB200_SORT_BENCH_FWD_ONLY"):
    f = ev(lambda: ops.tbe_forward(arena, idx, off, B, out=out, algo="direct"), n=10)
    knobs = {k: v for k, v in os.environ.items() if k.startswith("PB200_")}
    print(f"T={T} alpha={alpha} rows={rows} knobs={knobs} fwd_direct {f:.3f} ms")
    sys.exit(0)
for exact in (False, True):
    tag = "exact" if exact else "sorted"
    buf = ops.tbe_plan(arena.row_offsets, T, D, idx, off, B, rows, exact=exact).buf
    res[f"plan_{tag}"] = ev(lambda: ops.tbe_plan(arena.row_offsets, T, D, idx, off, B, rows, exact=exact, buf=buf))
    plan = ops.tbe_plan(arena.row_offsets, T, D, idx, off, B, rows, exact=exact, buf=buf)
    res[f"reduce_{tag}"] = ev(lambda: ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, out,
                                                       scale=-1e-6, algo=tag, max_table_rows=rows, plan=plan))
    res[f"inline_{tag}"] = ev(lambda: ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, out,
                                                       scale=-1e-6, algo=tag, max_table_rows=rows))
res["fwd"] = ev(lambda: ops.tbe_forward(arena, idx, off, B, out=out))
res["fwd_direct"] = ev(lambda: ops.tbe_forward(arena, idx, off, B, out=out, algo="direct"))
buf = ops.tbe_plan(arena.row_offsets, T, D, idx, off, B, rows).buf


def step_overlap():
    plan = ops.tbe_plan(arena.row_offsets, T, D, idx, off, B, rows, stream=side, buf=buf)
    ops.tbe_forward(arena, idx, off, B, out=out)
    ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, out, scale=-1e-6, algo="sorted",
                     max_table_rows=rows, plan=plan)


def step_serial():
    ops.tbe_forward(arena, idx, off, B, out=out)
    ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, out, scale=-1e-6, algo="sorted",
                     max_table_rows=rows)


res["step_serial"] = ev(step_serial)
res["step_presort_side_stream"] = ev(step_overlap)
knobs = {k: v for k, v in os.environ.items() if k.startswith("PB200_")}
lookups = T * B * L
print(f"T={T} alpha={alpha} rows={rows} knobs={knobs} " + "  ".join(f"{k} {v:.3f}" for k, v in res.items())
      + f"  | plan {lookups / res['plan_sorted'] / 1e6:.1f} G keys/s, x{256 // T} -> step "
      f"{res['step_presort_side_stream'] * 256 / T:.1f} ms at 256 tables")
