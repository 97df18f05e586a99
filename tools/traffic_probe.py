#!/usr/bin/env python3
"""One forward + one sorted backward (sort plan + segmented reduce) at the bench shape, to be run under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` — the source of profiles/traffic.json.
    python tools/traffic_probe.py <tables> <alpha>"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from param_b200 import ops  # noqa: E402
from param_b200.compute.pt.pytorch_emb import zipf_cdf  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 1.15
rows, B, L, D = 1_000_000, 65536, 20, 128
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
for _ in range(2):          # the second round is the one to read (first-use allocations are over)
    ops.tbe_forward(arena, idx, off, B, out=out)
    ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, out, scale=-1e-6, algo="sorted",
                     max_table_rows=rows)
torch.cuda.synchronize()
print("done")
