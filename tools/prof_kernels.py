#!/usr/bin/env python3
"""Launch each hot kernel a few times on a cfg2-shaped workload (fewer tables) so that ncu can pick
them with -k regex:...  Usage (on the GPU box):
  ncu --set full --clock-control none --import-source on -k regex:tbe_fwd_direct -s 2 -c 1 -o gpurun_out/fwd_direct \
      python tools/prof_kernels.py --tables 32
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from param_b200 import ops  # noqa: E402
from param_b200.compute.pt.pytorch_emb import zipf_cdf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tables", type=int, default=32)
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--bag", type=int, default=20)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--alpha", type=float, default=1.15)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--what", default="fwd_direct,fwd_staged,bwd_sorted,bwd_atomic")
a = ap.parse_args()
dev = torch.device("cuda:0")
T, B, L, D = a.tables, a.batch, a.bag, a.dim
arena = ops.TableArena.allocate([a.rows] * T, D, dev)
ops.fill_uniform_(arena.weights, -1e-3, 1e-3, seed=1)
idx = torch.empty(T * B * L, dtype=torch.int64, device=dev)
cdf = (torch.from_numpy(zipf_cdf(a.alpha, a.rows)).to(dev) if a.alpha > 0 else
       torch.linspace(1.0 / a.rows, 1.0, a.rows, dtype=torch.float64, device=dev))
for t in range(T):
    ops.fill_zipf_indices_(idx[t * B * L:(t + 1) * B * L], L, cdf, seed=1000 + t, dedupe=a.alpha > 0)
off = torch.arange(T * B + 1, dtype=torch.int64, device=dev) * L
out = torch.empty((B, T * D), device=dev)
torch.cuda.synchronize()
for what in a.what.split(","):
    for _ in range(a.iters):
        if what == "fwd_direct":
            ops.tbe_forward(arena, idx, off, B, algo="direct", out=out)
        elif what == "fwd_pipelined":
            ops.tbe_forward(arena, idx, off, B, algo="pipelined", out=out)
        elif what == "fwd_staged":
            ops.tbe_forward(arena, idx, off, B, algo="staged", out=out)
        elif what == "bwd_sorted":
            ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, out, scale=-1e-6, algo="sorted",
                             max_table_rows=a.rows)
        elif what == "bwd_exact":
            ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, out, scale=-1e-6, algo="exact")
        elif what == "bwd_adagrad":
            if "state" not in globals():
                state = torch.zeros(arena.total_rows, device=dev)
            ops.tbe_backward_fused(arena.weights, arena.row_offsets, T, D, idx, off, B, out,
                                   optimizer="exact_row_wise_adagrad", lr=1e-6, state=state)
        elif what == "fwd_f16":
            if "a16" not in globals():
                a16 = ops.TableArena(arena.weights.to(torch.float16), arena.row_offsets, arena.rows, D)
            ops.tbe_forward(a16, idx, off, B, out=out)
        elif what == "bwd_atomic":
            ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, out, scale=-1e-6, algo="atomic")
    torch.cuda.synchronize()
print("done")
