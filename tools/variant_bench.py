#!/usr/bin/env python3
"""Time forward / backward on a cfg2-shaped workload with fewer tables; the kernel variant is chosen
by environment knobs (PB200_FWD_OCC5, PB200_SEG, PB200_SEG_OCC4, PB200_SORT_BITS) read by the library."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from param_b200 import ops  # noqa: E402
from param_b200.compute.pt.pytorch_emb import zipf_cdf  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
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


f = ev(lambda: ops.tbe_forward(arena, idx, off, B, out=out))
bw = ev(lambda: ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, out, scale=-1e-6, algo="sorted"))
knobs = {k: v for k, v in os.environ.items() if k.startswith("PB200_")}
if os.environ.get("PB200_VARIANT_EXTRAS", "1") != "0":
    # fused-optimizer ("exact") backward and fp16 tables on the same request
    lookups = T * B * L
    ex = ev(lambda: ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, out, scale=-1e-6, algo="exact"))
    state = torch.zeros(arena.total_rows, device=dev)
    ada = ev(lambda: ops.tbe_backward_fused(arena.weights, arena.row_offsets, T, D, idx, off, B, out,
                                            optimizer="exact_row_wise_adagrad", lr=1e-6, state=state))
    a16 = ops.TableArena(arena.weights.to(torch.float16), arena.row_offsets, arena.rows, D)
    f16 = ev(lambda: ops.tbe_forward(a16, idx, off, B, out=out))
    ada16 = ev(lambda: ops.tbe_backward_fused(a16.weights, a16.row_offsets, T, D, idx, off, B, out,
                                              optimizer="exact_row_wise_adagrad", lr=1e-6, state=state,
                                              stochastic_rounding=True, sr_seed=7))
    print(f"T={T} alpha={alpha} extras: bwd_exact_sgd {ex:.3f} ms  bwd_exact_rowwise_adagrad {ada:.3f} ms  "
          f"fwd_fp16 {f16:.3f} ms ({lookups / f16 / 1e6:.2f} G lookups/s)  "
          f"bwd_fp16_rowwise_adagrad_sr {ada16:.3f} ms")
print(f"T={T} alpha={alpha} knobs={knobs}  fwd {f:.3f} ms  bwd_sorted {bw:.3f} ms  (x{256 // T} -> {f * 256 / T:.1f} / {bw * 256 / T:.1f} ms at 256 tables)")
