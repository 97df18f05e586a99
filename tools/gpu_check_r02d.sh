#!/bin/bash
# Round 2, second GPU pass: ncu --set full of the six kernels of one sort-plan build (2 passes) and of the
# segmented reduce, 16 tables, Zipf and uniform.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/prof_plan.py <<'PY'
import sys
sys.path.insert(0, ".")
import torch
from param_b200 import ops
from param_b200.compute.pt.pytorch_emb import zipf_cdf
alpha = float(sys.argv[1]); T = 16
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
out = torch.randn((B, T * D), device=dev)
for _ in range(3):
    ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, out, scale=-1e-6, algo="sorted", max_table_rows=rows)
torch.cuda.synchronize()
print("done")
PY
for a in 1.15; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'radix_|segment_reduce' -s 7 -c 7 -f \
    -o gpurun_out/r02d_sort_a$a python /tmp/prof_plan.py $a > gpurun_out/r02d_ncu_a$a.log 2>&1
echo "ncu a=$a rc=$?"
done
ls -la gpurun_out/*.ncu-rep
