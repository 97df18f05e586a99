#!/bin/bash
# Round 2: sort v3 (compile-time digit width, inline-PTX ballot ranking, 32-bit tile positions)
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
timeout 300 python -m pytest tests/test_gpu_sort_plan.py tests/test_gpu_embbag.py tests/test_gpu_tbe_fused.py -q -x --timeout 120 -p no:cacheprovider > gpurun_out/r02e_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r02e_tests.log
for a in 1.15 0; do
  timeout 120 python tools/sort_bench.py 64 $a > gpurun_out/r02e_sort_a$a.log 2>&1
done
timeout 120 python tools/sort_bench.py 16 1.15 10000000 > gpurun_out/r02e_sort_10Mrows.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'radix_|segment_reduce' -s 7 -c 7 -f \
    -o gpurun_out/r02e_sort_a1.15 python /tmp/prof_plan.py 1.15 > gpurun_out/r02e_ncu.log 2>&1
tail -n 4 gpurun_out/r02e_tests.log
tail -n 1 gpurun_out/r02e_sort_*.log
