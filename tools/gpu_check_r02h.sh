#!/bin/bash
# Round 2: HOT forward v2 (tables in lockstep, predicated loads, 768 threads), new rowwise-Adagrad golden on the GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/prof_fwd.py <<'PY'
import sys
sys.path.insert(0, ".")
import torch
from param_b200 import ops
from param_b200.compute.pt.pytorch_emb import zipf_cdf
alpha = 1.15; T = 64
rows, B, L, D = 1_000_000, 65536, 20, 128
dev = torch.device("cuda:0")
arena = ops.TableArena.allocate([rows] * T, D, dev)
ops.fill_uniform_(arena.weights, -1e-3, 1e-3, seed=1)
idx = torch.empty(T * B * L, dtype=torch.int64, device=dev)
cdf = torch.from_numpy(zipf_cdf(alpha, rows)).to(dev)
for t in range(T):
    ops.fill_zipf_indices_(idx[t * B * L:(t + 1) * B * L], L, cdf, seed=1000 + t, dedupe=True)
off = torch.arange(T * B + 1, dtype=torch.int64, device=dev) * L
out = torch.empty((B, T * D), device=dev)
for _ in range(2):
    ops.tbe_forward(arena, idx, off, B, out=out, algo="hot")
    ops.tbe_forward(arena, idx, off, B, out=out, algo="direct")
torch.cuda.synchronize()
PY
timeout 400 python -m pytest tests/test_gpu_embbag.py tests/test_gpu_tbe_fused.py -q -x --timeout 120 -p no:cacheprovider > gpurun_out/r02h_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r02h_tests.log
for a in 1.15 0; do
  timeout 120 python tools/sort_bench.py 64 $a > gpurun_out/r02h_a$a.log 2>&1
done
PB200_FWD_HOT_ROWS=128 timeout 120 python tools/sort_bench.py 64 1.15 > gpurun_out/r02h_a1.15_k128.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tbe_fwd_hot' -s 2 -c 1 -f \
    -o gpurun_out/r02h_fwd python /tmp/prof_fwd.py > gpurun_out/r02h_ncu.log 2>&1
tail -n 4 gpurun_out/r02h_tests.log
tail -n 1 gpurun_out/r02h_a*.log
