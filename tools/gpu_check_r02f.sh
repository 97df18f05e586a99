#!/bin/bash
# Round 2: HOT forward (head-of-table rows cached in shared memory by cp.async.bulk, one persistent CTA per SM)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_embbag.py -q -x --timeout 120 -p no:cacheprovider > gpurun_out/r02f_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r02f_tests.log
for a in 1.15 0; do
  timeout 120 python tools/sort_bench.py 64 $a > gpurun_out/r02f_a$a.log 2>&1
  PB200_FWD_HOT_ROWS=128 timeout 120 python tools/sort_bench.py 64 $a > gpurun_out/r02f_a${a}_k128.log 2>&1
  PB200_FWD_HOT_ROWS=320 timeout 120 python tools/sort_bench.py 64 $a > gpurun_out/r02f_a${a}_k320.log 2>&1
done
PB200_FWD_HOT_ROWS=64 timeout 120 python tools/sort_bench.py 64 1.15 > gpurun_out/r02f_a1.15_k64.log 2>&1
timeout 200 python tools/sort_bench.py 16 1.15 10000000 > gpurun_out/r02f_10Mrows.log 2>&1
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
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tbe_fwd' -s 2 -c 2 -f \
    -o gpurun_out/r02f_fwd python /tmp/prof_fwd.py > gpurun_out/r02f_ncu.log 2>&1
tail -n 4 gpurun_out/r02f_tests.log
tail -n 1 gpurun_out/r02f_a*.log gpurun_out/r02f_10Mrows.log
