#!/bin/bash
# Round 2, N=1: forward with two bags per warp (PB200_FWD_GROUP=16) — resident CTAs per SM asked of the compiler (4/5/6)
# x rows in flight per lane group (4/8), Zipf and uniform
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PB200_FWD_GROUP=16 PB200_SORT_BENCH_FWD_ONLY=1
for occ in 4 5 6; do for u in 4 8; do
  PB200_FWD_OCC=$occ PB200_FWD_U=$u timeout 120 python tools/sort_bench.py 64 1.15 > $O/r02p_fwd_a1.15_g16_occ${occ}_u$u.log 2>&1
done; done
for cfg in "5 4" "6 4" "5 8" "6 8"; do set -- $cfg
  PB200_FWD_OCC=$1 PB200_FWD_U=$2 timeout 120 python tools/sort_bench.py 64 0 > $O/r02p_fwd_a0_g16_occ$1_u$2.log 2>&1
done
PB200_FWD_OCC=5 PB200_FWD_U=4 timeout 120 python -m pytest tests/test_gpu_embbag.py -q --timeout 120 -p no:cacheprovider \
      -k "forward or large_shape or golden" > $O/r02p_tests_g16_occ5.log 2>&1
for f in $O/r02p_*.log; do echo "== $f"; tail -n 1 $f | cut -c1-300; done
