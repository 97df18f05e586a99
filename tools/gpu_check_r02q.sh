#!/bin/bash
# Round 2, N=1: full suite with the new defaults (two bags per warp in the forward and the fused forward, un-aggregated
# histograms, L2 policies), two-segments-per-warp trial for the reduce, the fused lookup+exchange kernel at world 1 with
# one / two bags per warp, then bench.py (default flags) and its ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 500 python -m pytest tests -m gpu -q --maxfail 10 --timeout 150 -p no:cacheprovider > $O/r02q_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee -a $O/r02q_tests_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02q_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $O/r02q_smoke.log
for a in 1.15 0; do
  timeout 150 python tools/sort_bench.py 64 $a > $O/r02q_sort_a$a.log 2>&1
  PB200_SEG_GROUP=16 timeout 150 python tools/sort_bench.py 64 $a > $O/r02q_sort_a${a}_seggroup16.log 2>&1
done
export RANK=0 LOCAL_RANK=0 WORLD_SIZE=1 MASTER_ADDR=127.0.0.1 MASTER_PORT=29751
for g in 32 16; do
  PB200_FUSED_GROUP=$g timeout 200 python -m param_b200.comms.pt.dlrm --mini-batch-size 65536 --num-batches 6 --warmup-batches 2 \
      --arch-embedding-size 1000000x64 --arch-sparse-feature-size 128 --num-indices-per-lookup 20 --alpha 1.15 --json \
      > $O/r02q_dlrm_w1_fusedgroup$g.log 2>&1
done
unset RANK LOCAL_RANK WORLD_SIZE MASTER_ADDR MASTER_PORT
timeout 900 python bench.py > $O/r02q_bench_n1.log 2> $O/r02q_bench_n1.err
echo "bench rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02q_launches_bench_py.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-uniform --quick > $O/r02q_bench_under_ncu.log 2>&1
for f in $O/r02q_tests_all.log $O/r02q_smoke.log $O/r02q_sort_*.log $O/r02q_dlrm_w1_*.log; do echo "== $f"; tail -n 2 $f | cut -c1-700; done
tail -c 1500 $O/r02q_bench_n1.log; tail -n 3 $O/r02q_bench_n1.err
