#!/bin/bash
# Round 2, first GPU pass: the hand-written radix sort (sort plan) — new tests first, then the whole GPU suite,
# smoke, and timings of the sort / reduce / overlapped step at 64 tables with a few knob settings.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sort_plan.py -q -x --timeout 120 -p no:cacheprovider > gpurun_out/r02a_tests_sort.log 2>&1
echo "sort tests rc=$?" | tee -a gpurun_out/r02a_tests_sort.log
timeout 90 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/r02a_smoke.log
timeout 400 python -m pytest tests -m gpu -q --maxfail 10 --timeout 120 -p no:cacheprovider > gpurun_out/r02a_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee -a gpurun_out/r02a_tests_all.log
for a in 1.15 0; do
  timeout 120 python tools/sort_bench.py 64 $a > gpurun_out/r02a_sort_a$a.log 2>&1
  PB200_SORT_GROUP=8 timeout 120 python tools/sort_bench.py 64 $a > gpurun_out/r02a_sort_a${a}_group8.log 2>&1
  PB200_SORT_DIGIT=7 timeout 120 python tools/sort_bench.py 64 $a > gpurun_out/r02a_sort_a${a}_digit7.log 2>&1
  PB200_SORT_TILE=6000 timeout 120 python tools/sort_bench.py 64 $a > gpurun_out/r02a_sort_a${a}_tile6000.log 2>&1
done
tail -n 5 gpurun_out/r02a_tests_sort.log
tail -n 3 gpurun_out/r02a_smoke.log
tail -n 15 gpurun_out/r02a_tests_all.log
tail -n 2 gpurun_out/r02a_sort_*.log
