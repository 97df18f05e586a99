#!/bin/bash
# Round 2, third GPU pass: ballot-ranked sort (512 threads x 8 items), window staging fix, list all_to_all, new bench.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sort_plan.py tests/test_gpu_a2a_local.py -q -x --timeout 120 -p no:cacheprovider > gpurun_out/r02c_tests_new.log 2>&1
echo "new tests rc=$?" | tee -a gpurun_out/r02c_tests_new.log
timeout 400 python -m pytest tests -m gpu -q --maxfail 10 --timeout 120 -p no:cacheprovider > gpurun_out/r02c_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee -a gpurun_out/r02c_tests_all.log
for a in 1.15 0; do
  timeout 120 python tools/sort_bench.py 64 $a > gpurun_out/r02c_sort_a$a.log 2>&1
  PB200_SORT_DIGIT=7 timeout 120 python tools/sort_bench.py 64 $a > gpurun_out/r02c_sort_a${a}_digit7.log 2>&1
  PB200_SORT_DIGIT=11 timeout 120 python tools/sort_bench.py 64 $a > gpurun_out/r02c_sort_a${a}_digit11.log 2>&1
done
timeout 120 python tools/sort_bench.py 16 1.15 10000000 > gpurun_out/r02c_sort_10Mrows.log 2>&1
timeout 600 python bench.py --steps 5 > gpurun_out/r02c_bench.log 2> gpurun_out/r02c_bench.err
echo "bench rc=$?"
tail -n 4 gpurun_out/r02c_tests_new.log
tail -n 12 gpurun_out/r02c_tests_all.log
tail -n 1 gpurun_out/r02c_sort_*.log
tail -c 6000 gpurun_out/r02c_bench.log; tail -n 5 gpurun_out/r02c_bench.err
