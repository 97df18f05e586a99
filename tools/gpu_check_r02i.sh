#!/bin/bash
# Round 2, N=1: L1-policy variant of the DIRECT forward (head rows evict_last, cold rows no_allocate), the reference's
# compute/python framework with the B200 operator, compute-sanitizer on the peer-window protocol tests, full suite,
# bench.py (default flags) and its ncu launch list + DRAM traffic per kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
for k in 0 128 256 512; do
  PB200_FWD_HEAD_ROWS=$k PB200_SORT_BENCH_FWD_ONLY=1 timeout 120 python tools/sort_bench.py 64 1.15 > $O/r02i_head${k}_zipf.log 2>&1
done
PB200_FWD_HEAD_ROWS=256 PB200_SORT_BENCH_FWD_ONLY=1 timeout 120 python tools/sort_bench.py 64 0 > $O/r02i_head256_uniform.log 2>&1
timeout 500 python -m pytest tests -m gpu -q --maxfail 10 --timeout 120 -p no:cacheprovider > $O/r02i_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee -a $O/r02i_tests_all.log
timeout 200 python -m param_b200.integration.param_plugin bench -c param_b200/compute/configs/b200_batched_embedding_bag.json \
    -d cuda -w 2 -i 5 -b --cuda-l2-cache on -o $O/r02i_compute_python > $O/r02i_compute_python.log 2>&1
echo "compute/python rc=$?" | tee -a $O/r02i_compute_python.log
timeout 120 python -m param_b200.integration.param_plugin emb --steps 20 --warmups 3 --device gpu emb --dataset B > $O/r02i_emb_driver.log 2>&1
echo "emb driver rc=$?" | tee -a $O/r02i_emb_driver.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_a2a_local.py -q -x --timeout 380 -p no:cacheprovider \
    -k "golden or zero_copy or staged_output or list_form" > $O/r02i_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a $O/r02i_sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_sort_plan.py -q -x --timeout 280 -p no:cacheprovider \
    -k "fixed_size or side_arrays" > $O/r02i_sanitizer_synccheck_sort.log 2>&1
echo "synccheck rc=$?" | tee -a $O/r02i_sanitizer_synccheck_sort.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_sort_plan.py -q -x --timeout 280 -p no:cacheprovider \
    -k "fixed_size" > $O/r02i_sanitizer_racecheck_sort.log 2>&1
echo "racecheck rc=$?" | tee -a $O/r02i_sanitizer_racecheck_sort.log
for f in $O/r02i_*.log; do echo "== $f"; tail -n 5 $f | cut -c1-500; done
