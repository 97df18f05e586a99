#!/bin/bash
# Second GPU-box pass of the round-1 additions (after the E1 prefetch / red rewrite and the lazy-load fix).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --maxfail 8 --timeout 120 -p no:cacheprovider > gpurun_out/r01e_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee -a gpurun_out/r01e_tests_all.log
timeout 120 python tools/variant_bench.py 64 1.15 > gpurun_out/r01e_variant_zipf.log 2>&1
timeout 120 python tools/variant_bench.py 64 0 > gpurun_out/r01e_variant_uniform.log 2>&1
PB200_SEG=256 timeout 120 python tools/variant_bench.py 64 1.15 > gpurun_out/r01e_variant_zipf_seg256.log 2>&1
PB200_SEG=256 timeout 120 python tools/variant_bench.py 64 0 > gpurun_out/r01e_variant_uniform_seg256.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on \
    -k regex:'exact_reduce_kernel|exact_boundary_kernel' -s 2 -c 2 -f -o gpurun_out/r01e_exact_adagrad \
    python tools/prof_kernels.py --tables 16 --what bwd_adagrad --iters 2 > gpurun_out/r01e_ncu.log 2>&1
echo "ncu rc=$?"
timeout 300 python bench.py > gpurun_out/r01e_bench_n1.log 2>&1
echo "bench rc=$?"
timeout 150 python bench.py --rows 10000000 --tables 25 --skip-cpu --skip-e2e > gpurun_out/r01e_bench_n1_10Mrows_25tables.log 2>&1
echo "bench 10M rc=$?"
tail -n 6 gpurun_out/r01e_tests_all.log
tail -n 2 gpurun_out/r01e_variant_*.log
tail -c 1500 gpurun_out/r01e_bench_n1.log
tail -c 800 gpurun_out/r01e_bench_n1_10Mrows_25tables.log
