#!/bin/bash
# Final 1-GPU pass of round 1: whole GPU suite (default knobs), the fused-optimizer tests again with the
# slim E1 kernel, timings of both E1 forms, one ncu capture of the slim kernel, the default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --maxfail 8 --timeout 120 -p no:cacheprovider > gpurun_out/r01i_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee -a gpurun_out/r01i_tests_all.log
PB200_EXACT_SLIM=1 timeout 200 python -m pytest tests/test_gpu_tbe_fused.py tests/test_gpu_embbag.py -m gpu -q --maxfail 8 \
    --timeout 120 -p no:cacheprovider > gpurun_out/r01i_tests_slim.log 2>&1
echo "slim tests rc=$?" | tee -a gpurun_out/r01i_tests_slim.log
timeout 100 python tools/variant_bench.py 64 1.15 > gpurun_out/r01i_variant_zipf.log 2>&1
timeout 100 python tools/variant_bench.py 64 0 > gpurun_out/r01i_variant_uniform.log 2>&1
PB200_EXACT_SLIM=1 timeout 100 python tools/variant_bench.py 64 1.15 > gpurun_out/r01i_variant_zipf_slim.log 2>&1
PB200_EXACT_SLIM=1 timeout 100 python tools/variant_bench.py 64 0 > gpurun_out/r01i_variant_uniform_slim.log 2>&1
PB200_EXACT_SLIM=1 timeout 100 ncu --set full --clock-control none --import-source on \
    -k regex:'exact_reduce_slim' -s 1 -c 1 -f -o gpurun_out/r01i_exact_slim_adagrad \
    python tools/prof_kernels.py --tables 16 --what bwd_adagrad --iters 2 > gpurun_out/r01i_ncu.log 2>&1
echo "ncu rc=$?"
timeout 200 python bench.py > gpurun_out/r01i_bench_n1.log 2>&1
echo "bench rc=$?"
tail -n 5 gpurun_out/r01i_tests_all.log gpurun_out/r01i_tests_slim.log
tail -n 2 gpurun_out/r01i_variant_*.log
tail -c 1200 gpurun_out/r01i_bench_n1.log
