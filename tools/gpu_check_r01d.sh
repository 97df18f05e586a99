#!/bin/bash
# One GPU-box pass for the round-1 additions: the whole GPU suite (all failures, not just the first),
# timings of the new kernel variants, one ncu --set full capture of the new kernels.
# Every step is bounded; logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r01d_gpu.txt 2>&1
timeout 400 python -m pytest tests -m gpu -q --maxfail 8 --timeout 120 -p no:cacheprovider > gpurun_out/r01d_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee -a gpurun_out/r01d_tests_all.log
timeout 150 python tools/variant_bench.py 64 1.15 > gpurun_out/r01d_variant_zipf.log 2>&1
echo "variant zipf rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on \
    -k regex:'exact_reduce_kernel|exact_boundary_kernel|tbe_fwd_direct_f16' -c 3 -f -o gpurun_out/r01d_exact_f16 \
    python tools/prof_kernels.py --tables 16 --what bwd_adagrad,fwd_f16 --iters 1 > gpurun_out/r01d_ncu.log 2>&1
echo "ncu rc=$?"
timeout 100 python tools/variant_bench.py 64 0 > gpurun_out/r01d_variant_uniform.log 2>&1
echo "variant uniform rc=$?"
tail -n 12 gpurun_out/r01d_tests_all.log
tail -n 3 gpurun_out/r01d_variant_zipf.log gpurun_out/r01d_variant_uniform.log gpurun_out/r01d_ncu.log
