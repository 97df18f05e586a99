#!/bin/bash
# One GPU-box pass for the round-1 additions: new parity tests first, then timings of the new kernel
# variants, one ncu --set full capture of the new kernels, then the whole GPU suite.
# Every step is bounded; logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r01d_gpu.txt 2>&1
timeout 280 python -m pytest tests/test_gpu_tbe_fused.py tests/test_gpu_a2a_local.py tests/test_gpu_sparse_dist.py \
    -q --timeout 120 -p no:cacheprovider > gpurun_out/r01d_tests_new.log 2>&1
echo "new tests rc=$?" | tee -a gpurun_out/r01d_tests_new.log
timeout 200 python tools/variant_bench.py 64 1.15 > gpurun_out/r01d_variant_zipf.log 2>&1
echo "variant zipf rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on \
    -k regex:'exact_reduce_kernel|exact_boundary_kernel|tbe_fwd_direct_f16' -c 3 -f -o gpurun_out/r01d_exact_f16 \
    python tools/prof_kernels.py --tables 16 --what bwd_adagrad,fwd_f16 --iters 1 > gpurun_out/r01d_ncu.log 2>&1
echo "ncu rc=$?"
timeout 330 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/r01d_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee -a gpurun_out/r01d_tests_all.log
timeout 120 python tools/variant_bench.py 64 0 > gpurun_out/r01d_variant_uniform.log 2>&1
echo "variant uniform rc=$?"
tail -5 gpurun_out/r01d_tests_new.log gpurun_out/r01d_variant_zipf.log gpurun_out/r01d_tests_all.log gpurun_out/r01d_variant_uniform.log
