#!/bin/bash
# Last GPU pass of round 1: smoke(), the whole GPU suite (256-entry segments are the default now; new: aten
# override, comms_compute span), one timing of the backward variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 python __graft_entry__.py smoke > gpurun_out/r01k_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/r01k_smoke.log
timeout 170 python -m pytest tests -m gpu -q --maxfail 8 --timeout 100 -p no:cacheprovider > gpurun_out/r01k_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee -a gpurun_out/r01k_tests_all.log
timeout 40 python tools/variant_bench.py 64 1.15 > gpurun_out/r01k_variant_zipf.log 2>&1
tail -n 3 gpurun_out/r01k_smoke.log
tail -n 12 gpurun_out/r01k_tests_all.log
tail -n 2 gpurun_out/r01k_variant_zipf.log
