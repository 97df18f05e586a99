#!/bin/bash
# Fourth GPU-box pass: hybrid boundary handling (lane-group E2 + CTA work-list E3), E1 at 3 CTAs/SM by default.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --maxfail 8 --timeout 120 -p no:cacheprovider > gpurun_out/r01g_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee -a gpurun_out/r01g_tests_all.log
timeout 100 python tools/variant_bench.py 64 1.15 > gpurun_out/r01g_variant_zipf.log 2>&1
timeout 100 python tools/variant_bench.py 64 0 > gpurun_out/r01g_variant_uniform.log 2>&1
PB200_EXACT_OCC3=0 timeout 100 python tools/variant_bench.py 64 1.15 > gpurun_out/r01g_variant_zipf_occ2.log 2>&1
PB200_EXACT_OCC3=0 timeout 100 python tools/variant_bench.py 64 0 > gpurun_out/r01g_variant_uniform_occ2.log 2>&1
timeout 120 python bench.py --rows 10000000 --tables 25 --skip-cpu --skip-e2e > gpurun_out/r01g_bench_n1_10Mrows_25tables.log 2>&1
PB200_CHUNK_MIN_PAIRS=0 timeout 120 python bench.py --rows 10000000 --tables 25 --skip-cpu --skip-e2e > gpurun_out/r01g_bench_n1_10Mrows_25tables_chunk24bit.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on \
    -k regex:'exact_' -s 3 -c 3 -f -o gpurun_out/r01g_exact_adagrad \
    python tools/prof_kernels.py --tables 16 --what bwd_adagrad --iters 2 > gpurun_out/r01g_ncu.log 2>&1
echo "ncu rc=$?"
tail -n 6 gpurun_out/r01g_tests_all.log
tail -n 2 gpurun_out/r01g_variant_*.log
for f in gpurun_out/r01g_bench_n1_10Mrows_25tables.log gpurun_out/r01g_bench_n1_10Mrows_25tables_chunk24bit.log; do
  python - "$f" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d = json.loads(line)
        print(sys.argv[1], 'step ms', round(d['ms_per_step'], 3), d['kernels'])
PY
done
