#!/bin/bash
# Round 2, N=1 (one call): full GPU suite + smoke, knob trials for the sorted backward (L2 policies in the segmented
# reduce, un-aggregated first-pass histogram), bench.py (default flags) with its ncu launch list, DRAM traffic per
# kernel at the bench shape (-> profiles/traffic.json), ncu --set full of the forward / sort / reduce kernels, and
# the reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 500 python -m pytest tests -m gpu -q --maxfail 10 --timeout 150 -p no:cacheprovider > $O/r02n_tests_all.log 2>&1
echo "all gpu tests rc=$?" | tee -a $O/r02n_tests_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02n_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $O/r02n_smoke.log
timeout 150 python tools/sort_bench.py 64 1.15 > $O/r02n_sort_a1.15.log 2>&1
PB200_SEG_L2HINT=1 timeout 150 python tools/sort_bench.py 64 1.15 > $O/r02n_sort_a1.15_l2hint.log 2>&1
PB200_SORT_HIST_PLAIN=1 timeout 150 python tools/sort_bench.py 64 1.15 > $O/r02n_sort_a1.15_histplain1.log 2>&1
PB200_SORT_HIST_PLAIN=3 timeout 150 python tools/sort_bench.py 64 1.15 > $O/r02n_sort_a1.15_histplain3.log 2>&1
timeout 150 python tools/sort_bench.py 64 0 > $O/r02n_sort_a0.log 2>&1
PB200_SEG_L2HINT=1 PB200_SORT_HIST_PLAIN=1 timeout 150 python tools/sort_bench.py 64 0 > $O/r02n_sort_a0_l2hint_histplain1.log 2>&1
timeout 900 python bench.py > $O/r02n_bench_n1.log 2> $O/r02n_bench_n1.err
echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > $O/r02n_bench_reference.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02n_launches_bench_py.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-uniform --quick > $O/r02n_bench_under_ncu.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 400 ncu --metrics $M --clock-control none -k regex:'tbe_fwd|radix_|segment_reduce' --csv --log-file $O/r02n_traffic_zipf.csv \
    python tools/traffic_probe.py 256 1.15 > $O/r02n_traffic_zipf.log 2>&1
timeout 400 ncu --metrics $M --clock-control none -k regex:'tbe_fwd|radix_|segment_reduce' --csv --log-file $O/r02n_traffic_uniform.csv \
    python tools/traffic_probe.py 256 0 > $O/r02n_traffic_uniform.log 2>&1
python tools/traffic_from_ncu.py $O/r02n_traffic_zipf.csv $O/r02n_traffic_uniform.csv > $O/r02n_traffic.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'tbe_fwd_direct|radix_|segment_reduce' -s 2 -c 8 -f \
    -o $O/r02n_full_16tables python tools/prof_kernels.py --tables 16 --iters 3 --what fwd_direct,bwd_sorted > $O/r02n_ncu_full.log 2>&1
for f in $O/r02n_tests_all.log $O/r02n_smoke.log $O/r02n_sort_*.log $O/r02n_bench_reference.log; do echo "== $f"; tail -n 3 $f | cut -c1-600; done
cat $O/r02n_traffic.json | head -c 1500
tail -c 2500 $O/r02n_bench_n1.log; tail -n 3 $O/r02n_bench_n1.err
