#!/bin/bash
# Round 2, N=1: DRAM traffic per kernel at the bench shape (-> profiles/traffic.json), ncu launch list of bench.py,
# the default bench.py run and the reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 400 ncu --metrics $M --clock-control none -k regex:'tbe_fwd|radix_|segment_reduce' --csv --log-file $O/r02j_traffic_zipf.csv \
    python tools/traffic_probe.py 256 1.15 > $O/r02j_traffic_zipf.log 2>&1
timeout 400 ncu --metrics $M --clock-control none -k regex:'tbe_fwd|radix_|segment_reduce' --csv --log-file $O/r02j_traffic_uniform.csv \
    python tools/traffic_probe.py 256 0 > $O/r02j_traffic_uniform.log 2>&1
python tools/traffic_from_ncu.py $O/r02j_traffic_zipf.csv $O/r02j_traffic_uniform.csv > $O/r02j_traffic.json && cp $O/r02j_traffic.json profiles/traffic.json
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02j_launches_bench_py.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-uniform --quick > $O/r02j_bench_under_ncu.log 2>&1
timeout 900 python bench.py > $O/r02j_bench_n1.log 2> $O/r02j_bench_n1.err
echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > $O/r02j_bench_reference.log 2>&1
cat $O/r02j_traffic.json
tail -c 3000 $O/r02j_bench_n1.log; tail -n 3 $O/r02j_bench_n1.err; tail -c 600 $O/r02j_bench_reference.log
