#!/bin/bash
# Round 2, N=1: forward lane-group trials (one / two / four bags per warp) with bit-exactness tests under each setting,
# the sorted backward with the new defaults, and ncu --set full of the two collective kernels run at world size 1
# (self-exchange: the same code path, HBM instead of NVLink — multi-rank commands are not run under ncu).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export RANK=0 LOCAL_RANK=0 WORLD_SIZE=1 MASTER_ADDR=127.0.0.1 MASTER_PORT=29741
for g in 16 8; do
  PB200_FWD_GROUP=$g timeout 200 python -m pytest tests/test_gpu_embbag.py -q --timeout 120 -p no:cacheprovider \
      -k "forward or large_shape or golden or module_matches or host_buffer" > $O/r02o_tests_group$g.log 2>&1
  echo "group $g tests rc=$?" | tee -a $O/r02o_tests_group$g.log
done
for a in 1.15 0; do
  for g in 32 16 8; do
    PB200_FWD_GROUP=$g PB200_SORT_BENCH_FWD_ONLY=1 timeout 120 python tools/sort_bench.py 64 $a > $O/r02o_fwd_a${a}_group$g.log 2>&1
  done
done
PB200_FWD_GROUP=16 PB200_FWD_OCC5=1 PB200_SORT_BENCH_FWD_ONLY=1 timeout 120 python tools/sort_bench.py 64 1.15 > $O/r02o_fwd_a1.15_group16_occ5.log 2>&1
timeout 150 python tools/sort_bench.py 64 1.15 > $O/r02o_sort_a1.15_defaults.log 2>&1
timeout 150 python tools/sort_bench.py 64 0 > $O/r02o_sort_a0_defaults.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:a2a_push_kernel -s 2 -c 1 -f -o $O/r02o_a2a_push_w1 \
    python -m param_b200.comms.pt.comms --collective all_to_all_single --begin-size 256M --end-size 256M --num-iters 3 \
    --num_warmup_iters 2 --backend b200 --json > $O/r02o_ncu_a2a_push.log 2>&1
echo "ncu a2a_push rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tbe_fwd_a2a_kernel -s 2 -c 1 -f -o $O/r02o_fwd_a2a_w1 \
    python -m param_b200.comms.pt.dlrm --mini-batch-size 65536 --num-batches 2 --warmup-batches 2 --arch-embedding-size 1000000x16 \
    --arch-sparse-feature-size 128 --num-indices-per-lookup 20 --alpha 1.15 --json > $O/r02o_ncu_fwd_a2a.log 2>&1
echo "ncu fwd_a2a rc=$?"
for f in $O/r02o_*.log; do echo "== $f"; tail -n 2 $f | cut -c1-400; done
