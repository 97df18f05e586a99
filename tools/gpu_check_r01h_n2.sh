#!/bin/bash
# 2-GPU pass: multi-process parity (tools/dist_check.py incl. the device-side SparseDataDist), the DLRM
# step bench at N = 2, and the comms/compute overlap runner with the emb_lookup kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29611 tools/dist_check.py > gpurun_out/r01h_dist_check_n2.log 2>&1
echo "dist_check rc=$?" | tee -a gpurun_out/r01h_dist_check_n2.log
timeout 240 $TR --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r01h_bench_n2.log 2>&1
echo "bench n2 rc=$?"
timeout 120 $TR --master-port 29613 -m param_b200.comms.pt.comms_compute --mode comms-compute --kernel emb_lookup \
    --collective all_to_all_single --b 16M --e 256M --f 4 --n 10 --w 3 --num-compute 2 --emb-dim 128 \
    --num-embs 1000000 --batch-size 16384 --ntables 64 --bag-size 20 --direction forward \
    > gpurun_out/r01h_comms_compute_n2_fwd.log 2>&1
echo "comms_compute fwd rc=$?"
timeout 120 $TR --master-port 29614 -m param_b200.comms.pt.comms_compute --mode comms-compute --kernel emb_lookup \
    --collective all_to_all_single --b 64M --e 64M --n 10 --w 3 --num-compute 1 --emb-dim 128 \
    --num-embs 1000000 --batch-size 16384 --ntables 64 --bag-size 20 --direction backward \
    > gpurun_out/r01h_comms_compute_n2_bwd.log 2>&1
echo "comms_compute bwd rc=$?"
grep -c PASS gpurun_out/r01h_dist_check_n2.log; grep -v PASS gpurun_out/r01h_dist_check_n2.log | tail -8
tail -c 2500 gpurun_out/r01h_bench_n2.log
tail -n 6 gpurun_out/r01h_comms_compute_n2_fwd.log gpurun_out/r01h_comms_compute_n2_bwd.log
