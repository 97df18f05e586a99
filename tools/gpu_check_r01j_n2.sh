#!/bin/bash
# 2-GPU pass: comms/compute overlap runner (emb_lookup kernel, forward and backward) and the DLRM
# pattern runner (times the device-side SparseDataDist with the parallel count kernel).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 100 $TR --master-port 29613 -m param_b200.comms.pt.comms_compute --mode comms-compute --kernel emb_lookup \
    --collective all_to_all_single --begin-size 16M --end-size 256M --step-factor 4 --num-iters 10 \
    --num_warmup_iters 3 --num-compute 2 --emb-dim 128 --num-embs 1000000 --batch-size 16384 --ntables 64 \
    --bag-size 20 --direction forward > gpurun_out/r01j_comms_compute_n2_fwd.log 2>&1
echo "comms_compute fwd rc=$?"
timeout 100 $TR --master-port 29614 -m param_b200.comms.pt.comms_compute --mode comms-compute --kernel emb_lookup \
    --collective all_to_all_single --begin-size 64M --end-size 64M --num-iters 10 --num_warmup_iters 3 \
    --num-compute 1 --emb-dim 128 --num-embs 1000000 --batch-size 16384 --ntables 64 --bag-size 20 \
    --direction backward > gpurun_out/r01j_comms_compute_n2_bwd.log 2>&1
echo "comms_compute bwd rc=$?"
timeout 100 $TR --master-port 29615 -m param_b200.comms.pt.dlrm --mini-batch-size 8192 --num-batches 10 \
    --warmup-batches 3 --arch-embedding-size 1000000x128 --arch-sparse-feature-size 128 \
    --num-indices-per-lookup 20 --alpha 1.15 --lr 0.01 > gpurun_out/r01j_dlrm_runner_n2.log 2>&1
echo "dlrm runner rc=$?"
tail -n 8 gpurun_out/r01j_comms_compute_n2_fwd.log gpurun_out/r01j_comms_compute_n2_bwd.log gpurun_out/r01j_dlrm_runner_n2.log
