#!/bin/bash
# Round 2, 2 GPUs: parity checks, the UNMODIFIED reference runners (from baseline/_ref) on the b200 backend next
# to their own NCCL path, cfg5 (ET capture of a DLRM step + the reference's comm_replay / et_replay on it), bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 240 "$@" > $O/r02k_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/r02k_$name.log; }

# --- config 4: reference dlrm.py, stock backend vs b200 (after the gradient-layout change) ---
DLRM="--device cuda --mini-batch-size 2048 --arch-embedding-size $(python -c "print('-'.join(['200000']*16))") --arch-sparse-feature-size 128 --num-indices-per-lookup 20 --num-indices-per-lookup-fixed True --num-batches 12 --warmup-batches 2"
PB200_PLUGIN_BACKEND=stock run dlrm_stock $TR --master-port 29705 -m -- param_b200.integration.param_plugin dlrm --backend nccl $DLRM
run dlrm_b200 $TR --master-port 29706 -m -- param_b200.integration.param_plugin dlrm --backend nccl $DLRM
# --- basic trace with "compute": "emb_lookup" entries through the reference's commsTraceReplay.py ---
run trace_replay_b200 $TR --master-port 29712 -m -- param_b200.integration.param_plugin trace_replay \
      --trace-path param_b200/comms/pt/traces/dlrm_step_basic.json --trace-type basic --backend b200 --device cuda --num-replays 3
# --- config 5: capture on the box, replay with the reference's tools ---
run cfg5_capture $TR --master-port 29707 tools/cfg5_capture.py --out $O/cfg5_trace --tables-per-rank 8 --rows 200000 --dim 128 --local-batch 2048 --bag 20
for be in nccl b200; do
  run cfg5_comm_replay_$be $TR --master-port 29708 -m -- param_b200.integration.param_plugin comm_replay --trace-type et \
      --trace-path $O/cfg5_trace --backend $be --num-replays 5
done
# the reference's own two-step procedure: first let et_replay find the nodes it cannot rebuild (--update-replay-config,
# needs CUDA_LAUNCH_BLOCKING=1, et_replay.py:1393-1412), then replay with the completed skip list
cp param_b200/et/replay-config-stock.json $O/cfg5_replay_config_stock.json
cp param_b200/et/replay-config-b200-aten.json $O/cfg5_replay_config_b200.json
for c in stock b200; do
  for pass in 1 2; do
    CUDA_LAUNCH_BLOCKING=1 run cfg5_skiplist_${c}_$pass python -m param_b200.integration.param_plugin et_replay --input $O/cfg5_trace/rank-0.json \
        -m comp --warmup-iter 1 --iter 1 --replay-config $O/cfg5_replay_config_$c.json --update-replay-config
  done
done
run cfg5_et_replay_stock $TR --master-port 29709 -m -- param_b200.integration.param_plugin et_replay --trace-path $O/cfg5_trace \
      -m full --warmup-iter 2 --iter 5 --backend nccl --replay-config $O/cfg5_replay_config_stock.json
run cfg5_et_replay_b200 $TR --master-port 29710 -m -- param_b200.integration.param_plugin et_replay --trace-path $O/cfg5_trace \
      -m full --warmup-iter 2 --iter 5 --backend b200 --replay-config $O/cfg5_replay_config_b200.json
run cfg5_et_replay_b200_comp python -m param_b200.integration.param_plugin et_replay --input $O/cfg5_trace/rank-0.json \
      -m comp --warmup-iter 2 --iter 5 --replay-config $O/cfg5_replay_config_b200.json
# --- bench (cfg4 shapes at N = 2) ---
rm -rf $O/cfg5_trace/*_resources   # keep the traces, drop the raw tensor dumps
for f in $O/r02k_*.log; do echo "== $f"; tail -n 6 $f | cut -c1-400; done
