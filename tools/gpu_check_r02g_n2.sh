#!/bin/bash
# Round 2, 2 GPUs: parity checks, the UNMODIFIED reference runners (from baseline/_ref) on the b200 backend next
# to their own NCCL path, cfg5 (ET capture of a DLRM step + the reference's comm_replay / et_replay on it), bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 240 "$@" > $O/r02g_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/r02g_$name.log; }

run dist_check $TR --master-port 29701 tools/dist_check.py
# --- config 3: reference comms.py, its own NCCL path vs --backend b200 (plugin registry) ---
for be in nccl b200; do
  run comms_$be $TR --master-port 29702 -m -- param_b200.integration.param_plugin comms --backend $be --device cuda \
      --collective all_to_all_single --b 1K --e 256M --f 4 --z 1 --c 1 --n 20 --w 5
done
run comms_b200_a2a_list $TR --master-port 29703 -m -- param_b200.integration.param_plugin comms --backend b200 --device cuda \
      --collective all_to_all --b 64K --e 16M --f 16 --z 1 --c 1 --n 10 --w 3
run comms_b200_a2av $TR --master-port 29704 -m -- param_b200.integration.param_plugin comms --backend b200 --device cuda \
      --collective all_to_allv --b 64K --e 16M --f 16 --z 1 --c 1 --n 10 --w 3
# --- config 4: reference dlrm.py, stock backend vs b200 ---
DLRM="--device cuda --mini-batch-size 2048 --arch-embedding-size $(python -c "print('-'.join(['200000']*16))") --arch-sparse-feature-size 128 --num-indices-per-lookup 20 --num-indices-per-lookup-fixed True --num-batches 12 --warmup-batches 2"
PB200_PLUGIN_BACKEND=stock run dlrm_stock $TR --master-port 29705 -m -- param_b200.integration.param_plugin dlrm --backend nccl $DLRM
run dlrm_b200 $TR --master-port 29706 -m -- param_b200.integration.param_plugin dlrm --backend nccl $DLRM
# --- config 1: reference compute driver with the module swapped ---
run emb_driver timeout 120 python -m param_b200.integration.param_plugin emb --steps 20 --warmups 3 --device gpu emb --dataset A
# --- config 5: capture on the box, replay with the reference's tools ---
run cfg5_capture $TR --master-port 29707 tools/cfg5_capture.py --out $O/cfg5_trace --tables-per-rank 8 --rows 200000 --dim 128 --local-batch 2048 --bag 20
for be in nccl b200; do
  run cfg5_comm_replay_$be $TR --master-port 29708 -m -- param_b200.integration.param_plugin comm_replay --trace-type et \
      --trace-path $O/cfg5_trace --backend $be --num-replays 5
done
run cfg5_et_replay_stock $TR --master-port 29709 -m -- param_b200.integration.param_plugin et_replay --trace-path $O/cfg5_trace \
      -m full --warmup-iter 2 --iter 5 --backend nccl
run cfg5_et_replay_b200 $TR --master-port 29710 -m -- param_b200.integration.param_plugin et_replay --trace-path $O/cfg5_trace \
      -m full --warmup-iter 2 --iter 5 --backend b200 --replay-config param_b200/et/replay-config-b200-aten.json
run cfg5_et_replay_b200_comp python -m param_b200.integration.param_plugin et_replay --input $O/cfg5_trace/rank-0.json \
      -m comp --warmup-iter 2 --iter 5 --replay-config param_b200/et/replay-config-b200-aten.json
# --- bench (cfg4 shapes at N = 2) ---
timeout 400 $TR --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02g_bench_n2.log 2> $O/r02g_bench_n2.err
echo "bench rc=$?"
rm -rf $O/cfg5_trace/*_resources   # keep the traces, drop the raw tensor dumps
for f in $O/r02g_*.log; do echo "== $f"; tail -n 6 $f | cut -c1-400; done
