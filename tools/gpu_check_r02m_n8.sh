#!/bin/bash
# Round 2, 8 GPUs (one call): parity, NVLink counters, bench, the reference's runners on the b200 backend next to NCCL,
# cfg5 (capture on 8 GPUs + the reference's comm_replay / et_replay).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 200 "$@" > $O/r02m_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/r02m_$name.log; }
run dist_check $TR --master-port 29701 tools/dist_check.py
run nvlink_probe $TR --master-port 29702 tools/nvlink_probe.py
timeout 400 $TR --master-port 29703 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02m_bench_n8.log 2> $O/r02m_bench_n8.err
echo "bench rc=$?"
PB200_DLRM_BWD_PARTS=1 timeout 300 $TR --master-port 29704 bench.py --gpus 8 --steps 10 --warmup 3 --skip-e2e > $O/r02m_bench_n8_parts1.log 2> $O/r02m_bench_n8_parts1.err
for be in nccl b200; do
  run comms_$be $TR --master-port 29705 -m -- param_b200.integration.param_plugin comms --backend $be --device cuda \
      --collective all_to_all_single --b 1K --e 1G --f 8 --z 1 --c 1 --n 20 --w 5
done
DLRM="--device cuda --mini-batch-size 4096 --arch-embedding-size $(python -c "print('-'.join(['500000']*128))") --arch-sparse-feature-size 128 --num-indices-per-lookup 20 --num-indices-per-lookup-fixed True --num-batches 8 --warmup-batches 2"
PB200_PLUGIN_BACKEND=stock run dlrm_stock $TR --master-port 29706 -m -- param_b200.integration.param_plugin dlrm --backend nccl $DLRM
run dlrm_b200 $TR --master-port 29707 -m -- param_b200.integration.param_plugin dlrm --backend nccl $DLRM
python tools/make_basic_trace.py --world 8 --out $O/dlrm_step_basic_w8.json > /dev/null
run trace_replay_b200 $TR --master-port 29712 -m -- param_b200.integration.param_plugin trace_replay \
      --trace-path $O/dlrm_step_basic_w8.json --trace-type basic --backend b200 --device cuda --num-replays 3
run cfg5_capture $TR --master-port 29708 tools/cfg5_capture.py --out $O/cfg5_trace_n8 --tables-per-rank 8 --rows 500000 --dim 128 --local-batch 4096 --bag 20
for be in nccl b200; do
  run cfg5_comm_replay_$be $TR --master-port 29709 -m -- param_b200.integration.param_plugin comm_replay --trace-type et \
      --trace-path $O/cfg5_trace_n8 --backend $be --num-replays 5
done
run cfg5_et_replay_stock $TR --master-port 29710 -m -- param_b200.integration.param_plugin et_replay --trace-path $O/cfg5_trace_n8 \
      -m full --warmup-iter 2 --iter 5 --backend nccl --replay-config param_b200/et/replay-config-stock.json
run cfg5_et_replay_b200 $TR --master-port 29711 -m -- param_b200.integration.param_plugin et_replay --trace-path $O/cfg5_trace_n8 \
      -m full --warmup-iter 2 --iter 5 --backend b200 --replay-config param_b200/et/replay-config-b200-aten.json
rm -rf $O/cfg5_trace_n8/*_resources
for r in 2 3 4 5 6 7; do rm -f $O/cfg5_trace_n8/rank-$r.json; done
for f in $O/r02m_*.log; do echo "== $f"; tail -n 8 $f | cut -c1-420; done
