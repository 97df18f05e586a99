#!/bin/bash
# Round 2, 2 GPUs, third pass: backward in pieces (parity + bench), cfg5 full replay with the plain configs (the
# --update-replay-config discovery pass of r02k wrongly put every op on the skip list), basic-trace replay with
# emb_lookup entries through the reference's commsTraceReplay.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 240 "$@" > $O/r02l_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/r02l_$name.log; }
timeout 200 python -m pytest tests/test_gpu_a2a_local.py tests/test_gpu_sort_plan.py -q -x --timeout 120 -p no:cacheprovider -k "pieces or table_groups" > $O/r02l_tests_pieces.log 2>&1
echo "pieces tests rc=$?" | tee -a $O/r02l_tests_pieces.log
run dist_check $TR --master-port 29701 tools/dist_check.py
run trace_replay_b200 $TR --master-port 29712 -m -- param_b200.integration.param_plugin trace_replay \
      --trace-path param_b200/comms/pt/traces/dlrm_step_basic.json --trace-type basic --backend b200 --device cuda --num-replays 3
run trace_replay_nccl_comms_only $TR --master-port 29713 -m -- param_b200.integration.param_plugin trace_replay \
      --trace-path param_b200/comms/pt/traces/dlrm_step_basic.json --trace-type basic --backend nccl --device cuda --num-replays 3
run cfg5_capture $TR --master-port 29707 tools/cfg5_capture.py --out $O/cfg5_trace --tables-per-rank 8 --rows 200000 --dim 128 --local-batch 2048 --bag 20
run cfg5_et_replay_stock $TR --master-port 29709 -m -- param_b200.integration.param_plugin et_replay --trace-path $O/cfg5_trace \
      -m full --warmup-iter 2 --iter 5 --backend nccl --replay-config param_b200/et/replay-config-stock.json
run cfg5_et_replay_b200 $TR --master-port 29710 -m -- param_b200.integration.param_plugin et_replay --trace-path $O/cfg5_trace \
      -m full --warmup-iter 2 --iter 5 --backend b200 --replay-config param_b200/et/replay-config-b200-aten.json
run cfg5_et_replay_b200_comp python -m param_b200.integration.param_plugin et_replay --input $O/cfg5_trace/rank-0.json \
      -m comp --warmup-iter 2 --iter 5 --replay-config param_b200/et/replay-config-b200-aten.json
run cfg5_et_replay_stock_comp python -m param_b200.integration.param_plugin et_replay --input $O/cfg5_trace/rank-0.json \
      -m comp --warmup-iter 2 --iter 5 --replay-config param_b200/et/replay-config-stock.json
timeout 400 $TR --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02l_bench_n2.log 2> $O/r02l_bench_n2.err
echo "bench rc=$?"
PB200_DLRM_BWD_PARTS=1 timeout 400 $TR --master-port 29714 bench.py --gpus 2 --steps 10 --warmup 3 --skip-e2e > $O/r02l_bench_n2_parts1.log 2> $O/r02l_bench_n2_parts1.err
PB200_DLRM_BWD_PARTS=4 timeout 400 $TR --master-port 29715 bench.py --gpus 2 --steps 10 --warmup 3 --skip-e2e > $O/r02l_bench_n2_parts4.log 2> $O/r02l_bench_n2_parts4.err
rm -rf $O/cfg5_trace/*_resources
for f in $O/r02l_*.log; do echo "== $f"; tail -n 6 $f | cut -c1-400; done
