#!/usr/bin/env python3
"""DLRM step (bench_dist.py's cfg4 shape) with the backward as ONE exchange + ONE reduce against the backward in
pieces (partial transpose exchanges on a side stream, each reduce starting when its part has landed), for several
(parts, CTA cap of the partial pushes) settings in one process.  CUDA events, max over ranks.

    torchrun --nproc-per-node 4 tools/pieces_bench.py
"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import ev_time  # noqa: E402
from param_b200.comms.pt.dlrm import DLRMParallelEmbedding, SparseBatch  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
T_l, b, D, L, rows = 64, 8192, 128, 20, 1_000_000
os.environ["PB200_DLRM_BWD_PARTS"] = "8"          # create the comm stream; the sweep sets bwd_parts itself
model = DLRMParallelEmbedding(dist.group.WORLD, [rows] * (T_l * world), D, b, L, dev, lr=1e-6, seed=3)
batch = SparseBatch.synthetic([rows] * (T_l * world), b, L, True, seed=17 + rank, device=dev, alpha=1.15)
state = {}


def step():
    o, i = model.sparse_data_dist(batch)
    state["out"] = model.forward(o, i)
    model.backward(state["out"])


def bwd_only():
    model.backward(state["out"])


def maxr(v):
    t = torch.tensor([float(v)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


res = []
sweep = os.environ.get("PB200_PIECES_SWEEP", "1:0,2:0,2:32,2:64,4:32,4:16,8:32,1:0")
for parts, cap in (tuple(int(v) for v in item.split(":")) for item in sweep.split(",")):
    model.bwd_parts = parts
    os.environ["PB200_A2A_PART_CTAS"] = str(cap)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    ms = maxr(ev_time(step, 10))
    dist.barrier()
    ms_b = maxr(ev_time(bwd_only, 10))
    dist.barrier()
    res.append({"parts": parts, "part_ctas": cap, "step_ms": round(ms, 3), "backward_ms": round(ms_b, 3)})
    if rank == 0:
        print(json.dumps(res[-1]), flush=True)
if model.window.error():
    raise SystemExit("peer wait timed out")
dist.barrier()
dist.destroy_process_group()
