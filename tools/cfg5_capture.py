#!/usr/bin/env python3
"""BASELINE.json configs[4]: capture a Chakra execution trace (ET) of one DLRM step on the box — stock PyTorch
modules only, the way the reference's own model runs it — so that the reference's et_replay can replay it on
the B200 kernels (aten override + b200 comm backend).

Recipe: docs/using_ET.md:10-68 of the reference (ExecutionTraceObserver, the `## process_group:init ##` record,
one rank-{r}.json per rank) and SURVEY.md appendix D (integral tensor DATA must be saved for the embedding
ops, into `<trace stem>_resources/`, or replay draws random offsets).

One step, per rank (table-parallel embeddings, batch-parallel dense part; train/comms/pt/dlrm.py:1200-1323,
train/workloads/dlrm/dlrm_s_pytorch.py:295-327):
    indices all_to_all_single (int64)  ->  T_l x nn.EmbeddingBag(sparse=True) over the GLOBAL batch
    -> cat -> all_to_all_single (fp32, table-parallel -> batch-parallel)  -> interaction (bmm + tril gather)
    -> top linear -> loss;  backward through all of it (transpose all-to-all, EmbeddingBag backward).

    torchrun --nproc-per-node 2 tools/cfg5_capture.py --out gpurun_out/cfg5_trace --tables-per-rank 8 --rows 100000 \
        --dim 128 --local-batch 2048 --bag 20
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.nn as nn

ET_OPS = ("aten::embedding_bag,aten::_embedding_bag,aten::_embedding_bag_forward_only,aten::_embedding_bag_backward,"
          "aten::_embedding_bag_sparse_backward,aten::_embedding_bag_dense_backward,aten::embedding_sparse_backward")


class _A2A(torch.autograd.Function):
    """dist.all_to_all_single with the transposed exchange as its backward (All2Allv_Req/Wait,
    train/comms/pt/dlrm.py:86-218, condensed: synchronous, one Function)."""

    @staticmethod
    def forward(ctx, inp, out_splits, in_splits, group):
        ctx.splits, ctx.group = (out_splits, in_splits), group
        out = inp.new_empty(sum(out_splits) if out_splits else inp.numel())
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=group)
        return out

    @staticmethod
    def backward(ctx, grad):
        out_splits, in_splits = ctx.splits
        gin = grad.new_empty(sum(in_splits) if in_splits else grad.numel())
        dist.all_to_all_single(gin, grad.contiguous(), in_splits, out_splits, group=ctx.group)
        return gin, None, None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--tables-per-rank", type=int, default=8)
    ap.add_argument("--rows", type=int, default=100_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--local-batch", type=int, default=2048)
    ap.add_argument("--bag", type=int, default=20)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--steps-before", type=int, default=2)
    a = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    out_dir = Path(a.out)
    out_dir.mkdir(parents=True, exist_ok=True)
    stem = out_dir / f"rank-{rank}"
    (out_dir / f"rank-{rank}_resources").mkdir(exist_ok=True)        # must exist BEFORE the observer starts
    os.environ["ENABLE_PYTORCH_EXECUTION_TRACE_SAVE_INTEGRAL_TENSOR_DATA"] = ET_OPS
    if a.device == "cuda":
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
    group = dist.group.WORLD
    torch.manual_seed(1234 + rank)
    T_l, E, b, L = a.tables_per_rank, a.dim, a.local_batch, a.bag
    T_g, N = T_l * world, b * world
    # dense gradients: et_replay cannot allocate the sparse COO tensors a sparse=True table hands to autograd
    # ("allocate tensor failed", et_replay.py:819); forward and the gradient scatter are the same ops either way
    embs = nn.ModuleList([nn.EmbeddingBag(a.rows, E, mode="sum", sparse=False) for _ in range(T_l)]).to(dev)
    n_pairs = (T_g + 1) * T_g // 2
    top = nn.Linear(E + n_pairs, 1).to(dev)
    dense = torch.randn(b, E, device=dev)
    li, lj = torch.tril_indices(T_g + 1, T_g + 1, offset=-1)
    # the pairs as flat positions of the [T_g + 1, T_g + 1] interaction matrix: index_select instead of Z[:, li, lj],
    # whose aten::index carries a None in its Tensor?[] argument that et_replay cannot rebuild (et_replay.py:1240)
    tril_flat = (li * (T_g + 1) + lj).to(dev)
    # this rank's sparse inputs: LOCAL batch, ALL tables (table-major), fixed bag size
    my_idx = torch.randint(0, a.rows, (T_g * b * L,), device=dev)
    offsets = torch.arange(0, N * L, L, device=dev)                   # per table, global batch
    # explicit unit per-sample weights: with None the backward node records an UNDEFINED tensor argument
    # ("Tensor(nullptr (uninitialized))"), which et_replay cannot allocate (et_replay.py:819) and then cannot run
    psw = torch.ones(N * L, device=dev)
    opt = torch.optim.SGD(list(embs.parameters()) + list(top.parameters()), lr=0.01)

    def step():
        # sparse input redistribution: every rank receives the indices of ITS tables from everyone
        recv = torch.empty(world * T_l * b * L, dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv, my_idx, group=group)
        per_table = recv.view(world, T_l, b * L).permute(1, 0, 2).contiguous()      # [T_l, W * b * L]
        ly = [embs[t](per_table[t].view(-1), offsets, per_sample_weights=psw) for t in range(T_l)]   # T_l x [N, E]
        pooled = torch.cat(ly, dim=1)                                                  # [N, T_l * E]
        out = _A2A.apply(pooled.view(-1), [b * T_l * E] * world, [b * T_l * E] * world, group)
        x = torch.cat([o.view(b, T_l * E) for o in out.split(b * T_l * E)], dim=1)   # [b, T_g * E]
        z = torch.cat([dense.view(b, 1, E), x.view(b, T_g, E)], dim=1)                # [b, T_g + 1, E]
        zz = torch.index_select(torch.bmm(z, z.transpose(1, 2)).view(b, -1), 1, tril_flat)   # interaction
        loss = top(torch.cat([dense, zz], dim=1)).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(a.steps_before):
        step()
    if dev.type == "cuda":
        torch.cuda.synchronize()
    from torch.profiler import ExecutionTraceObserver
    et = ExecutionTraceObserver()
    et.register_callback(str(stem) + ".json")
    et.start()
    h = torch.autograd._record_function_with_args_enter(
        "## process_group:init ##", json.dumps(dist.distributed_c10d._world.pg_config_info))
    torch.autograd._record_function_with_args_exit(h)
    loss = step()
    if dev.type == "cuda":
        torch.cuda.synchronize()
    et.stop()
    et.unregister_callback()
    dist.barrier()
    n_res = len(list((out_dir / f"rank-{rank}_resources").iterdir()))
    print(f"[rank {rank}] trace {stem}.json ({os.path.getsize(str(stem) + '.json')} B), {n_res} resource files, "
          f"loss {float(loss):.6f}; tables/rank {T_l} x {a.rows} x {E}, local batch {b}, bag {L}")
    dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
