"""bench.py, N > 1: BASELINE.json configs[3] — the DLRM table-parallel step on N B200s of one box.

Per rank: 64 tables (dim 128, rows scaled to fit), local batch 8192, bag 20; each rank looks up ITS
tables for the GLOBAL batch, the pooled vectors are exchanged by the fused peer-push all-to-all
(table-parallel -> batch-parallel, output permute included), and the backward is the transpose
exchange followed by the scatter-add into the local tables.  No collective other than that exchange
is on the data path; NCCL is used for bootstrap / barriers and as the comparator only.

value = lookups processed by ALL ranks per second over the step (lookup fwd + a2a fwd + a2a bwd +
scatter-add bwd), timed with CUDA events, max over ranks.  `a2a` reports the exchange alone as bus
bandwidth (PARAM's definition: bytes of the rank's OUTPUT tensor / time * (W-1)/W) next to NCCL's
all_to_all_single on the same tensors, and `sweep` a few all_to_all_single sizes (config 2).
"""
from __future__ import annotations

import json
import os

import torch
import torch.distributed as dist

from bench import METRIC, UNIT, ClockSampler, algorithmic_bytes, ev_time, measured_peaks

NVLINK_PEAK_MEASURED = 770.0   # GB/s per direction per GPU, peer copy (B200_PROFILING.md)
NVLINK_PEAK_NOMINAL = 900.0


def run_dist(args):
    from param_b200 import _cabi, ops
    from param_b200.comms.pt.dlrm import DLRMParallelEmbedding, SparseBatch

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    T_l = int(os.environ.get("PB200_TABLES_PER_GPU", 64))
    b = int(os.environ.get("PB200_LOCAL_BATCH", 8192))
    D, L = args.dim, args.bag
    rows = int(os.environ.get("PB200_ROWS", min(args.rows, 2_000_000)))
    T_g, N = T_l * world, b * world
    model = DLRMParallelEmbedding(group, [rows] * T_g, D, b, L, dev, lr=args.lr, seed=3,
                                  bwd_algo="sorted" if args.bwd_algo == "auto" else args.bwd_algo)
    win = model.window
    batch = SparseBatch.synthetic([rows] * T_g, b, L, True, seed=17 + rank, device=dev, alpha=args.alpha)
    offsets, indices = model.sparse_data_dist(batch)
    torch.cuda.synchronize()
    lookups_rank = T_l * N * L
    state = {"out": None}

    def fwd():
        state["out"] = model.forward(offsets, indices)

    def bwd():
        model.backward(state["out"])      # the received pooled tensor doubles as dOut

    def step():
        # one iteration of the reference's loop (dlrm.py:1200-1323): sparse input redistribution, lookup +
        # forward exchange, backward exchange + scatter-add.  The backward's sort plan is queued on a side
        # stream by model.forward() and runs under the exchanges.
        o, i = model.sparse_data_dist(batch)
        state["out"] = model.forward(o, i)
        bwd()

    def compute_only():
        # the same per-rank lookups with no exchange at all: what one GPU does alone on this rank's work
        ops.tbe_forward(model.arena, indices, offsets, N, layout="BTD", out=model._pooled_local)
        ops.tbe_backward(model.arena.weights, model.arena.row_offsets, T_l, D, indices, offsets, N,
                         model._pooled_local, scale=-args.lr, algo=model.bwd_algo, max_table_rows=rows)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    n0 = _cabi.launch_count()
    with ClockSampler(local_rank) as clk:
        ms_step = ev_time(step, args.steps)
    launches = _cabi.launch_count() - n0
    torch.cuda.synchronize()
    dist.barrier()

    def maxr(v):
        t = torch.tensor([float(v)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_step = maxr(ms_step)

    # ---- pieces ---------------------------------------------------------------------------------
    pooled = model._pooled_local
    ms_compute_only = maxr(ev_time(compute_only, args.steps))
    dist.barrier()
    ms_lookup = maxr(ev_time(lambda: ops.tbe_forward(model.arena, indices, offsets, N, layout="BTD", out=pooled),
                             args.steps))
    dist.barrier()
    ms_a2a_f = maxr(ev_time(lambda: win.pooled_forward(pooled, model.batch_split, model.tables_split, D,
                                                       out_window_off=model.off_pooled), args.steps))
    dist.barrier()
    ms_fused = maxr(ev_time(lambda: win.lookup_forward_fused(model.arena, indices, offsets, model.batch_split,
                                                             model.tables_split, out_window_off=model.off_pooled),
                            args.steps))
    dist.barrier()
    out = state["out"]
    ms_a2a_b = maxr(ev_time(lambda: win.pooled_backward(out, model.batch_split, model.tables_split, D,
                                                        out_window_off=model.off_grad), args.steps))
    dist.barrier()
    g_local = win.view(model.off_grad, N * T_l * D, torch.float32).view(N, T_l * D)
    ms_scatter = maxr(ev_time(lambda: ops.tbe_backward(model.arena.weights, model.arena.row_offsets, T_l, D,
                                                       indices, offsets, N, g_local, scale=-args.lr,
                                                       algo=model.bwd_algo, max_table_rows=rows), args.steps))
    dist.barrier()
    ms_plan = ms_reduce = None
    if model.bwd_algo in ("sorted", "auto"):
        ms_plan = maxr(ev_time(lambda: ops.tbe_plan(model.arena.row_offsets, T_l, D, indices, offsets, N, rows,
                                                    buf=model._plan_buf), args.steps))
        plan = ops.tbe_plan(model.arena.row_offsets, T_l, D, indices, offsets, N, rows, buf=model._plan_buf)
        ms_reduce = maxr(ev_time(lambda: ops.tbe_backward(model.arena.weights, model.arena.row_offsets, T_l, D,
                                                          indices, offsets, N, g_local, scale=-args.lr,
                                                          algo="sorted", max_table_rows=rows, plan=plan),
                                 args.steps))
        dist.barrier()
    ms_dist = maxr(ev_time(lambda: model.sparse_data_dist(batch), max(2, args.steps // 2)))
    dist.barrier()

    # ---- NCCL comparator: the reference's exchange on the same tensors (cat + a2a + split/cat) ------
    S = b * T_g * D * 4                                    # bytes of the rank's output tensor
    in_splits = [m * T_l * D for m in model.batch_split]
    out_splits = [b * t * D for t in model.tables_split]
    ly = pooled.view(N, T_l, D).permute(1, 0, 2).contiguous()   # [T_l, N, E], what apply_emb returns
    nccl_out = torch.empty(sum(out_splits), device=dev)

    def nccl_ref_fwd():
        inp = torch.cat(list(ly), dim=1).view(-1)
        dist.all_to_all_single(nccl_out, inp, out_splits, in_splits)
        return torch.cat([o.view(b, -1) for o in nccl_out.split(out_splits)], dim=1)

    def nccl_raw():
        dist.all_to_all_single(nccl_out, pooled.view(-1), out_splits, in_splits)

    for _ in range(3):
        nccl_ref_fwd()
    ms_nccl_ref = maxr(ev_time(nccl_ref_fwd, args.steps))
    ms_nccl_raw = maxr(ev_time(nccl_raw, args.steps))
    same = torch.equal(nccl_ref_fwd(), win.pooled_forward(pooled, model.batch_split, model.tables_split, D,
                                                          out_window_off=model.off_pooled))
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    del ly

    def bus(ms):
        return S / (ms * 1e-3) / 1e9 * (world - 1) / world

    # ---- all_to_all_single sweep (config 2), b200 vs nccl ------------------------------------------
    sweep = []
    for size in (1 << 10, 64 << 10, 1 << 20, 16 << 20, 256 << 20, 1 << 30):
        numel = size // 4 // world * world
        if numel == 0 or size > win.window_bytes // 2:
            continue
        xin = torch.arange(numel, device=dev, dtype=torch.float32) + rank
        xo_w = win.view(0, numel, torch.float32)
        xo_n = torch.empty(numel, device=dev)
        for _ in range(3):
            win.all_to_all_single(xo_w, xin)
            dist.all_to_all_single(xo_n, xin)
        dist.barrier()
        t_b = maxr(ev_time(lambda: win.all_to_all_single(xo_w, xin), 10))
        dist.barrier()
        t_n = maxr(ev_time(lambda: dist.all_to_all_single(xo_n, xin), 10))
        ok = torch.equal(xo_w, xo_n)
        f = (world - 1) / world
        sweep.append({"bytes": numel * 4, "b200_us": t_b * 1e3, "nccl_us": t_n * 1e3,
                      "b200_busbw": numel * 4 / t_b / 1e6 * f, "nccl_busbw": numel * 4 / t_n / 1e6 * f,
                      "bit_exact_vs_nccl": bool(ok)})
    err = win.error()

    peak, peak_src = measured_peaks()
    fwd_bytes, bwd_bytes = algorithmic_bytes(T_l, N, L, D)
    dom, dom_ms, dom_bytes = ("fwd", ms_lookup, fwd_bytes) if ms_lookup >= ms_scatter else ("bwd", ms_scatter, bwd_bytes)
    ach = dom_bytes / (dom_ms * 1e-3) / 1e9
    res = {
        "metric": METRIC, "value": world * lookups_rank / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cfg4: DLRM table-parallel step, {T_l} tables/GPU x {rows} rows x {D} dim, local batch {b} "
                               f"(global {N}), bag {L}, Zipf alpha={args.alpha}; sparse input redistribution -> lookup fwd "
                               "fused with the peer-push all-to-all (+output permute) -> transpose all-to-all -> "
                               "scatter-add bwd (sort plan on a side stream)",
                   "tables_per_gpu": T_l, "rows_per_table": rows, "dim": D, "local_batch": b, "bag": L,
                   "alpha": args.alpha, "parallelism": f"table-parallel x{world}", "peer_mapping": getattr(win, "mapping", "?"),
                   "l2_policy": "inputs larger than L2 (arena %.1f GB/GPU, exchange %.2f GB/GPU/direction)" %
                                (model.arena.weights.numel() * 4 / 1e9, S / 1e9)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "lookup_" + dom, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(ach / peak, 4), "peak_source": peak_src, "traffic": None,
                     "algorithmic_bytes": dom_bytes, "ms": round(dom_ms, 4)},
        "pieces_ms": {"sparse_input_dist": ms_dist, "fused_lookup_a2a_fwd_one_kernel": ms_fused,
                      "lookup_fwd": ms_lookup, "a2a_fwd_fused_permute": ms_a2a_f,
                      "a2a_bwd_fused_permute": ms_a2a_b, "scatter_add_bwd_inline_sort": ms_scatter,
                      "bwd_sort_plan": ms_plan, "bwd_segment_reduce": ms_reduce,
                      "note": "the step contains sparse_input_dist; the sort plan is queued on a side stream at "
                              "forward time and overlaps the exchanges"},
        # like-for-like scaling: the same per-rank lookups (this rank's tables over the global batch, forward +
        # backward) with no exchange, on one GPU.  value / (N * n1_equivalent_value) is the share of the step
        # that is not exposed communication; it does not mix workloads the way value(N) / value(1, cfg2) does.
        "compute_only_ms": ms_compute_only,
        "n1_equivalent_value": lookups_rank / (ms_compute_only * 1e-3),
        "efficiency_like_for_like": ms_compute_only / ms_step,
        "a2a": {"bytes_per_rank": S, "fwd_busbw_gbs": bus(ms_a2a_f), "bwd_busbw_gbs": bus(ms_a2a_b),
                "frac_of_measured_peer_copy_770": bus(ms_a2a_f) / NVLINK_PEAK_MEASURED,
                "frac_of_nominal_900": bus(ms_a2a_f) / NVLINK_PEAK_NOMINAL,
                "nccl_a2a_only_busbw_gbs": bus(ms_nccl_raw), "nccl_reference_path_ms": ms_nccl_ref,
                "nccl_reference_path_busbw_gbs": bus(ms_nccl_ref),
                "fused_matches_nccl_reference_bit_exact": bool(int(flag)), "timeout_errors": err},
        "sweep_all_to_all_single": sweep,
        "clocks": clk.summary(),
    }
    if not args.skip_e2e:
        try:
            res["e2e"] = e2e_dist(args, model, batch, lookups_rank, world, maxr)
        except Exception as exc:  # noqa: BLE001 — the device-timed line must survive a failing e2e leg
            res["e2e"] = {"value": None, "unit": UNIT, "error": repr(exc)[:300]}
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


def e2e_dist(args, model, batch, lookups_rank, world, maxr):
    """Through the public API from HOST buffers: this rank's sparse inputs start in pinned host memory; what ends
    in pinned host memory is the step's scalar result — the sum of this rank's batch-parallel pooled tensor
    (pb200_pooled_sum per sample row, then one add over the rows), as the reference's step keeps the pooled
    tensor on the device for the layers above (dlrm.py:1239-1257).  The form that copies the whole pooled tensor
    back is timed beside it (`full_output`)."""
    import time
    from param_b200 import ops
    from param_b200.comms.pt.dlrm import SparseBatch
    h_len = batch.lengths.cpu().pin_memory()
    h_idx = batch.indices.cpu().pin_memory()
    d_len, d_idx = torch.empty_like(batch.lengths), torch.empty_like(batch.indices)
    h_out = torch.empty((model.b, model.T_global * model.E), dtype=torch.float32).pin_memory()
    h_loss = torch.zeros(1, dtype=torch.float64).pin_memory()

    def call(full):
        d_len.copy_(h_len, non_blocking=True)
        d_idx.copy_(h_idx, non_blocking=True)
        sb = SparseBatch(batch.count, batch.batch_size, d_len, d_idx)
        offsets, indices = model.sparse_data_dist(sb)
        out = model.forward(offsets, indices)
        if full:
            h_out.copy_(out, non_blocking=True)
        else:
            h_loss.copy_(ops.pooled_sum(out, out.shape[0]).sum(dim=0, keepdim=True), non_blocking=True)
        model.backward(out)

    def timed(full):
        for _ in range(2):
            call(full)
        torch.cuda.synchronize()
        dist.barrier()
        n = max(3, args.steps // 2)
        t0 = time.perf_counter()
        for _ in range(n):
            call(full)
        torch.cuda.synchronize()
        return maxr((time.perf_counter() - t0) / n)

    dt = timed(False)
    dt_full = timed(True)
    h2d = int((h_len.numel() + h_idx.numel()) * 8)
    return {"value": world * lookups_rank / dt, "unit": UNIT, "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
            "result": "loss = sum of the rank's pooled [lN, T_g*E] tensor (float64)",
            "full_output": {"value": world * lookups_rank / dt_full, "ms_per_step": dt_full * 1e3,
                            "d2h_bytes_per_step": int(h_out.numel() * 4)},
            "path": "DLRMParallelEmbedding: H2D lengths+indices -> sparse_data_dist (2 peer-push a2a + regroup) -> "
                    "fused lookup + a2a -> pooled sum -> D2H loss -> transpose a2a -> scatter-add"}
