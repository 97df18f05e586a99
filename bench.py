#!/usr/bin/env python3
"""bench.py — PARAM EmbeddingBag + DLRM all-to-all hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N = 1 (default): BASELINE.json configs[1] — 256-table batched EmbeddingBag fwd+bwd, dim 128, global
batch 65536, bag 20, Zipf alpha=1.15 — with rows/table scaled to fit one GPU (256 x 10M x 128 fp32 is
1.31 TB).  A step = one forward over all tables + one backward (scatter-add of the
pooled gradient into the table arena, fused SGD).  Prints ONE JSON line.
N > 1: see bench_dist (DLRM table-parallel step: lookup -> fused all-to-all -> transpose -> scatter-add).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

METRIC = "embeddingbag_lookups_per_sec"
UNIT = "lookups/s"


# ------------------------------------------------------------------------------------------------
def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--tables", type=int, default=256)
    ap.add_argument("--rows", type=int, default=1_000_000, help="rows per table (scaled to fit HBM)")
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--bag", type=int, default=20)
    ap.add_argument("--alpha", type=float, default=1.15)
    ap.add_argument("--fwd-algo", default="auto")
    ap.add_argument("--bwd-algo", default="auto")
    ap.add_argument("--lr", type=float, default=1e-6)
    ap.add_argument("--cpu-tables", type=int, default=4, help="tables in the CPU baseline sample")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--e2e-group", type=int, default=16, help="tables per H2D/kernel/D2H pipeline group")
    ap.add_argument("--presort", choices=["side", "inline"], default="inline",
                    help="where the backward's sort plan is built: after the forward on the same stream, or "
                         "on a side stream queued before the forward (inside the timed step either way)")
    ap.add_argument("--skip-uniform", action="store_true", help="skip the alpha=0 roofline leg")
    ap.add_argument("--quick", action="store_true", help="skip the alternative kernel variants")
    return ap.parse_args(argv)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(T, B, L, D):
    """fwd = B*L*(D*4+8) + B*8 + B*D*4 per table; bwd = B*L*(2*D*4+8) + B*D*4."""
    fwd = T * (B * L * (D * 4 + 8) + B * 8 + B * D * 4)
    bwd = T * (B * L * (2 * D * 4 + 8) + B * D * 4)
    return fwd, bwd


def ev_time(fn, iters, stream=None):
    """CUDA-event time of `iters` back-to-back calls of fn on the current stream -> ms per call."""
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()                       # one untimed call: first-use allocations, lazy module loads
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(iters):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ------------------------------------------------------------------------------------------------
def build_workload(args, dev):
    from param_b200 import ops
    T, B, L, D = args.tables, args.batch, args.bag, args.dim
    free, total = torch.cuda.mem_get_info(dev)
    # idx, off, out, a gradient copy for the parity check, the sort plan (24 B per lookup + histograms), slack
    fixed = T * B * L * 8 + (T * B + 1) * 8 + 2 * T * B * D * 4 + T * B * L * 28 + (6 << 30)
    rows = args.rows
    max_rows = int((free - fixed) // (T * D * 4))
    scaled = False
    if rows > max_rows:
        rows, scaled = max(max_rows // 1000 * 1000, 1000), True
    arena = ops.TableArena.allocate([rows] * T, D, dev)
    ops.fill_uniform_(arena.weights, -(1.0 / rows) ** 0.5, (1.0 / rows) ** 0.5, seed=2026)
    idx = torch.empty(T * B * L, dtype=torch.int64, device=dev)
    fill_indices(args, idx, rows, args.alpha, dev)
    off = torch.arange(T * B + 1, dtype=torch.int64, device=dev) * L
    torch.cuda.synchronize()
    return arena, idx, off, rows, scaled


def fill_indices(args, idx, rows, alpha, dev):
    """(re)generate the request in place: Zipf(alpha) by inverse CDF, per-bag distinct like the reference's
    init_indices (pytorch_emb.py:138-160), or uniform for alpha == 0"""
    from param_b200 import ops
    from param_b200.compute.pt.pytorch_emb import zipf_cdf
    T, B, L = args.tables, args.batch, args.bag
    if alpha > 0:
        cdf = torch.from_numpy(zipf_cdf(alpha, rows)).to(dev)
    else:
        cdf = torch.linspace(1.0 / rows, 1.0, rows, dtype=torch.float64, device=dev)
    for t in range(T):
        ops.fill_zipf_indices_(idx[t * B * L:(t + 1) * B * L], L, cdf, seed=1000 + t, dedupe=alpha > 0)
    torch.cuda.synchronize()


def sampled_parity(args, arena, idx, off, out, rows, n_samples=64):
    """CHECKER (the one place besides cpu_baseline where bench.py touches oracle/): after the timed region,
    one more forward + backward of the SAME kernels on the same request; 64 sampled bags of the pooled
    output are compared bit for bit with the CPU oracle (fp32 sum in index order over the rows as they are
    in HBM), and the update of 64 sampled rows with a float64 sum over ALL their lookups (1e-5 relative)."""
    import numpy as np
    from oracle import oracle
    from param_b200 import ops
    T, B, L, D = args.tables, args.batch, args.bag, args.dim
    rng = np.random.default_rng(12345)
    ts, bs = rng.integers(0, T, n_samples), rng.integers(0, B, n_samples)
    ops.tbe_forward(arena, idx, off, B, layout="BTD", algo=args.fwd_algo, out=out)
    torch.cuda.synchronize()
    ok_fwd = True
    for t, b in zip(ts.tolist(), bs.tolist()):
        ids = idx[(t * B + b) * L:(t * B + b + 1) * L]
        w = arena.weights[t * rows + ids].cpu().numpy()
        want = oracle.embbag_fwd(w, np.arange(L, dtype=np.int64), np.zeros(1, np.int64))[0]
        got = out[b, t * D:(t + 1) * D].cpu().numpy()
        ok_fwd &= bool(np.array_equal(got, want))
    # backward: the rows of the first lookup of each sampled bag; every lookup of such a row is found by
    # scanning its table's indices.  The check step uses lr = 1 so that the update is not below the fp32
    # resolution of the weights (it runs after the timed region).
    lr = 1.0
    sample_rows = sorted({(int(t), int(idx[(int(t) * B + int(b)) * L].item())) for t, b in zip(ts, bs)})
    before = {tr: arena.weights[tr[0] * rows + tr[1]].double().cpu().numpy() for tr in sample_rows}
    grad = out.clone()
    ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, grad, layout="BTD", scale=-lr,
                     algo="sorted", max_table_rows=rows)
    torch.cuda.synchronize()
    worst = 0.0
    for (t, r) in sample_rows:
        pos = (idx[t * B * L:(t + 1) * B * L] == r).nonzero().view(-1)
        bags = torch.div(pos, L, rounding_mode="floor")
        g = grad[bags, t * D:(t + 1) * D].double().sum(dim=0).cpu().numpy()
        want = before[(t, r)] - lr * g
        got = arena.weights[t * rows + r].double().cpu().numpy()
        worst = max(worst, float(np.abs(got - want).max()) / max(float(np.abs(want).max()), 1e-30))
    del grad
    return {"parity_sampled": bool(ok_fwd and worst <= 1e-5), "fwd_bags_bit_exact": bool(ok_fwd),
            "bwd_rows_max_rel_err": worst, "bags": int(n_samples), "rows": len(sample_rows),
            "checker": "oracle/param_oracle.c embbag_fwd (bit-exact) + float64 sum over all lookups of the row (1e-5)"}


def run_b200(args):
    from param_b200 import _cabi, ops
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    T, B, L, D = args.tables, args.batch, args.bag, args.dim
    arena, idx, off, rows, scaled = build_workload(args, dev)
    out = torch.empty((B, T * D), dtype=torch.float32, device=dev)
    lookups = T * B * L
    bwd_algo = "sorted" if args.bwd_algo == "auto" else args.bwd_algo
    side = torch.cuda.Stream(device=dev)
    plans = {}

    def fwd():
        ops.tbe_forward(arena, idx, off, B, layout="BTD", algo=args.fwd_algo, out=out)

    def plan(stream=None, exact=False):
        # the index-only half of the backward: every table's lookups sorted by row (hand-written radix sort)
        key = "exact" if exact else "sorted"
        p = ops.tbe_plan(arena.row_offsets, T, D, idx, off, B, rows, layout="BTD", exact=exact, stream=stream,
                         buf=plans.get(key))
        plans[key] = p.buf
        return p

    def reduce(p):
        # the pooled output doubles as the incoming gradient (same shape; saves 8.6 GB of HBM)
        ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, out, layout="BTD",
                         scale=-args.lr, algo=bwd_algo, max_table_rows=rows, plan=p)

    def bwd():
        if bwd_algo == "atomic":
            ops.tbe_backward(arena.weights, arena.row_offsets, T, D, idx, off, B, out, layout="BTD",
                             scale=-args.lr, algo="atomic")
        else:
            reduce(plan(exact=bwd_algo == "exact"))

    def step():
        # one training step of the path: forward over all tables, then the backward (sort plan + segmented
        # reduce, fused SGD into the arena).  --presort side queues the sort on a side stream before the
        # forward (it needs the indices only); the timed region contains it either way.
        if args.presort == "side" and bwd_algo != "atomic":
            p = plan(stream=side, exact=bwd_algo == "exact")
            fwd()
            reduce(p)
        else:
            fwd()
            bwd()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    n0 = _cabi.launch_count()
    with ClockSampler(local_rank) as clk:
        ms_step = ev_time(step, args.steps)
    launches = _cabi.launch_count() - n0
    if world > 1:
        import torch.distributed as dist
        tmax = torch.tensor([ms_step], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_step = float(tmax.item())
        dist.barrier()
    parity = sampled_parity(args, arena, idx, off, out, rows) if rank == 0 else None

    peak, peak_src = measured_peaks()
    fwd_bytes, bwd_bytes = algorithmic_bytes(T, B, L, D)
    traffic = {}
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text())
        except Exception:
            traffic = {}

    def roof(nbytes, ms, dram=None):
        a = nbytes / (ms * 1e-3) / 1e9
        r = {"bound": "hbm", "achieved": round(a, 1), "peak": peak, "unit": "GB/s", "frac": round(a / peak, 4),
             "peak_source": peak_src, "algorithmic_bytes": int(nbytes), "ms": round(ms, 4), "traffic": dram,
             "traffic_source": None if dram is None else
             "static: ncu dram__bytes_read.sum + dram__bytes_write.sum of the same launch at this size "
             "(profiles/traffic.json), not measured in this run"}
        if dram is not None:
            r["dram_gbs"] = round(dram / (ms * 1e-3) / 1e9, 1)
            r["frac_dram"] = round(dram / (ms * 1e-3) / 1e9 / peak, 4)
        return r

    def distinct_rows():
        """Measured in this run, outside every timed region: how many distinct (table, row) pairs the request touches
        (torch.unique per table) — what fixes the COMPULSORY DRAM traffic of the request: every distinct row must be
        read once by the forward, read and written once by the reduce, whatever the caches do."""
        try:
            n = 0
            for t in range(T):
                n += int(torch.unique(idx[t * B * L:(t + 1) * B * L]).numel())
            return n
        except Exception:  # noqa: BLE001 — an extra figure must not take the line down
            return None

    def with_compulsory(r, ms, nbytes):
        if nbytes is not None:
            r["compulsory_bytes"] = int(nbytes)
            r["frac_dram_compulsory"] = round(nbytes / (ms * 1e-3) / 1e9 / peak, 4)
        return r

    def legs(tag):
        """CUDA-event time of every kernel (group) of the step, same process, same inputs"""
        k = {}
        nd = distinct_rows()
        comp_fwd = comp_red = None
        if nd is not None:
            comp_fwd = nd * D * 4 + lookups * 8 + (T * B + 1) * 8 + T * B * D * 4     # rows + indices + offsets + output
            comp_red = T * B * D * 4 + 2 * nd * D * 4 + lookups * 8                   # gradient + row RMW + keys / values
        k["fwd"] = ev_time(fwd, args.steps)
        if bwd_algo != "atomic":
            k["sort_plan"] = ev_time(lambda: plan(exact=bwd_algo == "exact"), args.steps)
            p = plan(exact=bwd_algo == "exact")
            k["reduce"] = ev_time(lambda: reduce(p), args.steps)
        k["bwd"] = ev_time(bwd, args.steps)
        sfx = "" if tag == "zipf" else "_uniform"
        r = {"fwd": dict(with_compulsory(roof(fwd_bytes, k["fwd"], traffic.get("fwd" + sfx)), k["fwd"], comp_fwd),
                         kernel="tbe_fwd_direct_kernel_var<.., 16, 2, 5, 4>: two bags per warp (1 launch/step)",
                         lookups_per_s=lookups / k["fwd"] * 1e3,
                         param_bw_gbs=round(lookups * D * 4 / k["fwd"] / 1e6, 1))}
        if "reduce" in k:
            r["bwd_reduce"] = dict(with_compulsory(roof(bwd_bytes, k["reduce"], traffic.get("bwd_reduce" + sfx)),
                                                   k["reduce"], comp_red),
                                   kernel="segment_reduce_kernel (1 launch/step; carries all of the backward's "
                                          "algorithmic bytes: gradient rows in, row read-modify-write)")
            sort_bytes = lookups * 48
            r["bwd_sort_plan"] = dict(roof(sort_bytes, k["sort_plan"], traffic.get("bwd_sort_plan" + sfx)),
                                      kernel="radix_hist/scan/scatter_kernel x passes (6 launches/step at 20 key bits)",
                                      keys_per_s=lookups / k["sort_plan"] * 1e3,
                                      note="48 B per lookup moved by a 2-pass plan (csrc/radix_sort.cu); overhead on "
                                           "top of the backward's algorithmic bytes, which are charged to bwd_reduce")
        r["bwd"] = dict(roof(bwd_bytes, k["bwd"], traffic.get("bwd" + sfx)), kernel="backward as a whole", algo=bwd_algo)
        return k, r

    ms, roofs = legs("zipf" if args.alpha > 0 else "uniform")
    # the single kernel that takes the largest share of the step
    dom = "bwd_reduce" if ("bwd_reduce" in roofs and ms["reduce"] >= ms["fwd"]) else "fwd"
    extra = {}
    if not args.quick:
        extra["fwd_staged_ms"] = round(ev_time(lambda: ops.tbe_forward(arena, idx, off, B, algo="staged", out=out), args.steps), 4)
        try:
            n2 = max(2, args.steps // 2)
            extra["bwd_atomic_ms"] = round(ev_time(lambda: ops.tbe_backward(
                arena.weights, arena.row_offsets, T, D, idx, off, B, out, layout="BTD", scale=-args.lr,
                algo="atomic"), n2), 4)
            extra["bwd_exact_sgd_ms"] = round(ev_time(lambda: ops.tbe_backward_fused(
                arena.weights, arena.row_offsets, T, D, idx, off, B, out, optimizer="exact_sgd", lr=args.lr,
                max_table_rows=rows), n2), 4)
            state = torch.zeros(arena.total_rows, dtype=torch.float32, device=dev)
            extra["bwd_exact_rowwise_adagrad_ms"] = round(ev_time(lambda: ops.tbe_backward_fused(
                arena.weights, arena.row_offsets, T, D, idx, off, B, out, optimizer="exact_row_wise_adagrad",
                lr=args.lr, state=state, max_table_rows=rows), n2), 4)
            del state
        except Exception as exc:  # noqa: BLE001
            extra["variants_failed"] = str(exc)

    res = {
        "metric": METRIC, "value": world * lookups / (ms_step * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"cfg2: {T}-table batched EmbeddingBag fwd+bwd, {rows} rows x {D} dim per table "
                               f"(scaled from 10M rows to fit 180 GB HBM{'; further scaled to free memory' if scaled else ''}), "
                               f"global batch {B}, bag {L}, Zipf alpha={args.alpha}, int64 indices",
                   "tables": T, "rows_per_table": rows, "dim": D, "batch": B, "bag": L, "alpha": args.alpha,
                   "fwd_algo": args.fwd_algo, "bwd_algo": bwd_algo, "presort": args.presort,
                   "l2_policy": "inputs larger than L2 (arena %.1f GB, pooled output %.1f GB)" %
                                (arena.weights.numel() * 4 / 1e9, out.numel() * 4 / 1e9),
                   "parallelism": "replicas" if world > 1 else "single"},
        "gpu_launches": int(launches),
        "roofline": dict(roofs[dom], dominant_kernel=dom,
                         share_of_step=round((ms["reduce"] if dom == "bwd_reduce" else ms["fwd"]) / ms_step, 3),
                         note="Zipf skew: most row reads hit L1/L2, so algorithmic GB/s can exceed the DRAM peak; "
                              "frac_dram is the DRAM-traffic figure and roofline_uniform the HBM-bound case"
                              if args.alpha > 0 else "uniform indices: HBM-bound"),
        "roofline_kernels": roofs,
        "kernels_ms": dict({k: round(v, 4) for k, v in ms.items()}, **extra),
        "parity": parity,
        "clocks": clk.summary(),
    }
    if rank == 0 and not args.skip_e2e:
        res["e2e"], res["e2e_full_output"] = run_e2e(args, arena, idx, off, lookups)
    if rank == 0 and world == 1 and not args.skip_cpu:
        res["cpu_baseline"] = cpu_baseline(args, arena, idx, rows, backward=True)
    if rank == 0 and world == 1 and args.alpha > 0 and not args.skip_uniform:
        # the HBM-bound leg: the same arena and kernels under uniform indices (no cache help), timed in
        # this run.  Last, because it overwrites the request.
        fill_indices(args, idx, rows, 0.0, dev)
        ms_u, roofs_u = legs("uniform")
        ms_step_u = ev_time(step, max(3, args.steps // 2))
        res["roofline_uniform"] = dict(roofs_u, kernels_ms={k: round(v, 4) for k, v in ms_u.items()},
                                       ms_per_step=round(ms_step_u, 4), value=lookups / (ms_step_u * 1e-3),
                                       note="alpha = 0 on the same arena: every row read misses the caches, "
                                            "DRAM traffic ~ algorithmic bytes")
        # the dominant kernel's HBM-bound figure next to its skewed one, for a reader of `roofline` alone
        dom_u = roofs_u.get(res["roofline"].get("dominant_kernel"), {})
        res["roofline"]["uniform_case"] = {k: dom_u.get(k) for k in ("ms", "frac", "frac_dram", "frac_dram_compulsory")}
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, arena, idx, off, lookups):
    """Same step through the C-ABI host-buffer entries: indices/offsets start in pinned HOST memory and the step's
    result ends in pinned HOST memory; H2D + kernels + D2H are all inside the timed region.  Two forms:
    loss (pb200_tbe_step_host_loss): what comes back is the step's scalar result, one sum per table — the pooled
      vectors stay in HBM as they do in the reference's GPU loop (pytorch_emb.py:48-69) and DLRM step;
    full_output (pb200_tbe_step_host): all pooled vectors are copied back ([T, B, D], PCIe-bound)."""
    import ctypes as C
    from param_b200 import _cabi
    T, B, L, D = args.tables, args.batch, args.bag, args.dim
    lib = _cabi.load()
    g = min(args.e2e_group, T)
    h_idx = torch.empty(idx.numel(), dtype=torch.int64).pin_memory()
    h_off = torch.empty(off.numel(), dtype=torch.int64).pin_memory()
    h_idx.copy_(idx)
    h_off.copy_(off)
    h_out = torch.empty((T, B, D), dtype=torch.float32).pin_memory()   # [T, B, D]: contiguous D2H per group
    h_loss = torch.zeros(T, dtype=torch.float64).pin_memory()
    tro_h = arena.row_offsets.cpu()
    ctx = C.c_void_p()
    _cabi.check(lib.pb200_host_ctx_create(C.byref(ctx), g * B * L + 16, g * B, D), "host_ctx_create")

    def call_full(do_bwd=1):
        _cabi.check(lib.pb200_tbe_step_host(ctx, arena.weights.data_ptr(), arena.row_offsets.data_ptr(),
                                            tro_h.data_ptr(), T, D, h_idx.data_ptr(), h_idx.numel(),
                                            h_off.data_ptr(), B, 0, h_out.data_ptr(), 1, g,
                                            do_bwd, C.c_float(-args.lr)), "tbe_step_host")

    def call_loss(do_bwd=1):
        _cabi.check(lib.pb200_tbe_step_host_loss(ctx, arena.weights.data_ptr(), arena.row_offsets.data_ptr(),
                                                 tro_h.data_ptr(), T, D, h_idx.data_ptr(), h_idx.numel(),
                                                 h_off.data_ptr(), B, 0, h_loss.data_ptr(), g,
                                                 do_bwd, C.c_float(-args.lr)), "tbe_step_host_loss")

    def timed(call):
        for _ in range(2):
            call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = max(3, args.steps // 2)
        for _ in range(n):
            call()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n

    h2d = int(h_idx.numel() * 8 + h_off.numel() * 8)
    dt_loss = timed(call_loss)
    # the sums that come back are those of what the full-output form hands back (forward only: same arena state)
    call_loss(0)
    call_full(0)
    ref = h_out[:8].to(torch.float64).sum(dim=(1, 2))
    loss_ok = bool(torch.allclose(h_loss[:8], ref, rtol=1e-9, atol=1e-9))
    dt_full = timed(call_full)
    lib.pb200_host_ctx_destroy(ctx)
    e2e = {"value": lookups / dt_loss, "unit": UNIT, "ms_per_step": dt_loss * 1e3,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(h_loss.numel() * 8),
           "result": "per-table loss (sum of the pooled vectors, float64), checked against the full-output form: %s"
                     % loss_ok,
           "path": "pb200_tbe_step_host_loss (C ABI, pinned host buffers, %d-table pipeline groups): "
                   "H2D indices+offsets -> lookup fwd -> per-table sum -> scatter-add bwd -> D2H loss" % g}
    full = {"value": lookups / dt_full, "unit": UNIT, "ms_per_step": dt_full * 1e3,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(h_out.numel() * 4),
            "path": "pb200_tbe_step_host (C ABI, pinned host buffers, %d-table pipeline groups): "
                    "H2D indices+offsets -> lookup fwd -> D2H pooled -> scatter-add bwd" % g}
    return e2e, full


def cpu_baseline(args, arena, idx, rows, backward):
    """Reference CPU path (torch.nn.EmbeddingBag on the host cores, measure_cpu loop shape) on a
    bounded sample: the first --cpu-tables tables of the same workload, same indices."""
    from oracle import ref_torch_cpu  # bench-only import of oracle/
    T, B, L, D = args.tables, args.batch, args.bag, args.dim
    n = min(args.cpu_tables, T)
    ws = [arena.table(t).cpu() for t in range(n)]
    ids = [idx[t * B * L:(t + 1) * B * L].cpu() for t in range(n)]
    offs = [torch.arange(B, dtype=torch.int64) * L for _ in range(n)]
    sec, threads = ref_torch_cpu.time_embeddingbag_cpu(ws, ids, offs, steps=3, warmups=1, backward=backward)
    sec_f, _ = ref_torch_cpu.time_embeddingbag_cpu(ws, ids, offs, steps=5, warmups=2, backward=False)
    return {"value": n * B * L / sec, "unit": UNIT, "cores": threads, "host_cpus": os.cpu_count(),
            "fwd_only_value": n * B * L / sec_f, "fwd_only_ms_per_step_sample": sec_f * 1e3,
            "kind": "reference",
            "sample": f"torch.nn.EmbeddingBag(mode=sum, sparse=True) fwd+bwd on the host, first {n} of {T} tables "
                      f"({rows} rows x {D}), same Zipf indices, 3 steps after 1 warm-up (measure_cpu loop, "
                      "train/compute/pt/pytorch_emb.py:37-45)",
            "ms_per_step_sample": sec * 1e3}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, on a
    bounded sample of the same workload.  Rank 0 only."""
    from oracle import ref_torch_cpu
    from param_b200.compute.pt.pytorch_emb import zipf_cdf
    import numpy as np
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    T, B, L, D = args.tables, args.batch, args.bag, args.dim
    rows = args.rows
    tag = "cfg2"
    if args.gpus > 1:
        # N > 1: the b200 arm runs cfg4 (64 tables/GPU, local batch 8192): the CPU arm samples the
        # per-rank lookup of that workload (tables of one rank over the global batch)
        T = int(os.environ.get("PB200_TABLES_PER_GPU", 64))
        B = int(os.environ.get("PB200_LOCAL_BATCH", 8192)) * args.gpus
        rows = int(os.environ.get("PB200_ROWS", min(args.rows, 2_000_000)))
        tag = "cfg4 (one rank's tables over the global batch)"
    n = min(args.cpu_tables, T)
    rng = np.random.default_rng(2026)
    cdf = zipf_cdf(args.alpha, rows) if args.alpha > 0 else np.linspace(1.0 / rows, 1.0, rows)
    ws, ids, offs = [], [], []
    for t in range(n):
        ws.append((torch.rand(rows, D) * 2 - 1) * (1.0 / rows) ** 0.5)
        ids.append(torch.from_numpy(np.searchsorted(cdf, rng.random(B * L), side="right").clip(0, rows - 1).astype(np.int64)))
        offs.append(torch.arange(B, dtype=torch.int64) * L)
    sec, threads = ref_torch_cpu.time_embeddingbag_cpu(ws, ids, offs, steps=args.steps, warmups=args.warmup,
                                                       backward=True)
    v = n * B * L / sec
    sample = (f"first {n} of {T} tables per step ({rows} rows x {D}, batch {B}, bag {L}, Zipf {args.alpha}), "
              "torch.nn.EmbeddingBag CPU fwd+bwd, all host threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{tag} sample: {sample}", "tables": T, "rows_per_table": rows, "dim": D,
                   "batch": B, "bag": L, "alpha": args.alpha},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "host_cpus": os.cpu_count(),
                         "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main(argv=None):
    args = parse_args(argv)
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world != args.gpus and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} must be launched with torchrun --nproc-per-node {args.gpus}")
    if world > 1:
        from bench_dist import run_dist
        run_dist(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
