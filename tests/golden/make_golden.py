#!/usr/bin/env python3
"""Generate the golden vectors that pin oracle/ against the REFERENCE's own code.

Run in the build container only (it imports /root/reference, which does not exist on the GPU
box):      python tests/golden/make_golden.py

The reference holds no golden vectors / known-answer tests for this path, so the
fixtures are produced by running the reference's own call sites here:

  embbag_torch_cpu.npz  torch.nn.EmbeddingBag on CPU — the exact op the reference calls at
                        train/compute/pt/pytorch_emb.py:179,40 and train/comms/pt/dlrm.py:379-380 —
                        forward (sum / mean / per_sample_weights) and autograd dense backward.
  init_indices_ref.npz  the reference's init_indices (train/compute/pt/pytorch_emb.py:138-160),
                        imported from /root/reference, uniform and Zipf branches, seeded.
  dlrm_sparse_ref.npz   the reference's calculateLengths / splitPerTable / lengthsToOffsets
                        (train/comms/pt/dlrm.py:226-251, 430-504), imported from /root/reference.
  a2a_gloo_ref.npz      3 gloo ranks: c10d all_to_all_single with uneven splits (the call at
                        train/comms/pt/pytorch_dist_backend.py:336-351) and the reference's own
                        All2Allv_Req / All2Allv_Wait autograd Functions + torch.cat
                        (train/comms/pt/dlrm.py:86-218, 858-878, 1253), forward and backward.
  tbe_optim_torch.npz   the reference's TBE benchmark step — forward, create_grad = ones_like, backward with
                        the optimizer fused in (train/compute/python/workloads/pytorch/
                        split_table_batched_embeddings_ops.py:311-324) — computed WITHOUT fbgemm_gpu (absent) by
                        the per-table nn.EmbeddingBag loop + torch.optim.SGD / torch.optim.Adagrad.  With a
                        ones gradient every row gradient is constant along the row, where fbgemm's rowwise
                        Adagrad (mean over the row of g^2) coincides with the elementwise torch.optim.Adagrad.
                        Generate alone with:  python tests/golden/make_golden.py tbe_optim
  dlrm_report_ref.npz   the reference's own report printer — commsDLRMBench.reportBenchTime
                        (train/comms/pt/dlrm.py:1011-1193) — run on seeded synthetic samples of its 21 regions for
                        1 and for 3 ranks (the all_gather it calls is stood in by a copy of the per-rank sample
                        arrays); the printed text is the fixture param_b200.comms.pt.dlrm.format_report must equal.
                        Generate alone with:  python tests/golden/make_golden.py dlrm_report
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")


def _ref_paths():
    """Make `param_bench.train.comms.pt` and the script-dir imports of dlrm.py resolvable."""
    alias_root = Path(tempfile.gettempdir()) / "pb200_ref_alias"
    alias_root.mkdir(exist_ok=True)
    link = alias_root / "param_bench"
    if not link.exists():
        link.symlink_to(REF)
    sys.dont_write_bytecode = True
    for p in (str(alias_root), str(REF / "train" / "comms" / "pt"), str(REF / "train" / "compute" / "pt")):
        if p not in sys.path:
            sys.path.insert(0, p)


# ----------------------------------------------------------------------------------------------
def gen_embbag():
    rng = np.random.default_rng(20260925)
    cases = []

    def case(rows, dim, offsets, n_idx, mode="sum", weighted=False, idx=None):
        if idx is None:
            idx = rng.integers(0, rows, size=n_idx)
        cases.append(dict(rows=rows, dim=dim, offsets=np.asarray(offsets, np.int64),
                          indices=np.asarray(idx, np.int64), mode=mode, weighted=weighted))

    # 0: empty bags, trailing bag runs to the end, duplicate indices inside a bag
    case(37, 8, [0, 0, 3, 3, 7, 7], 11, idx=[5, 5, 9, 0, 36, 36, 36, 1, 2, 2, 30])
    # 1: cfg1-like (dim 64, fixed bag 20)
    case(1000, 64, np.arange(16) * 20, 320)
    # 2: dataset B shape (dim 56, bag 34)
    case(500, 56, np.arange(8) * 34, 272)
    # 3: dim 128, mean pooling, ragged incl. one empty bag
    case(300, 128, [0, 4, 4, 9, 30], 41, mode="mean")
    # 4: per_sample_weights (sum mode), dim 12
    case(64, 12, [0, 2, 5, 6], 9, weighted=True)
    # 5: odd dim -> scalar path, bag size 1
    case(40, 7, np.arange(10), 10)
    # 6: dim 256 (two float4 chunks per lane), ragged, bag longer than 32
    case(200, 256, [0, 40, 41, 41, 80], 97)
    # 7: dim 4, tiny
    case(10, 4, [0, 1, 3], 6)
    # 8: all bags empty except the last
    case(20, 16, [0, 0, 0, 0], 5)
    # 9: dim 512 upper edge of the vector path
    case(50, 512, [0, 3, 10], 12, mode="mean")

    out = {"n_cases": np.int64(len(cases))}
    for k, c in enumerate(cases):
        torch.manual_seed(100 + k)
        emb = nn.EmbeddingBag(c["rows"], c["dim"], mode=c["mode"])  # default N(0,1) init
        idx = torch.from_numpy(c["indices"])
        off = torch.from_numpy(c["offsets"])
        psw = torch.rand(idx.numel()) + 0.5 if c["weighted"] else None
        res = emb(idx, off, per_sample_weights=psw)
        g = torch.randn_like(res)
        res.backward(g)
        p = f"c{k}_"
        out[p + "weight"] = emb.weight.detach().numpy().copy()
        out[p + "indices"] = c["indices"]
        out[p + "offsets"] = c["offsets"]
        out[p + "mode"] = np.array(c["mode"])
        out[p + "psw"] = psw.numpy() if psw is not None else np.zeros(0, np.float32)
        out[p + "out"] = res.detach().numpy().copy()
        out[p + "grad_out"] = g.numpy().copy()
        out[p + "grad_weight"] = emb.weight.grad.numpy().copy()
    np.savez_compressed(HERE / "embbag_torch_cpu.npz", **out)
    print("embbag_torch_cpu.npz:", len(cases), "cases")


# ----------------------------------------------------------------------------------------------
def gen_init_indices():
    _ref_paths()
    import pytorch_emb as ref_emb  # /root/reference/train/compute/pt/pytorch_emb.py

    out = {}
    torch.manual_seed(7)
    out["uniform_seed"] = np.int64(7)
    out["uniform_args"] = np.array([5000, 64, 6], np.int64)  # features, batch, nnz
    out["uniform"] = ref_emb.init_indices(0.0, 5000, 64, 6).numpy()
    np.random.seed(1234)
    out["zipf_seed"] = np.int64(1234)
    out["zipf_args"] = np.array([2000, 48, 5], np.int64)
    out["zipf_alpha"] = np.float64(1.15)
    out["zipf"] = ref_emb.init_indices(1.15, 2000, 48, 5).numpy()
    np.savez_compressed(HERE / "init_indices_ref.npz", **out)
    print("init_indices_ref.npz")


# ----------------------------------------------------------------------------------------------
def gen_dlrm_sparse():
    _ref_paths()
    import dlrm as ref_dlrm  # /root/reference/train/comms/pt/dlrm.py

    rng = np.random.default_rng(99)
    out = {}
    # --- calculateLengths: per-feature offsets/indices -> flat lengths/indices (dlrm.py:226-242)
    feat, b = 4, 5
    offs, idxs = [], []
    for f in range(feat):
        lens = rng.integers(0, 4, size=b)
        lens[rng.integers(0, b)] = 0
        o = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
        offs.append(torch.from_numpy(o))
        idxs.append(torch.from_numpy(rng.integers(0, 100, size=int(lens.sum())).astype(np.int64)))
    lengths, indices = ref_dlrm.calculateLengths(feat, offs, idxs)
    out["cl_feat"], out["cl_batch"] = np.int64(feat), np.int64(b)
    for f in range(feat):
        out[f"cl_off{f}"] = offs[f].numpy()
        out[f"cl_idx{f}"] = idxs[f].numpy()
    out["cl_lengths"] = lengths.numpy()
    out["cl_indices"] = indices.numpy()

    # --- splitPerTable + lengthsToOffsets (dlrm.py:430-504, 245-251)
    for name, (W, Tl, bb) in {"a": (3, 2, 4), "b": (2, 3, 5), "c": (4, 1, 3)}.items():
        lens = rng.integers(0, 5, size=W * Tl * bb).astype(np.int64)
        ind = rng.integers(0, 1000, size=int(lens.sum())).astype(np.int64)
        offsets, indices_pt = ref_dlrm.paramDLRM_Net.splitPerTable(
            None, torch.from_numpy(lens), torch.from_numpy(ind), bb, Tl, W, 0, torch.device("cpu"))
        out[f"sp_{name}_dims"] = np.array([W, Tl, bb], np.int64)
        out[f"sp_{name}_lengths"] = lens
        out[f"sp_{name}_indices"] = ind
        for f in range(Tl):
            out[f"sp_{name}_off{f}"] = offsets[f].numpy().astype(np.int64)
            out[f"sp_{name}_idx{f}"] = indices_pt[f].numpy().astype(np.int64)
    np.savez_compressed(HERE / "dlrm_sparse_ref.npz", **out)
    print("dlrm_sparse_ref.npz")


# ----------------------------------------------------------------------------------------------
W_A2A = 3
A2A_SPLITS = [[1, 2, 3], [0, 4, 2], [5, 0, 1]]  # elements rank s sends to rank d
POOL_N, POOL_E, POOL_TG = 7, 4, 5


def _a2a_worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _ref_paths()
    import dlrm as ref_dlrm

    res = {}
    # (a) raw c10d all_to_all_single, uneven splits, position-coded payloads
    in_splits = A2A_SPLITS[rank]
    out_splits = [A2A_SPLITS[s][rank] for s in range(world)]
    for dt, tag in ((torch.int64, "i64"), (torch.float32, "f32")):
        inp = (torch.arange(sum(in_splits)) + 1000 * rank).to(dt)
        outp = torch.empty(sum(out_splits), dtype=dt)
        dist.all_to_all_single(outp, inp, out_splits, in_splits)
        res[f"raw_{tag}_in"] = inp.numpy()
        res[f"raw_{tag}_out"] = outp.numpy()

    # (b) the reference's pooled-embedding exchange: All2Allv_Req / All2Allv_Wait + cat
    class StubBackend:
        def sync_barrier(self, ca):
            dist.barrier()

        def complete_accel_ops(self, ca):
            pass

        def all_to_allv(self, ca, retFlag=False):
            return dist.all_to_all_single(ca.opTensor, ca.ipTensor, ca.opTensor_split,
                                          ca.ipTensor_split, async_op=True)

    get_split = lambda n, r, w: ref_dlrm.paramDLRM_Net.get_split_lengths_by_len(None, n, r, w)  # noqa: E731
    _, n_emb_per_rank = get_split(POOL_TG, rank, world)
    bench = types.SimpleNamespace(
        measured_regions={"fwd_a2a": {"memory": []}, "bwd_a2a": {"memory": []}},
        commDetails=[], collectiveArgs=types.SimpleNamespace(timers={}),
        backendFuncs=StubBackend(), my_size=world,
        paramNN=types.SimpleNamespace(get_split_lengths_by_len=get_split))
    bench.myreq = ref_dlrm.Request(bench)
    T_l = n_emb_per_rank[rank]
    g = torch.Generator().manual_seed(50 + rank)
    ly = torch.randn(T_l, POOL_N, POOL_E, generator=g).requires_grad_()
    dims_sum_per_rank = [t * POOL_E for t in n_emb_per_rank]
    req = ref_dlrm.commsDLRMBench.alltoallv(bench, ly, rank, dims_sum_per_rank, n_emb_per_rank)
    B = req.wait()
    tempB = torch.cat(B, dim=1)
    C = torch.randn(tempB.shape, generator=g)
    tempB.backward(C)
    res["pool_ly"] = ly.detach().numpy()
    res["pool_out"] = tempB.detach().numpy()
    res["pool_gradout"] = C.numpy()
    res["pool_gradin"] = ly.grad.numpy()
    res["pool_tables_split"] = np.array(n_emb_per_rank, np.int64)
    lN, gNS = get_split(POOL_N, rank, world)
    res["pool_batch_split"] = np.array(gNS, np.int64)
    np.savez(os.path.join(tmpdir, f"rank{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


def gen_a2a():
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_a2a_worker, args=(W_A2A, 29611, td), nprocs=W_A2A, join=True)
        out = {"world": np.int64(W_A2A), "splits": np.array(A2A_SPLITS, np.int64),
               "pool_dims": np.array([POOL_N, POOL_E, POOL_TG], np.int64)}
        for r in range(W_A2A):
            d = np.load(os.path.join(td, f"rank{r}.npz"))
            for k in d.files:
                out[f"r{r}_{k}"] = d[k]
    np.savez_compressed(HERE / "a2a_gloo_ref.npz", **out)
    print("a2a_gloo_ref.npz")


def gen_tbe_optim():
    """3 training steps of a 3-table TBE op on CPU with torch only (no oracle code involved)."""
    torch.manual_seed(7)
    rng = np.random.default_rng(20261017)
    rows, dim, B, lr, eps, steps = [60, 20, 150], 32, 32, 0.05, 1.0e-8, 3
    T = len(rows)
    w0 = [torch.randn(r, dim) * 0.1 for r in rows]
    out = {"rows": np.array(rows), "dim": dim, "batch": B, "lr": lr, "eps": eps, "steps": steps,
           "w0": torch.cat(w0).numpy()}
    requests = []
    for s in range(steps):
        lens = rng.integers(0, 9, size=T * B)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        idx = np.concatenate([rng.integers(0, rows[t], size=int(lens[t * B:(t + 1) * B].sum()))
                              for t in range(T)]).astype(np.int64)
        requests.append((idx, offsets))
        out[f"s{s}_indices"], out[f"s{s}_offsets"] = idx, offsets
    for name, make_opt in (("sgd", lambda ps: torch.optim.SGD(ps, lr=lr)),
                           ("adagrad", lambda ps: torch.optim.Adagrad(ps, lr=lr, eps=eps,
                                                                      initial_accumulator_value=0.0))):
        embs = [nn.EmbeddingBag(r, dim, mode="sum", _weight=w.clone()) for r, w in zip(rows, w0)]
        opt = make_opt([e.weight for e in embs])
        for s, (idx, offsets) in enumerate(requests):
            opt.zero_grad()
            pooled = []
            for t, e in enumerate(embs):                     # the per-table loop of dlrm.py:363-388
                lo, hi = offsets[t * B], offsets[(t + 1) * B]
                pooled.append(e(torch.from_numpy(idx[lo:hi]), torch.from_numpy(offsets[t * B:(t + 1) * B] - lo)))
            fwd_out = torch.cat(pooled, dim=1)               # [B, T*dim], the TBE output layout
            out[f"{name}_s{s}_out"] = fwd_out.detach().numpy().copy()
            fwd_out.backward(torch.ones_like(fwd_out))       # create_grad (:315-316) + backward (:318-324)
            opt.step()
            out[f"{name}_s{s}_w"] = torch.cat([e.weight.detach() for e in embs]).numpy().copy()
        if name == "adagrad":
            # elementwise accumulator; constant along a row here, so column 0 is fbgemm's rowwise state
            out["adagrad_state"] = torch.cat([opt.state[e.weight]["sum"][:, 0] for e in embs]).numpy().copy()
            acc = torch.cat([opt.state[e.weight]["sum"] for e in embs])
            assert bool((acc == acc[:, :1]).all()), "the accumulator must be constant along a row"
    np.savez_compressed(HERE / "tbe_optim_torch.npz", **out)
    print("wrote tbe_optim_torch.npz")


def gen_tbe_rowwise():
    """fbgemm's exact rowwise Adagrad on NON-constant gradients, computed with torch only and independently of
    oracle/: the dense gradient of every table comes from torch autograd (nn.EmbeddingBag, sparse=False, a random
    incoming gradient, one step with per-sample weights), the update is written out in float64 exactly as
    published — m[row] += mean_d(g[row, d]^2);  w[row] -= lr / (sqrt(m[row]) + eps) * g[row] — for the rows the
    batch touched.  Three steps, state carried over."""
    torch.manual_seed(11)
    rng = np.random.default_rng(20261018)
    rows, dim, B, lr, eps, steps = [90, 17, 400], 24, 48, 0.07, 1.0e-8, 3
    T = len(rows)
    w = [(torch.randn(r, dim) * 0.1).double() for r in rows]
    m = [torch.zeros(r, dtype=torch.float64) for r in rows]
    out = {"rows": np.array(rows), "dim": dim, "batch": B, "lr": lr, "eps": eps, "steps": steps,
           "w0": torch.cat(w).float().numpy()}
    for s in range(steps):
        lens = rng.integers(0, 11, size=T * B)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        idx = np.concatenate([rng.integers(0, rows[t], size=int(lens[t * B:(t + 1) * B].sum()))
                              for t in range(T)]).astype(np.int64)
        psw = (rng.random(idx.size).astype(np.float32) + 0.5) if s == 1 else None
        grad = torch.from_numpy(rng.standard_normal((B, T * dim)).astype(np.float32))
        out[f"s{s}_indices"], out[f"s{s}_offsets"], out[f"s{s}_grad"] = idx, offsets, grad.numpy()
        out[f"s{s}_psw"] = psw if psw is not None else np.zeros(0, np.float32)
        pooled, embs = [], []
        for t in range(T):
            e = nn.EmbeddingBag(rows[t], dim, mode="sum", _weight=w[t].float().clone())
            lo, hi = offsets[t * B], offsets[(t + 1) * B]
            pw = None if psw is None else torch.from_numpy(psw[lo:hi])
            pooled.append(e(torch.from_numpy(idx[lo:hi]), torch.from_numpy(offsets[t * B:(t + 1) * B] - lo),
                            per_sample_weights=pw))
            embs.append(e)
        fwd_out = torch.cat(pooled, dim=1)
        fwd_out.backward(grad)
        for t, e in enumerate(embs):
            g = e.weight.grad.double()                                   # dense dW of the table, from autograd
            touched = torch.zeros(rows[t], dtype=torch.bool)
            lo, hi = offsets[t * B], offsets[(t + 1) * B]
            touched[torch.from_numpy(idx[lo:hi])] = True
            m[t] = torch.where(touched, m[t] + (g * g).mean(dim=1), m[t])
            step = lr / (m[t].sqrt() + eps)
            w[t] = torch.where(touched[:, None], w[t] - step[:, None] * g, w[t])
        out[f"s{s}_out"] = fwd_out.detach().numpy().copy()
        out[f"s{s}_w"] = torch.cat(w).numpy().copy()                    # float64
        out[f"s{s}_state"] = torch.cat(m).numpy().copy()
    np.savez_compressed(HERE / "tbe_rowwise_adagrad_torch.npz", **out)
    print("wrote tbe_rowwise_adagrad_torch.npz")


def gen_dlrm_report():
    """The text the reference prints for known per-rank samples (both of its tables: percentiles over all samples,
    and over the per-rank means)."""
    import contextlib
    import io
    _ref_paths()
    import dlrm as ref_dlrm  # /root/reference/train/comms/pt/dlrm.py
    rng = np.random.default_rng(20261018)
    out = {}
    for world in (1, 3):
        iters, warm = 6, 2
        bench = ref_dlrm.commsDLRMBench()
        bench.initTimers()
        names = list(bench.measured_regions)
        lat = (rng.random((world, len(names), iters)) * 5000.0).astype(np.float32)       # microseconds
        mem = rng.integers(0, 1 << 20, size=(world, len(names), warm + iters)).astype(np.int64)
        mem[:, [i for i, n in enumerate(names) if "xchg" not in n and "a2a" not in n and "_ar" not in n], :] = 0
        for i, n in enumerate(names):
            bench.measured_regions[n]["samples"] = [float(v) for v in lat[0, i]]
            bench.measured_regions[n]["memory"] = [int(v) for v in mem[0, i]]
        calls = []

        class _Gather:        # stands in for backendFuncs.all_gather: rank r's tensor is row r of the prepared arrays
            def all_gather(self, ca):
                src = lat if not calls else mem
                calls.append(1)
                for r in range(world):
                    ca.opTensor[r].copy_(torch.from_numpy(src[r]).to(ca.opTensor[r].dtype))

            def complete_accel_ops(self, ca):
                pass

        bench.backendFuncs = _Gather()
        bench.collectiveArgs = types.SimpleNamespace()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            bench.reportBenchTime(0, warm, iters, world, "cpu")
        out[f"w{world}_lat"] = lat
        out[f"w{world}_mem"] = mem
        out[f"w{world}_warm"] = np.int64(warm)
        out[f"w{world}_text"] = np.array(buf.getvalue())
    out["regions"] = np.array(names)
    np.savez_compressed(HERE / "dlrm_report_ref.npz", **out)
    print("wrote dlrm_report_ref.npz")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "dlrm_report":
        gen_dlrm_report()
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "tbe_optim":      # torch only: no reference import needed
        gen_tbe_optim()
        gen_tbe_rowwise()
        raise SystemExit(0)
    if not REF.exists():
        raise SystemExit("/root/reference not present: golden vectors can only be regenerated "
                         "in the build container")
    gen_embbag()
    gen_init_indices()
    gen_dlrm_sparse()
    gen_a2a()
    gen_tbe_optim()
    gen_tbe_rowwise()
    gen_dlrm_report()
