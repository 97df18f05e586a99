"""numpy front-end of oracle/param_oracle.c — TEST INFRASTRUCTURE ONLY.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Nothing under param_b200/ may import this module.  See param_oracle.c for the reference
citations of every function and for how the oracle is pinned (tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libparam_oracle.so"
_SRC = _HERE / "param_oracle.c"
_lib = None

POOL = {"sum": 0, "mean": 1}


def build(force: bool = False) -> Path:
    if force or not _SO.exists() or _SO.stat().st_mtime < _SRC.stat().st_mtime:
        cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
               "-o", str(_SO), str(_SRC), "-lm"]
        subprocess.run(cmd, check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
    return _lib


def _p(a, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


def embbag_fwd(weight, indices, offsets, mode="sum", psw=None, include_last_offset=False,
               threads=1):
    weight, indices, offsets, psw = _f32(weight), _i64(indices), _i64(offsets), _f32(psw)
    n_bags = offsets.size - (1 if include_last_offset else 0)
    rows, dim = weight.shape
    out = np.empty((n_bags, dim), dtype=np.float32)
    _load().oracle_embbag_fwd(_p(weight, C.c_float), C.c_int64(rows), C.c_int32(dim),
                              _p(indices, C.c_int64), C.c_int64(indices.size),
                              _p(offsets, C.c_int64), C.c_int64(n_bags),
                              C.c_int32(1 if include_last_offset else 0), _p(psw, C.c_float),
                              C.c_int32(POOL[mode]), _p(out, C.c_float), C.c_int64(dim),
                              C.c_int32(threads))
    return out


def embbag_bwd(rows, dim, indices, offsets, grad_out, mode="sum", psw=None,
               include_last_offset=False, scale=1.0, dtype=np.float32, dst=None):
    """Dense-equivalent gradient [rows, dim] (fp32 sequential, or float64 when dtype=np.float64)."""
    indices, offsets, psw, grad_out = _i64(indices), _i64(offsets), _f32(psw), _f32(grad_out)
    n_bags = offsets.size - (1 if include_last_offset else 0)
    if dst is None:
        dst = np.zeros((rows, dim), dtype=dtype)
    d32 = dst if dst.dtype == np.float32 else None
    d64 = dst if dst.dtype == np.float64 else None
    _load().oracle_embbag_bwd(_p(d32, C.c_float), _p(d64, C.c_double), C.c_int32(dim),
                              _p(indices, C.c_int64), C.c_int64(indices.size),
                              _p(offsets, C.c_int64), C.c_int64(n_bags),
                              C.c_int32(1 if include_last_offset else 0), _p(psw, C.c_float),
                              C.c_int32(POOL[mode]), _p(grad_out, C.c_float), C.c_int64(dim),
                              C.c_float(scale))
    return dst


def tbe_fwd(weights, table_row_offsets, dim, indices, offsets, batch, mode="sum", psw=None,
            layout="BTD", threads=1):
    weights, indices, offsets, psw = _f32(weights), _i64(indices), _i64(offsets), _f32(psw)
    tro = _i64(table_row_offsets)
    T = tro.size - 1
    if layout == "BTD":
        out = np.empty((batch, T * dim), dtype=np.float32)
        st_t, st_b = dim, T * dim
    else:
        out = np.empty((T, batch, dim), dtype=np.float32)
        st_t, st_b = batch * dim, dim
    _load().oracle_tbe_fwd(_p(weights, C.c_float), _p(tro, C.c_int64), C.c_int32(T),
                           C.c_int32(dim), _p(indices, C.c_int64), _p(offsets, C.c_int64),
                           C.c_int64(batch), _p(psw, C.c_float), C.c_int32(POOL[mode]),
                           _p(out, C.c_float), C.c_int64(st_t), C.c_int64(st_b), C.c_int32(threads))
    return out


def tbe_bwd(total_rows, table_row_offsets, dim, indices, offsets, batch, grad_out, mode="sum",
            psw=None, layout="BTD", scale=1.0, dtype=np.float32, dst=None):
    indices, offsets, psw, grad_out = _i64(indices), _i64(offsets), _f32(psw), _f32(grad_out)
    tro = _i64(table_row_offsets)
    T = tro.size - 1
    st_t, st_b = (dim, T * dim) if layout == "BTD" else (batch * dim, dim)
    if dst is None:
        dst = np.zeros((total_rows, dim), dtype=dtype)
    d32 = dst if dst.dtype == np.float32 else None
    d64 = dst if dst.dtype == np.float64 else None
    _load().oracle_tbe_bwd(_p(d32, C.c_float), _p(d64, C.c_double), _p(tro, C.c_int64),
                           C.c_int32(T), C.c_int32(dim), _p(indices, C.c_int64),
                           _p(offsets, C.c_int64), C.c_int64(batch), _p(psw, C.c_float),
                           C.c_int32(POOL[mode]), _p(grad_out, C.c_float), C.c_int64(st_t),
                           C.c_int64(st_b), C.c_float(scale))
    return dst


def all_to_all_single(inputs, in_splits):
    """inputs: list of W 1-D arrays (same dtype); in_splits[s][d] = ELEMENTS rank s sends to d.
    Returns the list of W output arrays (c10d all_to_all_single semantics)."""
    W = len(inputs)
    dt = inputs[0].dtype
    es = dt.itemsize
    splits = np.asarray(in_splits, dtype=np.int64).reshape(W, W) * es
    flat_in = np.concatenate([np.ascontiguousarray(x).view(np.uint8).reshape(-1) for x in inputs]) \
        if sum(x.size for x in inputs) else np.zeros(0, np.uint8)
    in_off = np.zeros(W + 1, dtype=np.int64)
    in_off[1:] = np.cumsum([x.size * es for x in inputs])
    out_sizes = splits.sum(axis=0)
    out_off = np.zeros(W + 1, dtype=np.int64)
    out_off[1:] = np.cumsum(out_sizes)
    flat_out = np.zeros(int(out_off[-1]), dtype=np.uint8)
    _load().oracle_all_to_all_single(C.c_int32(W), _p(flat_in, C.c_uint8), _p(in_off, C.c_int64),
                                     _p(np.ascontiguousarray(splits), C.c_int64),
                                     _p(flat_out, C.c_uint8), _p(out_off, C.c_int64))
    return [flat_out[out_off[d]:out_off[d + 1]].view(dt).copy() for d in range(W)]


def pooled_a2a_fwd(pooled, batch_split, tables_split, emb_dim):
    """pooled[r]: [T_r, N, E] per rank -> list of [lN_j, T_global*E]."""
    W = len(pooled)
    bs, ts = _i64(batch_split), _i64(tables_split)
    Tg = int(ts.sum())
    flat = np.concatenate([_f32(p).reshape(-1) for p in pooled])
    poff = np.zeros(W + 1, np.int64)
    poff[1:] = np.cumsum([p.size for p in pooled])
    ooff = np.zeros(W + 1, np.int64)
    ooff[1:] = np.cumsum(bs * Tg * emb_dim)
    out = np.zeros(int(ooff[-1]), np.float32)
    _load().oracle_pooled_a2a_fwd(C.c_int32(W), C.c_int32(emb_dim), _p(bs, C.c_int64),
                                  _p(ts, C.c_int64), _p(flat, C.c_float), _p(poff, C.c_int64),
                                  _p(out, C.c_float), _p(ooff, C.c_int64))
    return [out[ooff[j]:ooff[j + 1]].reshape(int(bs[j]), Tg * emb_dim).copy() for j in range(W)]


def pooled_a2a_bwd(grads, batch_split, tables_split, emb_dim):
    """grads[j]: [lN_j, T_global*E] per rank -> list of [T_r, N, E]."""
    W = len(grads)
    bs, ts = _i64(batch_split), _i64(tables_split)
    N = int(bs.sum())
    flat = np.concatenate([_f32(g).reshape(-1) for g in grads])
    goff = np.zeros(W + 1, np.int64)
    goff[1:] = np.cumsum([g.size for g in grads])
    ooff = np.zeros(W + 1, np.int64)
    ooff[1:] = np.cumsum(ts * N * emb_dim)
    out = np.zeros(int(ooff[-1]), np.float32)
    _load().oracle_pooled_a2a_bwd(C.c_int32(W), C.c_int32(emb_dim), _p(bs, C.c_int64),
                                  _p(ts, C.c_int64), _p(flat, C.c_float), _p(goff, C.c_int64),
                                  _p(out, C.c_float), _p(ooff, C.c_int64))
    return [out[ooff[r]:ooff[r + 1]].reshape(int(ts[r]), N, emb_dim).copy() for r in range(W)]


def split_per_table(lengths, indices, world, tables_local, local_batch):
    lengths, indices = _i64(lengths).reshape(-1), _i64(indices).reshape(-1)
    n = world * tables_local * local_batch
    lengths_out = np.empty(n, np.int64)
    offsets_out = np.empty(n + 1, np.int64)
    indices_out = np.empty_like(indices)
    _load().oracle_split_per_table(_p(lengths, C.c_int64), _p(indices, C.c_int64),
                                   C.c_int32(world), C.c_int32(tables_local),
                                   C.c_int64(local_batch), _p(lengths_out, C.c_int64),
                                   _p(offsets_out, C.c_int64), _p(indices_out, C.c_int64))
    return lengths_out.reshape(tables_local, world * local_batch), offsets_out, indices_out


def calculate_lengths(offsets, n_indices):
    offsets = _i64(offsets)
    out = np.empty(offsets.size, np.int64)
    _load().oracle_calculate_lengths(_p(offsets, C.c_int64), C.c_int64(offsets.size),
                                     C.c_int64(n_indices), _p(out, C.c_int64))
    return out


def fused_optimizer_step(weights, grad, optimizer="exact_sgd", lr=0.01, eps=1.0e-8, state=None,
                         touched=None):
    """The "exact" fused optimizers of fbgemm's SplitTableBatchedEmbeddingBagsCodegen as the reference
    builds it (train/comms/pt/comms_utils.py:2015 optimizer=OptimType.EXACT_ROWWISE_ADAGRAD;
    train/compute/python/workloads/pytorch/split_table_batched_embeddings_ops.py:279-301 lr / eps):
    `grad` is the dense-equivalent gradient [rows, dim] of ONE step (tbe_bwd above), i.e. the sum over
    all lookups of a row; every touched row gets ONE update.

      exact_sgd              : w -= lr * g
      exact_row_wise_adagrad : m[row] += mean_d(g[row, d]^2);  w[row] -= lr / (sqrt(m[row]) + eps) * g[row]

    fbgemm_gpu is not in the reference tree: the formulas restate its published
    optimizers; tests/test_oracle_golden.py pins the Adagrad arithmetic (sqrt / eps placement, state
    accumulation over steps) against torch.optim.Adagrad on gradients that are constant along a row
    (the reference's own create_grad = ones_like, :315-316), where rowwise == elementwise.
    Rows outside `touched` (bool [rows]; default: rows with a non-zero gradient are irrelevant because
    an all-zero gradient leaves both w and m unchanged) are left alone.
    Returns (new_weights float64, new_state float64 or None)."""
    w = np.ascontiguousarray(weights, dtype=np.float64).copy()
    g = np.ascontiguousarray(grad, dtype=np.float64)
    if touched is not None:
        g = np.ascontiguousarray(g * np.asarray(touched, dtype=np.float64)[:, None])
    if optimizer in ("sgd", "exact_sgd"):
        code, m = 1, None
    elif optimizer in ("rowwise_adagrad", "exact_row_wise_adagrad", "exact_rowwise_adagrad"):
        code = 2
        m = np.zeros(w.shape[0]) if state is None else np.ascontiguousarray(state, dtype=np.float64).copy()
    else:
        raise ValueError(f"unknown optimizer {optimizer}")
    _load().oracle_fused_optimizer_step(_p(w, C.c_double), _p(m, C.c_double), _p(g, C.c_double),
                                        C.c_int64(w.shape[0]), C.c_int32(w.shape[1]), C.c_int32(code),
                                        C.c_double(lr), C.c_double(eps))
    return w, m
