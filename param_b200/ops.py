"""Torch-tensor host layer over the C ABI (include/param_b200.h).

PyTorch is used here for device memory, streams and dtype bookkeeping only: every function
passes raw device pointers + the current CUDA stream to libparam_b200.so.  Nothing in this
module computes on the CPU or falls back to ATen kernels; a non-CUDA tensor is an error.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _cabi
from ._cabi import (BWD_ATOMIC, BWD_AUTO, BWD_EXACT, BWD_SORTED, FWD_AUTO, FWD_DIRECT,  # noqa: F401
                    FWD_STAGED, IDX_I32, IDX_I64, OPT_ROWWISE_ADAGRAD, OPT_SGD, POOL_MEAN, POOL_SUM,
                    W_F16, W_F32, PB200Error)

_MODE = {"sum": POOL_SUM, "mean": POOL_MEAN}
_FWD_ALGO = {"auto": FWD_AUTO, "direct": FWD_DIRECT, "staged": FWD_STAGED, "pipelined": _cabi.FWD_PIPELINED}
_BWD_ALGO = {"auto": BWD_AUTO, "atomic": BWD_ATOMIC, "sorted": BWD_SORTED, "exact": BWD_EXACT}
# fbgemm OptimType values as they appear in the reference's op configs
# (split_table_batched_embeddings_ops.py:290 `OptimType(optimizer)`)
_OPTIMIZER = {"sgd": OPT_SGD, "exact_sgd": OPT_SGD,
              "rowwise_adagrad": OPT_ROWWISE_ADAGRAD, "exact_row_wise_adagrad": OPT_ROWWISE_ADAGRAD,
              "exact_rowwise_adagrad": OPT_ROWWISE_ADAGRAD}
_W_TYPE = {torch.float32: W_F32, torch.float16: W_F16}


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise PB200Error(
                "param_b200 kernels run on CUDA tensors only (no CPU fallback); got a "
                f"{t.device} tensor"
            )


def _idx_type(indices: torch.Tensor, offsets: torch.Tensor) -> int:
    if indices.dtype != offsets.dtype:
        raise PB200Error("indices and offsets must have the same integer dtype")
    if indices.dtype == torch.int64:
        return IDX_I64
    if indices.dtype == torch.int32:
        return IDX_I32
    raise PB200Error(f"unsupported index dtype {indices.dtype}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# --------------------------------------------------------------------------------------
# single-table EmbeddingBag forward (nn.EmbeddingBag call contract)
# --------------------------------------------------------------------------------------
def embedding_bag_forward(weight: torch.Tensor, indices: torch.Tensor, offsets: torch.Tensor,
                          mode: str = "sum", per_sample_weights: Optional[torch.Tensor] = None,
                          include_last_offset: bool = False, algo: str = "auto",
                          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b] = pool_{i in bag b} weight[indices[i]]  — reference call site
    train/compute/pt/pytorch_emb.py:61 / train/comms/pt/dlrm.py:380."""
    _need_cuda(weight, indices, offsets, per_sample_weights, out)
    if weight.dtype != torch.float32 or weight.dim() != 2 or not weight.is_contiguous():
        raise PB200Error("weight must be a contiguous fp32 [rows, dim] tensor")
    indices = indices.contiguous().view(-1)
    offsets = offsets.contiguous().view(-1)
    it = _idx_type(indices, offsets)
    n_bags = offsets.numel() - (1 if include_last_offset else 0)
    if n_bags < 0:
        raise PB200Error("include_last_offset requires at least one offset")
    rows, dim = weight.shape
    if out is None:
        out = torch.empty((n_bags, dim), dtype=torch.float32, device=weight.device)
    elif out.shape != (n_bags, dim) or out.dtype != torch.float32 or out.stride(1) != 1:
        raise PB200Error("out must be fp32 [n_bags, dim] with unit inner stride")
    psw = None
    if per_sample_weights is not None:
        psw = per_sample_weights.contiguous().view(-1).to(torch.float32)
        if psw.numel() != indices.numel():
            raise PB200Error("per_sample_weights must match indices")
    if n_bags == 0:
        return out
    rc = _cabi.load().pb200_embbag_fwd(
        weight.data_ptr(), rows, dim, _ptr(indices), indices.numel(), _ptr(offsets), n_bags,
        1 if include_last_offset else 0, it, _ptr(psw), _MODE[mode], out.data_ptr(),
        out.stride(0) if n_bags > 0 else dim, _FWD_ALGO[algo], _stream_ptr(weight))
    _cabi.check(rc, "pb200_embbag_fwd")
    return out


# --------------------------------------------------------------------------------------
# batched multi-table (TBE layout)
# --------------------------------------------------------------------------------------
@dataclass
class TableArena:
    """T embedding tables in one arena [sum(rows), dim] (fp32, or fp16 for weights_precision=fp16);
    table t = rows [row_offsets[t], row_offsets[t+1]).  Mirrors the storage of fbgemm's
    SplitTableBatchedEmbeddingBagsCodegen (weights_offsets) that
    train/comms/pt/comms_utils.py:1995-2017 constructs, with a uniform dim."""
    weights: torch.Tensor            # fp32 / fp16 [total_rows, dim], CUDA
    row_offsets: torch.Tensor        # int64 [T+1], CUDA
    rows: Sequence[int]              # host copy of per-table row counts
    dim: int

    @property
    def num_tables(self) -> int:
        return len(self.rows)

    @property
    def total_rows(self) -> int:
        return int(self.weights.shape[0])

    def table(self, t: int) -> torch.Tensor:
        lo = sum(self.rows[:t])
        return self.weights[lo:lo + self.rows[t]]

    @staticmethod
    def allocate(rows: Sequence[int], dim: int, device, dtype=torch.float32) -> "TableArena":
        total = int(sum(rows))
        if total >= 2 ** 32 - 1:
            raise PB200Error("arena row count must stay below 2^32 - 1")
        if dtype not in _W_TYPE:
            raise PB200Error("tables are fp32 or fp16")
        w = torch.empty((total, dim), dtype=dtype, device=device)
        ro = torch.tensor([0] + list(torch.tensor(list(rows), dtype=torch.int64).cumsum(0).tolist()),
                          dtype=torch.int64, device=device)
        return TableArena(w, ro, list(int(r) for r in rows), int(dim))


def tbe_forward(arena: TableArena, indices: torch.Tensor, offsets: torch.Tensor, batch: int,
                mode: str = "sum", per_sample_weights: Optional[torch.Tensor] = None,
                layout: str = "BTD", algo: str = "auto",
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Batched lookup over all tables of the arena.  offsets has T*batch + 1 entries (TBE request
    layout, split_table_batched_embeddings_ops.py:93-135).  layout "BTD" -> out [batch, T*dim]
    (TBE / all-to-all-ready), "TBD" -> out [T, batch, dim] (dlrm.py:387 torch.stack layout)."""
    _need_cuda(arena.weights, indices, offsets, per_sample_weights, out)
    T, D = arena.num_tables, arena.dim
    indices = indices.contiguous().view(-1)
    offsets = offsets.contiguous().view(-1)
    it = _idx_type(indices, offsets)
    if offsets.numel() != T * batch + 1:
        raise PB200Error(f"offsets must have T*batch+1 = {T * batch + 1} entries, got {offsets.numel()}")
    if layout == "BTD":
        shape, st_t, st_b = (batch, T * D), D, T * D
    elif layout == "TBD":
        shape, st_t, st_b = (T, batch, D), batch * D, D
    else:
        raise PB200Error("layout must be 'BTD' or 'TBD'")
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=arena.weights.device)
    elif tuple(out.shape) != shape or not out.is_contiguous() or out.dtype != torch.float32:
        raise PB200Error(f"out must be contiguous fp32 {shape}")
    psw = None
    if per_sample_weights is not None:
        psw = per_sample_weights.contiguous().view(-1).to(torch.float32)
        if psw.numel() != indices.numel():
            raise PB200Error("per_sample_weights must match indices")
    if batch == 0:
        return out
    if arena.weights.dtype == torch.float16:
        if algo not in ("auto", "direct"):
            raise PB200Error("fp16 tables run on the DIRECT forward variant only")
        rc = _cabi.load().pb200_tbe_fwd_f16(
            arena.weights.data_ptr(), arena.row_offsets.data_ptr(), T, D, _ptr(indices),
            indices.numel(), _ptr(offsets), batch, it, _ptr(psw), _MODE[mode], out.data_ptr(),
            st_t, st_b, _stream_ptr(arena.weights))
        _cabi.check(rc, "pb200_tbe_fwd_f16")
        return out
    if arena.weights.dtype != torch.float32:
        raise PB200Error("tables are fp32 or fp16")
    rc = _cabi.load().pb200_tbe_fwd(
        arena.weights.data_ptr(), arena.row_offsets.data_ptr(), T, D, _ptr(indices),
        indices.numel(), _ptr(offsets), batch, it, _ptr(psw), _MODE[mode], out.data_ptr(),
        st_t, st_b, _FWD_ALGO[algo], _stream_ptr(arena.weights))
    _cabi.check(rc, "pb200_tbe_fwd")
    return out


_scratch_cache: dict = {}


def _scratch(device, nbytes: int) -> torch.Tensor:
    """Grow-only scratch buffer, one per (device, current stream): kernels of two streams never share
    one, and the caching allocator hands a replaced buffer back only in the order of the stream it was
    allocated on."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _scratch_cache[key] = buf
    return buf


def _layout_strides(layout: str, T: int, D: int, batch: int):
    if layout == "BTD":
        return D, T * D
    if layout == "TBD":
        return batch * D, D
    raise PB200Error("layout must be 'BTD' or 'TBD'")


@dataclass
class SortPlan:
    """The index-only half of the sort-based backward (C ABI section 3c), built ahead of the gradient —
    typically on a side stream while the forward lookup of the same request runs.  `buf` is the scratch
    buffer the backward call consumes with plan_ready = 1; `ready` is recorded on the building stream."""
    buf: torch.Tensor
    signature: tuple
    ready: Optional[torch.cuda.Event]
    exact: bool


def _plan_signature(indices, offsets, T, D, batch, layout, mode, psw):
    return (indices.data_ptr(), indices.numel(), offsets.data_ptr(), T, D, batch, layout, mode,
            None if psw is None else psw.data_ptr())


def tbe_plan(row_offsets: torch.Tensor, num_tables: int, dim: int, indices: torch.Tensor,
             offsets: torch.Tensor, batch: int, max_table_rows: int, layout: str = "BTD",
             mode: str = "sum", per_sample_weights: Optional[torch.Tensor] = None,
             exact: bool = False, stream: Optional[torch.cuda.Stream] = None,
             buf: Optional[torch.Tensor] = None) -> SortPlan:
    """Sort the lookups of a TBE request by row (pb200_tbe_plan_build).  With `stream` the sort is
    queued on that stream after everything already queued on the current one (inputs are ready
    there), so it overlaps the forward lookup; tbe_backward(..., plan=) then waits for it.
    `exact` sizes the buffer for tbe_backward_fused / algo="exact".  `buf`: reuse a buffer of a
    previous plan of the same shape (steady-state training loops)."""
    _need_cuda(row_offsets, indices, offsets, per_sample_weights)
    indices = indices.contiguous().view(-1)
    offsets = offsets.contiguous().view(-1)
    it = _idx_type(indices, offsets)
    T, D = num_tables, dim
    if offsets.numel() != T * batch + 1:
        raise PB200Error("offsets must have T*batch+1 entries")
    st_t, st_b = _layout_strides(layout, T, D, batch)
    psw = None
    if per_sample_weights is not None:
        psw = per_sample_weights.contiguous().view(-1)
        if psw.dtype != torch.float32 or psw.numel() != indices.numel():
            raise PB200Error("per_sample_weights must be fp32 and match indices")
    lib = _cabi.load()
    if exact:
        nbytes = int(lib.pb200_tbe_bwd_fused_scratch_bytes(indices.numel(), T, batch, D))
    else:
        nbytes = int(lib.pb200_tbe_bwd_scratch_bytes(indices.numel(), T, batch, 0, BWD_SORTED))
    dev = indices.device
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    cur = torch.cuda.current_stream(dev)
    run_on = stream if stream is not None else cur
    if stream is not None:
        stream.wait_stream(cur)
        for t in (buf, indices, offsets, row_offsets, psw):
            if t is not None:
                t.record_stream(stream)
    rc = lib.pb200_tbe_plan_build(buf.data_ptr(), nbytes, row_offsets.data_ptr(), int(max_table_rows), T, D,
                                  _ptr(indices), indices.numel(), _ptr(offsets), batch, it, _ptr(psw),
                                  _MODE[mode], st_t, st_b, run_on.cuda_stream)
    _cabi.check(rc, "pb200_tbe_plan_build")
    ready = None
    if stream is not None:
        ready = torch.cuda.Event()
        ready.record(stream)
    return SortPlan(buf, _plan_signature(indices, offsets, T, D, batch, layout, mode, psw), ready, exact)


def _use_plan(plan: SortPlan, sig: tuple, need_bytes: int, dev) -> int:
    if plan.signature != sig:
        raise PB200Error("the sort plan was built for another request (indices / offsets / layout / mode / "
                         "per_sample_weights differ)")
    if plan.buf.numel() < need_bytes:
        raise PB200Error("the sort plan buffer is too small for this backward variant (build it with exact=True)")
    if plan.ready is not None:
        torch.cuda.current_stream(dev).wait_event(plan.ready)
    return plan.buf.data_ptr()


def tbe_backward(dst: torch.Tensor, row_offsets: torch.Tensor, num_tables: int, dim: int,
                 indices: torch.Tensor, offsets: torch.Tensor, batch: int,
                 grad_out: torch.Tensor, layout: str = "BTD", scale: float = 1.0,
                 mode: str = "sum", per_sample_weights: Optional[torch.Tensor] = None,
                 algo: str = "auto", max_table_rows: int = 0,
                 plan: Optional[SortPlan] = None) -> None:
    """dst[row(t, idx[i])] += scale * w_i * grad_out[(t, bag(i))].  dst is either a dense grad
    buffer shaped like the arena (scale=1) or the arena itself (scale=-lr, fused SGD).
    Reference: autograd backward driven at pytorch_dist_backend.py:849-857 / dlrm.py:1296.
    max_table_rows: rows of the largest table (fixes the radix passes of the sort; 0 = the arena's
    row count, always a valid bound).  plan: a SortPlan built by tbe_plan for this request — the call
    is then the segmented reduce alone."""
    _need_cuda(dst, row_offsets, indices, offsets, grad_out, per_sample_weights)
    indices = indices.contiguous().view(-1)
    offsets = offsets.contiguous().view(-1)
    it = _idx_type(indices, offsets)
    T, D = num_tables, dim
    if offsets.numel() != T * batch + 1:
        raise PB200Error("offsets must have T*batch+1 entries")
    grad_out = grad_out.contiguous()
    if grad_out.dtype != torch.float32 or grad_out.numel() != T * batch * D:
        raise PB200Error("grad_out must be fp32 with T*batch*dim elements")
    st_t, st_b = _layout_strides(layout, T, D, batch)
    if dst.dtype != torch.float32:
        raise PB200Error("tbe_backward scatters into fp32 rows; fp16 tables use tbe_backward_fused")
    if dst.dim() != 2 or dst.shape[1] != D or not dst.is_contiguous():
        raise PB200Error("dst must be a contiguous fp32 [rows, dim] tensor")
    psw = None
    if per_sample_weights is not None:
        psw = per_sample_weights.contiguous().view(-1).to(torch.float32)
        if psw.numel() != indices.numel():
            raise PB200Error("per_sample_weights must match indices")
    lib = _cabi.load()
    a = _BWD_ALGO[algo]
    if a == BWD_AUTO:
        # a plan says which reducer it was sized for; otherwise the sort pays from ~64 k lookups on
        # (below that the launch latency of the sort passes exceeds what the atomics cost)
        if plan is not None:
            a = BWD_EXACT if plan.exact else BWD_SORTED
        else:
            a = BWD_SORTED if indices.numel() >= 65536 else BWD_ATOMIC
    scratch_ptr, scratch_bytes = None, 0
    if a in (BWD_SORTED, BWD_AUTO):
        scratch_bytes = int(lib.pb200_tbe_bwd_scratch_bytes(indices.numel(), T, batch, dst.shape[0], BWD_SORTED))
    elif a == BWD_EXACT:
        scratch_bytes = int(lib.pb200_tbe_bwd_fused_scratch_bytes(indices.numel(), T, batch, D))
    if plan is not None:
        if a == BWD_ATOMIC:
            raise PB200Error("the atomic backward takes no sort plan")
        scratch_ptr = _use_plan(plan, _plan_signature(indices, offsets, T, D, batch, layout, mode, psw),
                                scratch_bytes, dst.device)
    elif scratch_bytes:
        scratch_ptr = _scratch(dst.device, scratch_bytes).data_ptr()
    mtr = int(max_table_rows) if max_table_rows and max_table_rows > 0 else int(dst.shape[0])
    rc = lib.pb200_tbe_bwd(dst.data_ptr(), row_offsets.data_ptr(), T, D, _ptr(indices),
                           indices.numel(), _ptr(offsets), batch, it, _ptr(psw), _MODE[mode],
                           grad_out.data_ptr(), st_t, st_b, float(scale), a, mtr, scratch_ptr,
                           scratch_bytes, 1 if plan is not None else 0, _stream_ptr(dst))
    _cabi.check(rc, "pb200_tbe_bwd")


def part_range(n: int, part: int, parts: int):
    """tables [lo, hi) of piece `part` when n tables are cut into `parts` contiguous pieces (remainder to the low
    pieces) — the split pb200_a2a_pooled_bwd_part uses"""
    k, m = divmod(int(n), int(parts))
    lo = part * k + min(part, m)
    return lo, lo + k + (1 if part < m else 0)


def tbe_backward_tables(dst: torch.Tensor, row_offsets: torch.Tensor, num_tables: int, dim: int,
                        indices: torch.Tensor, offsets: torch.Tensor, batch: int, grad_out: torch.Tensor,
                        plan: SortPlan, table_lo: int, table_hi: int, layout: str = "BTD", scale: float = 1.0,
                        mode: str = "sum", per_sample_weights: Optional[torch.Tensor] = None) -> None:
    """The segmented reduce of tbe_backward(algo="sorted", plan=plan) for tables [table_lo, table_hi) only
    (pb200_tbe_bwd_tables): the backward of a table group can start when ITS gradient columns have arrived."""
    _need_cuda(dst, row_offsets, indices, offsets, grad_out, per_sample_weights)
    indices = indices.contiguous().view(-1)
    offsets = offsets.contiguous().view(-1)
    it = _idx_type(indices, offsets)
    T, D = num_tables, dim
    if offsets.numel() != T * batch + 1:
        raise PB200Error("offsets must have T*batch+1 entries")
    if grad_out.dtype != torch.float32 or not grad_out.is_contiguous() or grad_out.numel() != T * batch * D:
        raise PB200Error("grad_out must be contiguous fp32 with T*batch*dim elements")
    if dst.dtype != torch.float32 or dst.dim() != 2 or dst.shape[1] != D or not dst.is_contiguous():
        raise PB200Error("dst must be a contiguous fp32 [rows, dim] tensor")
    if plan.exact:
        raise PB200Error("table-range backward runs on a SORTED plan (exact=False)")
    st_t, st_b = _layout_strides(layout, T, D, batch)
    psw = None
    if per_sample_weights is not None:
        psw = per_sample_weights.contiguous().view(-1).to(torch.float32)
    lib = _cabi.load()
    nbytes = int(lib.pb200_tbe_bwd_scratch_bytes(indices.numel(), T, batch, dst.shape[0], BWD_SORTED))
    ptr = _use_plan(plan, _plan_signature(indices, offsets, T, D, batch, layout, mode, psw), nbytes, dst.device)
    rc = lib.pb200_tbe_bwd_tables(dst.data_ptr(), row_offsets.data_ptr(), T, D, _ptr(indices), indices.numel(),
                                  _ptr(offsets), batch, it, _ptr(psw), _MODE[mode], grad_out.data_ptr(), st_t, st_b,
                                  float(scale), int(table_lo), int(table_hi), ptr, nbytes, _stream_ptr(dst))
    _cabi.check(rc, "pb200_tbe_bwd_tables")


def tbe_backward_fused(weights: torch.Tensor, row_offsets: torch.Tensor, num_tables: int, dim: int,
                       indices: torch.Tensor, offsets: torch.Tensor, batch: int,
                       grad_out: torch.Tensor, optimizer: str = "exact_sgd", lr: float = 0.01,
                       eps: float = 1.0e-8, state: Optional[torch.Tensor] = None,
                       layout: str = "BTD", mode: str = "sum",
                       per_sample_weights: Optional[torch.Tensor] = None,
                       stochastic_rounding: bool = False, sr_seed: int = 0,
                       max_table_rows: int = 0, plan: Optional[SortPlan] = None) -> None:
    """Backward with the optimizer fused in, one deterministic update per touched row (C ABI §3b):
    exact_sgd `w -= lr*g` or exact_row_wise_adagrad `m += mean(g^2); w -= lr/(sqrt(m)+eps)*g`,
    fp32 or fp16 tables.  Stands in for the fused backward of fbgemm's TBE op that the reference
    builds at comms_utils.py:1995-2017 / split_table_batched_embeddings_ops.py:279-301.
    max_table_rows / plan: as for tbe_backward (the plan must have been built with exact=True)."""
    _need_cuda(weights, row_offsets, indices, offsets, grad_out, per_sample_weights, state)
    if weights.dtype not in _W_TYPE or weights.dim() != 2 or not weights.is_contiguous():
        raise PB200Error("weights must be a contiguous fp32 or fp16 [rows, dim] tensor")
    if optimizer not in _OPTIMIZER:
        raise PB200Error(f"optimizer must be one of {sorted(_OPTIMIZER)}")
    opt = _OPTIMIZER[optimizer]
    if opt == OPT_ROWWISE_ADAGRAD:
        if state is None or state.dtype != torch.float32 or state.numel() != weights.shape[0] \
                or not state.is_contiguous():
            raise PB200Error("rowwise Adagrad needs a contiguous fp32 state with one element per arena row")
    indices = indices.contiguous().view(-1)
    offsets = offsets.contiguous().view(-1)
    it = _idx_type(indices, offsets)
    T, D = num_tables, dim
    if weights.shape[1] != D:
        raise PB200Error("weights must be [rows, dim]")
    if offsets.numel() != T * batch + 1:
        raise PB200Error("offsets must have T*batch+1 entries")
    grad_out = grad_out.contiguous()
    if grad_out.dtype != torch.float32 or grad_out.numel() != T * batch * D:
        raise PB200Error("grad_out must be fp32 with T*batch*dim elements")
    st_t, st_b = _layout_strides(layout, T, D, batch)
    psw = None
    if per_sample_weights is not None:
        psw = per_sample_weights.contiguous().view(-1).to(torch.float32)
        if psw.numel() != indices.numel():
            raise PB200Error("per_sample_weights must match indices")
    lib = _cabi.load()
    sb = int(lib.pb200_tbe_bwd_fused_scratch_bytes(indices.numel(), T, batch, D))
    if plan is not None:
        scratch_ptr = _use_plan(plan, _plan_signature(indices, offsets, T, D, batch, layout, mode, psw), sb,
                                weights.device)
    else:
        scratch_ptr = _scratch(weights.device, sb).data_ptr()
    mtr = int(max_table_rows) if max_table_rows and max_table_rows > 0 else int(weights.shape[0])
    rc = lib.pb200_tbe_bwd_fused(weights.data_ptr(), _W_TYPE[weights.dtype], _ptr(state),
                                 row_offsets.data_ptr(), T, D, _ptr(indices), indices.numel(),
                                 _ptr(offsets), batch, it, _ptr(psw), _MODE[mode],
                                 grad_out.data_ptr(), st_t, st_b, opt, float(lr), float(eps),
                                 1 if stochastic_rounding else 0,
                                 C.c_uint64(int(sr_seed) & (2 ** 64 - 1)), mtr, scratch_ptr, sb,
                                 1 if plan is not None else 0, _stream_ptr(weights))
    _cabi.check(rc, "pb200_tbe_bwd_fused")


_row_offsets_cache: dict = {}


def single_table_row_offsets(num_rows: int, device) -> torch.Tensor:
    """int64 [0, num_rows] on `device`, cached: building it from a Python list on every backward is a
    synchronising H2D copy (it cost the aten override 370 us per op under et_replay, profiles/r02l_*)."""
    key = (int(num_rows), device.type, device.index)
    t = _row_offsets_cache.get(key)
    if t is None:
        t = torch.tensor([0, int(num_rows)], dtype=torch.int64, device=device)
        _row_offsets_cache[key] = t
    return t


def close_offsets(offsets: torch.Tensor, n_indices: int) -> torch.Tensor:
    """nn.EmbeddingBag offsets (one per bag) -> the closed form with the trailing end offset, built on the device
    (new_full is a fill kernel: no host-to-device copy, no synchronisation)"""
    offsets = offsets.view(-1)
    return torch.cat([offsets, offsets.new_full((1,), int(n_indices))])


def embedding_bag_backward_sparse(grad_out: torch.Tensor, indices: torch.Tensor, offsets: torch.Tensor,
                                  num_rows: int, mode: str = "sum",
                                  per_sample_weights: Optional[torch.Tensor] = None,
                                  include_last_offset: bool = False) -> torch.Tensor:
    """Gradient of a single-table EmbeddingBag as the uncoalesced sparse COO tensor nn.EmbeddingBag(sparse=True)
    produces: indices = the lookup indices, values[i] = w_i * grad_out[bag(i)] (pb200_embbag_bwd_sparse).  No
    dense [rows, dim] buffer is allocated."""
    _need_cuda(grad_out, indices, offsets, per_sample_weights)
    indices = indices.contiguous().view(-1)
    offsets = offsets.contiguous().view(-1)
    it = _idx_type(indices, offsets)
    if grad_out.dtype != torch.float32 or grad_out.dim() != 2 or grad_out.stride(1) != 1:
        raise PB200Error("grad_out must be fp32 [n_bags, dim] with unit inner stride")
    n_bags = offsets.numel() - (1 if include_last_offset else 0)
    if grad_out.shape[0] != n_bags:
        raise PB200Error("grad_out must have one row per bag")
    dim = int(grad_out.shape[1])
    psw = None
    if per_sample_weights is not None:
        psw = per_sample_weights.contiguous().view(-1).to(torch.float32)
        if psw.numel() != indices.numel():
            raise PB200Error("per_sample_weights must match indices")
    values = torch.empty((indices.numel(), dim), dtype=torch.float32, device=grad_out.device)
    if indices.numel():
        rc = _cabi.load().pb200_embbag_bwd_sparse(grad_out.data_ptr(), grad_out.stride(0), dim, _ptr(offsets), n_bags,
                                                  1 if include_last_offset else 0, indices.numel(), it, _ptr(psw),
                                                  _MODE[mode], values.data_ptr(), _stream_ptr(grad_out))
        _cabi.check(rc, "pb200_embbag_bwd_sparse")
    return torch.sparse_coo_tensor(indices.to(torch.int64).view(1, -1), values, size=(int(num_rows), dim))


def check_indices(row_offsets: torch.Tensor, num_tables: int, indices: torch.Tensor,
                  offsets: torch.Tensor, batch: int) -> int:
    """Debug-mode bounds check; returns the number of out-of-range lookups (synchronises)."""
    _need_cuda(row_offsets, indices, offsets)
    indices = indices.contiguous().view(-1)
    offsets = offsets.contiguous().view(-1)
    bad = torch.zeros(1, dtype=torch.int64, device=indices.device)
    rc = _cabi.load().pb200_check_indices(row_offsets.data_ptr(), num_tables, _ptr(indices),
                                          indices.numel(), _ptr(offsets), batch,
                                          _idx_type(indices, offsets), bad.data_ptr(),
                                          _stream_ptr(indices))
    _cabi.check(rc, "pb200_check_indices")
    return int(bad.item())


# --------------------------------------------------------------------------------------
# sparse-input regroup (dlrm.py:430-504)
# --------------------------------------------------------------------------------------
def regroup_sparse(lengths: torch.Tensor, indices: torch.Tensor, world: int, tables_local: int,
                   local_batch: int):
    """lengths [W][T_l][b] + indices in the same order  ->  (lengths_out [T_l, W*b],
    offsets_out [T_l*W*b+1], indices_out table-major).  int64, bit-exact."""
    _need_cuda(lengths, indices)
    lengths = lengths.contiguous().view(-1)
    indices = indices.contiguous().view(-1)
    if lengths.dtype != torch.int64 or indices.dtype != torch.int64:
        raise PB200Error("regroup_sparse takes int64 lengths and indices")
    n = world * tables_local * local_batch
    if lengths.numel() != n:
        raise PB200Error("lengths must have W*T_l*b elements")
    dev = lengths.device
    lengths_out = torch.empty(n, dtype=torch.int64, device=dev)
    offsets_out = torch.empty(n + 1, dtype=torch.int64, device=dev)
    indices_out = torch.empty_like(indices)
    lib = _cabi.load()
    sb = int(lib.pb200_regroup_scratch_bytes(world, tables_local, local_batch))
    scratch = _scratch(dev, sb)
    rc = lib.pb200_regroup_sparse(lengths.data_ptr(), _ptr(indices), indices.numel(), world,
                                  tables_local, local_batch, lengths_out.data_ptr(),
                                  offsets_out.data_ptr(), _ptr(indices_out), scratch.data_ptr(),
                                  sb, _stream_ptr(lengths))
    _cabi.check(rc, "pb200_regroup_sparse")
    return lengths_out.view(tables_local, world * local_batch), offsets_out, indices_out


# --------------------------------------------------------------------------------------
# per-block sums of pooled vectors (the scalar result of a step)
# --------------------------------------------------------------------------------------
def pooled_sum(pooled: torch.Tensor, n_blocks: int = 1) -> torch.Tensor:
    """float64 [n_blocks]: the sum of each of the n_blocks equal contiguous blocks of `pooled` (fp32), accumulated
    in double in a fixed order (pb200_pooled_sum).  n_blocks = tables for a [T, B, D] tensor gives per-table
    sums; n_blocks = rows for a batch-major [B, T*D] tensor gives per-sample sums."""
    _need_cuda(pooled)
    if pooled.dtype != torch.float32 or not pooled.is_contiguous():
        raise PB200Error("pooled_sum takes a contiguous fp32 tensor")
    n_blocks = int(n_blocks)
    if n_blocks < 1 or pooled.numel() % n_blocks != 0 or pooled.numel() == 0:
        raise PB200Error("pooled_sum: the tensor does not split into n_blocks equal blocks")
    lib = _cabi.load()
    out = torch.empty(n_blocks, dtype=torch.float64, device=pooled.device)
    nbytes = int(lib.pb200_pooled_sum_scratch_bytes(n_blocks))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=pooled.device)
    rc = lib.pb200_pooled_sum(pooled.data_ptr(), n_blocks, pooled.numel() // n_blocks, out.data_ptr(),
                              scratch.data_ptr(), nbytes, _stream_ptr(pooled))
    _cabi.check(rc, "pb200_pooled_sum")
    return out


# --------------------------------------------------------------------------------------
# synthetic data on the device
# --------------------------------------------------------------------------------------
def fill_uniform_(t: torch.Tensor, lo: float, hi: float, seed: int) -> torch.Tensor:
    _need_cuda(t)
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise PB200Error("fill_uniform_ takes a contiguous fp32 tensor")
    rc = _cabi.load().pb200_fill_uniform(t.data_ptr(), t.numel(), float(lo), float(hi),
                                         C.c_uint64(seed & (2 ** 64 - 1)), _stream_ptr(t))
    _cabi.check(rc, "pb200_fill_uniform")
    return t


def fill_zipf_indices_(t: torch.Tensor, nnz: int, cdf: torch.Tensor, seed: int,
                       dedupe: bool = True) -> torch.Tensor:
    """t: int64 [n_bags * nnz]; bag-wise truncated-Zipf draws (distinct inside a bag if dedupe)."""
    _need_cuda(t, cdf)
    if t.dtype != torch.int64 or cdf.dtype != torch.float64 or not t.is_contiguous():
        raise PB200Error("fill_zipf_indices_ takes a contiguous int64 destination and a float64 cdf")
    if t.numel() % nnz != 0:
        raise PB200Error("destination size must be a multiple of nnz")
    rc = _cabi.load().pb200_fill_zipf_indices(t.data_ptr(), t.numel() // nnz, nnz, cdf.data_ptr(),
                                              cdf.numel(), 1 if dedupe else 0,
                                              C.c_uint64(seed & (2 ** 64 - 1)), _stream_ptr(t))
    _cabi.check(rc, "pb200_fill_zipf_indices")
    return t
