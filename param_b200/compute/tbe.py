"""Batched multi-table EmbeddingBag module on the B200 kernels.

Call-compatible with how PARAM drives fbgemm's SplitTableBatchedEmbeddingBagsCodegen
(train/compute/python/workloads/pytorch/split_table_batched_embeddings_ops.py:248-324 and
train/comms/pt/comms_utils.py:1995-2017, pytorch_dist_backend.py:832-857):

    op = B200TBE([(rows, dim)] * T, optimizer="exact_row_wise_adagrad", learning_rate=0.01)
    out = op.forward(indices, offsets, per_sample_weights)    # [B, T*dim], TBE request layout
    out.backward(grad)                                        # fused optimizer step in the arena

The optimizer is fused into the backward like fbgemm's: `exact_sgd` (W -= lr*dW) or
`exact_row_wise_adagrad` (fbgemm's default and what comms_utils.py:2015 asks for), on fp32 or fp16
tables (`weights_precision`), fp16 optionally with stochastic rounding (the reference's op config
sets stochastic_rounding=True, split_table_batched_embeddings_ops.py:292).  "exact" = the gradient
of a row is summed over all its lookups of the batch, then ONE update is applied
(pb200_tbe_bwd_fused, deterministic).  `bwd_algo` "sorted"/"atomic" keep the scatter-add SGD
kernels of pb200_tbe_bwd for fp32 + SGD.  Arithmetic parity at this boundary is pinned through the
per-table nn.EmbeddingBag loop for the forward and SGD; fbgemm_gpu itself is absent, so the rowwise
Adagrad update follows its published formula and is checked against the oracle's restatement
(parity unpinned).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import ops
from .._cabi import PB200Error


class _TBEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, op, indices, offsets, psw):
        B = (offsets.numel() - 1) // op.arena.num_tables
        if psw is not None:      # one fp32 copy shared by the lookup, the sort plan and the backward
            psw = psw.contiguous().view(-1).float()
        indices, offsets = indices.contiguous().view(-1), offsets.contiguous().view(-1)
        out = ops.tbe_forward(op.arena, indices, offsets, B, mode=op.pooling_mode,
                              per_sample_weights=psw, layout="BTD", algo=op.fwd_algo)
        ctx.op, ctx.B = op, B
        ctx.save_for_backward(indices, offsets, psw if psw is not None else torch.empty(0))
        ctx.weighted = psw is not None
        ctx.plan = None
        if op.presort and op.bwd_algo != "atomic" and ctx.needs_input_grad[0] and indices.numel() >= 65536:
            # the sort of the backward depends on the request only: queue it beside the lookup
            if op._side is None:
                op._side = torch.cuda.Stream(device=out.device)
            ctx.plan = ops.tbe_plan(op.arena.row_offsets, op.arena.num_tables, op.arena.dim,
                                    indices, offsets, B, op.max_table_rows, layout="BTD", mode=op.pooling_mode,
                                    per_sample_weights=psw,
                                    exact=op.bwd_algo == "exact", stream=op._side)
        return out

    @staticmethod
    def backward(ctx, grad):
        indices, offsets, psw = ctx.saved_tensors
        op = ctx.op
        psw = psw if ctx.weighted else None
        if op.bwd_algo in ("sorted", "atomic"):
            ops.tbe_backward(op.arena.weights, op.arena.row_offsets, op.arena.num_tables, op.arena.dim,
                             indices, offsets, ctx.B, grad.contiguous(), layout="BTD", scale=-op.lr,
                             mode=op.pooling_mode, per_sample_weights=psw, algo=op.bwd_algo,
                             max_table_rows=op.max_table_rows, plan=ctx.plan)
        else:
            op.step += 1
            ops.tbe_backward_fused(op.arena.weights, op.arena.row_offsets, op.arena.num_tables,
                                   op.arena.dim, indices, offsets, ctx.B, grad.contiguous(),
                                   optimizer=op.optimizer, lr=op.lr, eps=op.eps, state=op.momentum1,
                                   layout="BTD", mode=op.pooling_mode, per_sample_weights=psw,
                                   stochastic_rounding=op.stochastic_rounding,
                                   sr_seed=op.seed * 0x9E3779B1 + op.step,
                                   max_table_rows=op.max_table_rows, plan=ctx.plan)
        return None, None, None, None, None


class B200TBE(nn.Module):
    def __init__(self, embedding_specs: Sequence[Tuple[int, int]], lr: Optional[float] = None,
                 pooling_mode: str = "sum", device=None, fwd_algo: str = "auto",
                 bwd_algo: Optional[str] = None, seed: int = 0, optimizer: str = "exact_sgd",
                 weights_precision: str = "fp32", learning_rate: float = 0.01, eps: float = 1.0e-8,
                 stochastic_rounding: bool = False) -> None:
        super().__init__()
        dims = {int(d) for _, d in embedding_specs}
        if len(dims) != 1:
            raise PB200Error("B200TBE needs one embedding dim for all tables (mixed dims: not yet)")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        dtype = {"fp32": torch.float32, "float32": torch.float32,
                 "fp16": torch.float16, "float16": torch.float16}.get(str(weights_precision).lower())
        if dtype is None:
            raise PB200Error("weights_precision must be fp32 or fp16")
        if optimizer not in ops._OPTIMIZER:
            raise PB200Error(f"optimizer must be one of {sorted(ops._OPTIMIZER)}")
        self.optimizer = optimizer
        adagrad = ops._OPTIMIZER[optimizer] == ops.OPT_ROWWISE_ADAGRAD
        if bwd_algo is None:   # scatter-add SGD kernels where they apply, the exact path otherwise
            bwd_algo = "exact" if (adagrad or dtype == torch.float16) else "sorted"
        if bwd_algo in ("sorted", "atomic") and (adagrad or dtype == torch.float16):
            raise PB200Error("bwd_algo sorted/atomic is fp32 + SGD only; use bwd_algo='exact'")
        self.embedding_specs = [(int(r), int(d)) for r, d in embedding_specs]
        self.arena = ops.TableArena.allocate([r for r, _ in self.embedding_specs], dims.pop(), device,
                                             dtype=dtype)
        tmp = None
        for t, (r, _) in enumerate(self.embedding_specs):
            lim = (1.0 / r) ** 0.5
            if dtype == torch.float32:
                ops.fill_uniform_(self.arena.table(t), -lim, lim, seed=seed * 65537 + t)
            else:      # initialisation only: generate in fp32, store rounded to fp16
                if tmp is None or tmp.shape[0] < r:
                    tmp = torch.empty((r, self.arena.dim), dtype=torch.float32, device=device)
                ops.fill_uniform_(tmp[:r], -lim, lim, seed=seed * 65537 + t)
                self.arena.table(t).copy_(tmp[:r])
        # rowwise Adagrad state: one fp32 per arena row (fbgemm's momentum1), zero-initialised
        self.momentum1 = torch.zeros(self.arena.total_rows, dtype=torch.float32, device=device) \
            if adagrad else None
        self.lr = float(learning_rate if lr is None else lr)
        self.eps, self.pooling_mode = float(eps), pooling_mode
        self.stochastic_rounding = bool(stochastic_rounding) and dtype == torch.float16
        self.seed, self.step = int(seed), 0
        self.fwd_algo, self.bwd_algo = fwd_algo, bwd_algo
        self.max_table_rows = max(r for r, _ in self.embedding_specs)
        # presort: build the backward's sort plan on a side stream during the forward (large requests)
        self.presort, self._side = True, None
        # autograd needs one differentiable input to route the backward through
        self._anchor = nn.Parameter(torch.zeros(1, device=device))

    @property
    def weights(self) -> torch.Tensor:
        return self.arena.weights

    def forward(self, indices: torch.Tensor, offsets: torch.Tensor,
                per_sample_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
        return _TBEFn.apply(self._anchor, self, indices, offsets, per_sample_weights)
