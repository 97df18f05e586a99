"""Batched multi-table EmbeddingBag module on the B200 kernels.

Call-compatible with how PARAM drives fbgemm's SplitTableBatchedEmbeddingBagsCodegen
(train/compute/python/workloads/pytorch/split_table_batched_embeddings_ops.py:248-324 and
train/comms/pt/comms_utils.py:1995-2017, pytorch_dist_backend.py:832-857):

    op = B200TBE([(rows, dim)] * T, lr=0.01)                 # embedding_specs
    out = op.forward(indices, offsets, per_sample_weights)    # [B, T*dim], TBE request layout
    out.backward(grad)                                        # fused optimizer step (SGD) in the arena

The optimizer is fused into the backward like fbgemm's (there: EXACT_ROWWISE_ADAGRAD by default; here
plain SGD `W -= lr * dW` — rowwise Adagrad is a §8f "next" item).  Arithmetic parity at this boundary
is pinned through the per-table nn.EmbeddingBag loop (fbgemm_gpu itself is absent: parity unpinned).
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
        out = ops.tbe_forward(op.arena, indices, offsets, B, mode=op.pooling_mode,
                              per_sample_weights=psw, layout="BTD", algo=op.fwd_algo)
        ctx.op, ctx.B = op, B
        ctx.save_for_backward(indices, offsets, psw if psw is not None else torch.empty(0))
        ctx.weighted = psw is not None
        return out

    @staticmethod
    def backward(ctx, grad):
        indices, offsets, psw = ctx.saved_tensors
        op = ctx.op
        ops.tbe_backward(op.arena.weights, op.arena.row_offsets, op.arena.num_tables, op.arena.dim,
                         indices, offsets, ctx.B, grad.contiguous(), layout="BTD", scale=-op.lr,
                         mode=op.pooling_mode, per_sample_weights=psw if ctx.weighted else None,
                         algo=op.bwd_algo)
        return None, None, None, None, None


class B200TBE(nn.Module):
    def __init__(self, embedding_specs: Sequence[Tuple[int, int]], lr: float = 0.01,
                 pooling_mode: str = "sum", device=None, fwd_algo: str = "auto",
                 bwd_algo: str = "sorted", seed: int = 0) -> None:
        super().__init__()
        dims = {int(d) for _, d in embedding_specs}
        if len(dims) != 1:
            raise PB200Error("B200TBE needs one embedding dim for all tables (mixed dims: not yet)")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.embedding_specs = [(int(r), int(d)) for r, d in embedding_specs]
        self.arena = ops.TableArena.allocate([r for r, _ in self.embedding_specs], dims.pop(), device)
        for t, (r, _) in enumerate(self.embedding_specs):
            ops.fill_uniform_(self.arena.table(t), -(1.0 / r) ** 0.5, (1.0 / r) ** 0.5, seed=seed * 65537 + t)
        self.lr, self.pooling_mode = float(lr), pooling_mode
        self.fwd_algo, self.bwd_algo = fwd_algo, bwd_algo
        # autograd needs one differentiable input to route the backward through
        self._anchor = nn.Parameter(torch.zeros(1, device=device))

    @property
    def weights(self) -> torch.Tensor:
        return self.arena.weights

    def forward(self, indices: torch.Tensor, offsets: torch.Tensor,
                per_sample_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
        return _TBEFn.apply(self._anchor, self, indices, offsets, per_sample_weights)
