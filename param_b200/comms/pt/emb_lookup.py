"""`emb_lookup` compute kernel set-up without fbgemm_gpu.

Mirror of init_emb_lookup (train/comms/pt/comms_utils.py:1956-2039) and of the TBE request
generator it relies on (train/compute/python/workloads/pytorch/
split_table_batched_embeddings_ops.py:93-135, 191-213): same collectiveArgs fields are filled
(emb, embRequests, direction, emb_dim, batch_size, num_emb_tables_batched, num_emb_ops, LookupOut,
grad_output), so that `backendFuncs.emb_lookup(collectiveArgs)` (pytorch_dist_backend.py:832-857)
and the `"compute": "emb_lookup"` trace entries run on the B200 batched op.

Differences: the op is B200TBE with optimizer="exact_row_wise_adagrad" fused into the backward, as the
reference asks of fbgemm (OptimType.EXACT_ROWWISE_ADAGRAD, comms_utils.py:2015); a missing
`commsParams.direction` defaults to "forward" (the reference raises AttributeError there when driven
from commsComputeBench).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from ...compute.tbe import B200TBE

Request = Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]


def _table_indices(B: int, L: int, E: int, alpha: float, rng: np.random.Generator) -> torch.Tensor:
    """one table's B*L indices, the reference's four regimes keyed on alpha (:104-120)"""
    n = B * L
    if alpha == 0:
        return torch.arange(n, dtype=torch.int64) % L          # linear sequence by pooling factor
    if alpha <= 0.5:
        return torch.arange(n, dtype=torch.int64) % E          # linear sequence by embedding size
    if alpha <= 1.0:
        return torch.from_numpy(rng.integers(0, E, size=n, dtype=np.int64))
    return torch.from_numpy(rng.zipf(a=alpha, size=n).astype(np.int64) % E)   # folded unbounded Zipf


def generate_requests(iters: int, B: int, T: int, L: int, E: int, alpha: float = 1.0,
                      weighted: bool = False, seed: int = 0, device=None) -> List[Request]:
    """`iters` TBE requests: indices = cat over T tables (table-major), offsets int64[T*B + 1]
    cumulative over the concatenation (first table [0, L, ..., B*L], later tables continue),
    optional per-sample weights."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(iters):
        idx = torch.cat([_table_indices(B, L, E, alpha, rng) for _ in range(T)])
        offsets = torch.arange(T * B + 1, dtype=torch.int64) * L
        w = torch.randn(idx.numel(), dtype=torch.float32) if weighted else None
        if device is not None:
            idx, offsets = idx.to(device), offsets.to(device)
            w = w.to(device) if w is not None else None
        out.append((idx, offsets, w))
    return out


def init_emb_lookup(collectiveArgs, commsParams, backendFuncs) -> None:
    collectiveArgs.direction = getattr(commsParams, "direction", "forward")
    collectiveArgs.emb_dim = commsParams.emb_dim
    num_embeddings = commsParams.num_embs
    collectiveArgs.batch_size = commsParams.batch_size
    tables_per_device = commsParams.num_emb_tables_per_device
    collectiveArgs.num_emb_tables_batched = commsParams.num_emb_tables_batched
    batched = tables_per_device if collectiveArgs.num_emb_tables_batched == -1 \
        else collectiveArgs.num_emb_tables_batched
    collectiveArgs.num_emb_ops = tables_per_device // batched
    dev = backendFuncs.get_device()
    collectiveArgs.emb = [
        B200TBE([(num_embeddings, collectiveArgs.emb_dim)] * batched, device=dev,
                optimizer=getattr(commsParams, "emb_optimizer", "exact_row_wise_adagrad"),
                learning_rate=getattr(commsParams, "emb_lr", 0.01), seed=i)
        for i in range(collectiveArgs.num_emb_ops)
    ]
    collectiveArgs.embRequests = generate_requests(collectiveArgs.num_emb_ops, collectiveArgs.batch_size,
                                                   batched, commsParams.bag_size, num_embeddings, device=dev)
    if collectiveArgs.direction == "backward":
        # backward needs a forward output to differentiate and a gradient to push through it
        for i, (indices, offsets, weights) in enumerate(collectiveArgs.embRequests):
            collectiveArgs.LookupOut = collectiveArgs.emb[i].forward(indices, offsets, weights)
        collectiveArgs.grad_output = torch.rand_like(collectiveArgs.LookupOut)
