"""The B200 batched EmbeddingBag inside the reference's config-driven micro-benchmark framework
(train/compute/python): operator + input iterator + input-data generator registered into the reference's own
registries, and the `_clear_cache` entry its OpExecutor lacks for sm_100.

Reference pieces this stands beside (the originals import fbgemm_gpu at module level, which is absent, so the
reference logs "failed to import module" for its own workload and goes on — lib/init_helper.py:42-56):
  workloads/pytorch/split_table_batched_embeddings_ops.py
      :31-89    SplitTableBatchedEmbeddingBagsCodegenInputIterator   -> B200BatchedEmbeddingBagInputIterator
      :93-135   generate_requests (alpha semantics)                   -> generate_requests
      :138-229  SplitTableBatchedEmbeddingBagsCodegenInputDataGenerator -> B200BatchedEmbeddingBagInputDataGenerator
      :239-329  SplitTableBatchedEmbeddingBagsCodegenOp                -> compute/operator.py
  lib/pytorch/op_executor.py:16-28  _clear_cache: L2 sizes for sm_70/80/90 only — KeyError on a B200 as soon as
      --cuda-l2-cache on is given; patched here with the sm_100 / sm_103 entry (126 MB).
The JSON schema is the reference's (examples/pytorch/configs/split_table_batched_embeddings_ops.json): build args
num_tables, rows, dim, pooling, weighted, weights_precision (+ kwargs optimizer ...), input args batch_size,
pooling_factor.  Config: param_b200/compute/configs/b200_batched_embedding_bag.json.

    python -m param_b200.integration.param_plugin bench -c param_b200/compute/configs/b200_batched_embedding_bag.json \\
        -d cuda -w 2 -i 5 -b
"""
from __future__ import annotations

import copy
import os
from typing import Any, Dict, Optional

import numpy as np
import torch

from .operator import OP_NAME, B200BatchedEmbeddingBagOp

ITERATOR_NAME = "B200BatchedEmbeddingBagInputIterator"
GENERATOR_NAME = "B200BatchedEmbeddingBagInputDataGenerator"
L2_BYTES_SM100 = 126 * 1024 * 1024


def generate_requests(B: int, L: int, E: int, offset_start: int, alpha: float = 1.0, weighted: bool = False,
                      rng: Optional[np.random.Generator] = None):
    """One table's part of a TBE request, with the reference's alpha convention
    (split_table_batched_embeddings_ops.py:93-135): 0 -> arange % L, <= 0.5 -> arange % E, <= 1 -> uniform,
    > 1 -> numpy Zipf(alpha) % E.  offsets continue from offset_start (the first table contributes the
    leading 0)."""
    n = B * L
    if alpha == 0:
        indices = torch.arange(0, n).long() % L
    elif alpha <= 0.5:
        indices = torch.arange(0, n).long() % E
    elif alpha <= 1.0:
        indices = torch.randint(low=0, high=E, size=(n,), dtype=torch.int64)
    else:
        draw = (rng or np.random).zipf(a=alpha, size=n)
        indices = torch.as_tensor(draw).long() % E
    lengths = np.ones(B, dtype=np.int64) * L
    if offset_start == 0:
        offsets = torch.tensor(np.cumsum([0] + lengths.tolist()))
    else:
        offsets = torch.tensor(offset_start + np.cumsum(lengths))
    weights = torch.randn(n, dtype=torch.float32) if weighted else None
    return indices, offsets, weights


class B200BatchedEmbeddingBagInputDataGenerator:
    """get_data(config, device) -> ([indices, offsets, per_sample_weights], {}) in the TBE request layout.
    Same config positions as the reference generator: args[0] num_tables, [1] rows, [3] batch_size,
    [4] pooling_factor, [5] weighted."""

    def get_data(self, config: Dict[str, Any], device: str, alpha: float = 1.0):
        a = config["args"]
        num_tables = int(a[0]["value"])
        rows, pooling = a[1]["value"], a[4]["value"]
        rows = list(rows) if isinstance(rows, (list, tuple)) else [rows] * num_tables
        pooling = list(pooling) if isinstance(pooling, (list, tuple)) else [pooling] * num_tables
        if len(rows) == 1:
            rows = rows * num_tables
        if len(pooling) == 1:
            pooling = pooling * num_tables
        batch, weighted = int(a[3]["value"]), bool(a[5]["value"])
        dist = os.getenv("split_embedding_distribution")     # the reference reads the same variable (:160)
        alpha = float(dist) if dist is not None else float(alpha)
        ind, off, wts, start = [], [], [], 0
        for t in range(num_tables):
            i, o, w = generate_requests(batch, int(pooling[t]), int(rows[t]), start, alpha, weighted)
            ind.append(i)
            off.append(o)
            start = int(o[-1])
            if weighted:
                wts.append(w)
        dev = torch.device(device)
        return ([torch.cat(ind).to(dev), torch.cat(off).to(dev), torch.cat(wts).to(dev) if weighted else None], {})


def _make_iterator_class():
    from param_bench.train.compute.python.lib.generator import full_range, IterableList, ListProduct, TableProduct
    from param_bench.train.compute.python.lib.iterator import ConfigIterator, remove_meta_attr

    class B200BatchedEmbeddingBagInputIterator(ConfigIterator):
        """yields (id, {"args": [num_tables, rows, dim, batch_size, pooling_factor, weighted, weights_precision]})
        for every combination of the input ranges, like the reference iterator (:31-86)"""

        def __init__(self, configs, key, device):
            super().__init__(configs, key, device)
            b = configs["build"]["args"]
            self.num_tables, self.rows, self.dim, self.weighted, self.precision = b[0], b[1], b[2], b[4], b[5]
            self.generator = self._generator()

        def _generator(self):
            for var_id, inp in enumerate(self.configs[self.key]):
                args = []
                for arg in copy.deepcopy(inp)["args"]:
                    if "__range__" in arg:
                        arg["value"] = full_range(*arg["value"])
                    if "__list__" in arg:
                        arg["value"] = IterableList(arg["value"])
                    args.append(TableProduct(arg))
                for config_id, (batch_size, pooling_factor) in enumerate(ListProduct(args)):
                    result = {"args": [self.num_tables, self.rows, self.dim, batch_size, pooling_factor,
                                       self.weighted, self.precision], "kwargs": {}}
                    yield (f"{var_id}_{config_id}", remove_meta_attr(result))

        def __next__(self):
            return next(self.generator)

    return B200BatchedEmbeddingBagInputIterator


def patch_clear_cache() -> None:
    """op_executor._clear_cache knows sm_70 / 80 / 90 only (op_executor.py:16-28): add Blackwell."""
    from param_bench.train.compute.python.lib.pytorch import op_executor

    if getattr(op_executor._clear_cache, "_pb200", False):
        return
    orig = op_executor._clear_cache

    def _clear_cache(device: torch.device):
        cap = torch.cuda.get_device_capability(device)
        if cap[0] >= 10:
            with torch.autograd.profiler.record_function("[param|clear_cache]"):
                _ = torch.zeros(L2_BYTES_SM100 // 4, device=device).float() * 2    # write a buffer larger than L2
                del _
                torch.cuda.empty_cache()
            return
        return orig(device)

    _clear_cache._pb200 = True
    op_executor._clear_cache = _clear_cache


def register_all() -> B200BatchedEmbeddingBagOp:
    """Operator, input iterator and input-data generator into the reference's registries (idempotent)."""
    from param_bench.train.compute.python.lib import data as ref_data, iterator as ref_iter, operator as ref_op

    op = ref_op.op_map.get(OP_NAME)
    if op is None:
        op = B200BatchedEmbeddingBagOp()
        ref_op.register_operator(OP_NAME, op)
    if ITERATOR_NAME not in ref_iter.config_iterator_map:
        ref_iter.register_config_iterator(ITERATOR_NAME, _make_iterator_class())
    if GENERATOR_NAME not in ref_data.data_generator_map:
        ref_data.register_data_generator(GENERATOR_NAME, B200BatchedEmbeddingBagInputDataGenerator)
    patch_clear_cache()
    return op


def run_benchmark(argv) -> None:
    """the reference's train/compute/python/pytorch/run_benchmark.py, unmodified, with the plugin registered"""
    import sys

    register_all()
    from param_bench.train.compute.python.pytorch import run_benchmark as rb

    sys.argv = ["run_benchmark.py"] + list(argv)
    rb.main()
