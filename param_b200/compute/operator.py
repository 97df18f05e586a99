"""Operator plugin for the reference's config-driven micro-benchmark framework
(train/compute/python): an object with the OperatorInterface protocol —
build / cleanup / forward / create_grad / backward (lib/operator.py:8-45) — wrapping the B200
batched EmbeddingBag, with the build() signature of the fbgemm wrapper it stands in for
(workloads/pytorch/split_table_batched_embeddings_ops.py:248-303) so that the same JSON op configs
drive it.  `register()` puts it into the reference's op_map (lib/operator.py:48-68) when that
package is importable; the class itself has no dependency on it (OperatorInterface accepts any
class with a callable `forward` through __subclasshook__, :18-24).
"""
from __future__ import annotations

from typing import List, Union

import torch

from .._cabi import PB200Error
from .tbe import B200TBE

OP_NAME = "B200BatchedEmbeddingBag"


class B200BatchedEmbeddingBagOp:
    def __init__(self) -> None:
        self.device = None          # set by the framework before build() (lib/operator.py:26-27)
        self.op = None
        self.fwd_out = None
        self.grad_in = None

    def build(self, num_tables: int, rows: Union[int, List[int]], dims: Union[int, List[int]],
              pooling: int = 0, weighted: bool = False, weights_precision: str = "fp32",
              optimizer: str = "exact_sgd", lr: float = 0.01, eps: float = 1.0e-8,
              weight_decay: float = 0.0, weight_decay_mode=None) -> None:
        dev = str(self.device or "cuda")
        if not dev.startswith("cuda"):
            raise PB200Error(f"{OP_NAME} runs on CUDA devices only (got {dev})")
        if float(weight_decay) != 0.0:
            raise PB200Error("weight decay is not implemented in the fused optimizers")
        rows_list = rows if isinstance(rows, list) else [rows] * num_tables
        dims_list = dims if isinstance(dims, list) else [dims] * num_tables
        if len(rows_list) == 1:
            rows_list = rows_list * num_tables
        if len(dims_list) == 1:
            dims_list = dims_list * num_tables
        mode = {0: "sum", 1: "mean"}.get(int(pooling))      # fbgemm PoolingMode: SUM = 0, MEAN = 1
        if mode is None:
            raise PB200Error("pooling must be 0 (sum) or 1 (mean)")
        self.weighted = bool(weighted)
        self.op = None                                       # release the previous build's tables first
        # the reference's wrapper hard-codes stochastic_rounding=True (:292); it only acts on fp16 tables
        self.op = B200TBE(list(zip(rows_list, dims_list)), learning_rate=lr, eps=eps, pooling_mode=mode,
                          optimizer=str(getattr(optimizer, "value", optimizer)),
                          weights_precision=str(getattr(weights_precision, "value", weights_precision)),
                          stochastic_rounding=True, device=torch.device(dev))

    def cleanup(self) -> None:
        # the framework calls this before every build AND after every input run (build_executor.py:158, :229):
        # only the outputs go, so that several inputs can run against one build (the reference wrapper drops
        # the op too, :303-307, and then fails on the second input of a build)
        self.fwd_out = self.grad_in = None

    def forward(self, *args, **kwargs):
        indices, offsets = args[0], args[1]
        psw = args[2] if len(args) > 2 else None
        self.fwd_out = self.op.forward(indices, offsets, psw)
        return self.fwd_out

    def create_grad(self) -> None:
        self.grad_in = torch.ones_like(self.fwd_out)

    def backward(self, grad=None) -> None:
        if grad is None:
            if self.grad_in is None:
                self.create_grad()
            grad = self.grad_in
        self.fwd_out.backward(grad)


def register(name: str = OP_NAME):
    """Add the op to the reference's global registry (needs `param_bench.train.compute.python`
    importable).  Raises ValueError on a duplicate name, like the reference."""
    from param_bench.train.compute.python.lib.operator import register_operator

    op = B200BatchedEmbeddingBagOp()
    register_operator(name, op)
    return op
