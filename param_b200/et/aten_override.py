"""Route stock `nn.EmbeddingBag` models — and execution traces captured from them — to the B200 kernels.

Importing this module re-registers the CUDA kernels of
    aten::_embedding_bag, aten::_embedding_bag_forward_only      (forward)
    aten::_embedding_bag_backward                                 (backward of the weight)
with implementations that call libparam_b200 (pb200_embbag_fwd / pb200_tbe_bwd).  This is the second
route SURVEY §8b names for et_replay compute nodes: a DLRM trace captured from the reference records
`aten::embedding_bag` (child `aten::_embedding_bag`) and `aten::_embedding_bag_backward` by name and
schema; et_replay rebuilds those ops through TorchScript IR
(et_replay/et_replay_utils.py:129-212) and the dispatcher then lands on the kernels registered here.
It is opt-in because the registration is process-wide: put the module in the replay config's
"import modules" (param_b200/et/replay-config-b200-aten.json) or import it before building the model.

Covered: fp32 CUDA weights, int64/int32 indices, mode sum / mean, per_sample_weights (sum),
include_last_offset.  Not covered (raises PB200Error, there is no fallback to the ATen kernels once
they are replaced): mode max, padding_idx >= 0, scale_grad_by_freq, non-fp32 weights.
The backward returns a DENSE gradient also when the module was built with sparse=True (the reference
builds its tables that way, train/comms/pt/pytorch_dist_backend.py:923-934): autograd accepts a strided
gradient for a strided parameter, and the replay does not look at the layout.
"""
from __future__ import annotations

import warnings

import torch

from .. import ops as _ops
from .._cabi import PB200Error

_MODES = {0: "sum", 1: "mean"}


def _check(weight, scale_grad_by_freq, mode, padding_idx):
    if mode not in _MODES:
        raise PB200Error("param_b200 aten override: EmbeddingBag mode max is not covered")
    if scale_grad_by_freq:
        raise PB200Error("param_b200 aten override: scale_grad_by_freq is not covered")
    if padding_idx is not None and padding_idx >= 0:
        raise PB200Error("param_b200 aten override: padding_idx is not covered")
    if weight is not None and weight.dtype != torch.float32:
        raise PB200Error("param_b200 aten override: fp32 tables only")


def _embedding_bag(weight, indices, offsets, scale_grad_by_freq=False, mode=0, sparse=False,
                   per_sample_weights=None, include_last_offset=False, padding_idx=-1):
    _check(weight, scale_grad_by_freq, mode, padding_idx)
    if indices.dtype != offsets.dtype:          # ATen promotes mixed int32/int64 index types
        indices, offsets = indices.to(torch.int64), offsets.to(torch.int64)
    out = _ops.embedding_bag_forward(weight.contiguous(), indices, offsets, mode=_MODES[mode],
                                     per_sample_weights=per_sample_weights,
                                     include_last_offset=include_last_offset)
    # offset2bag / bag_size / max_indices exist for ATen's own backward; the backward registered below
    # works from (indices, offsets) and ignores them
    empty = torch.empty(0, dtype=indices.dtype, device=indices.device)
    return out, empty, empty, empty


def _embedding_bag_backward(grad, indices, offsets, offset2bag, bag_size, maximum_indices, num_weights,
                            scale_grad_by_freq, mode, sparse, per_sample_weights, padding_idx=-1):
    _check(None, scale_grad_by_freq, mode, padding_idx)
    if grad.dtype != torch.float32:
        raise PB200Error("param_b200 aten override: fp32 gradients only")
    if indices.dtype != offsets.dtype:
        indices, offsets = indices.to(torch.int64), offsets.to(torch.int64)
    indices = indices.contiguous().view(-1)
    n_bags, dim = int(grad.shape[0]), int(grad.shape[1])
    # nn.EmbeddingBag offsets hold n_bags entries (n_bags + 1 with include_last_offset); the batched
    # kernel takes the closed form with the trailing end offset
    if offsets.numel() == n_bags:
        offsets = _ops.close_offsets(offsets, indices.numel())
    elif offsets.numel() != n_bags + 1:
        raise PB200Error("param_b200 aten override: offsets do not match the gradient's bag count")
    if sparse:
        # aten::_embedding_bag_backward(sparse=True) returns an uncoalesced COO tensor with one entry per
        # lookup (-> _embedding_bag_sparse_backward); same here, no dense [num_weights, dim] buffer
        return _ops.embedding_bag_backward_sparse(grad.contiguous(), indices, offsets, int(num_weights),
                                                  mode=_MODES[mode], per_sample_weights=per_sample_weights,
                                                  include_last_offset=True)
    dst = torch.zeros((int(num_weights), dim), dtype=torch.float32, device=grad.device)
    if n_bags == 0 or indices.numel() == 0:
        return dst
    row_offsets = _ops.single_table_row_offsets(int(num_weights), grad.device)
    _ops.tbe_backward(dst, row_offsets, 1, dim, indices, offsets, n_bags, grad.contiguous(), layout="TBD",
                      scale=1.0, mode=_MODES[mode], per_sample_weights=per_sample_weights, algo="auto")
    return dst


_lib = None


def enable() -> None:
    """Register the overrides (idempotent)."""
    global _lib
    if _lib is not None:
        return
    lib = torch.library.Library("aten", "IMPL")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # "Overriding a previously registered kernel ..."
        lib.impl("_embedding_bag", _embedding_bag, "CUDA")
        lib.impl("_embedding_bag_forward_only", _embedding_bag, "CUDA")
        lib.impl("_embedding_bag_backward", _embedding_bag_backward, "CUDA")
    _lib = lib


enable()
