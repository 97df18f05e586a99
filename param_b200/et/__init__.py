"""et_replay dispatch hook.

`import param_b200.et` registers the B200 kernels as dispatcher ops in the `b200::` namespace, with
full schema strings.  That is all et_replay needs to replay a trace that calls them: compute nodes
are re-created purely from node.name + node.op_schema through TorchScript IR
(et_replay/et_replay_utils.py:129-212 build_torchscript_func), and third-party ops are pulled in
through the replay config's "import modules" list (et_replay_utils.py:242-249,
tools/et_replay.py:383-389; precedent configs/replay-config-fbgemm.json).  The matching config is
param_b200/et/replay-config-b200.json.

Ops (all CUDA only; a CPU tensor raises — there is no fallback):
  b200::embedding_bag(Tensor weight, Tensor indices, Tensor offsets, int mode,
                      Tensor? per_sample_weights, bool include_last_offset) -> Tensor
  b200::tbe_forward(Tensor weights, Tensor row_offsets, int dim, Tensor indices, Tensor offsets,
                    int batch, int mode, Tensor? per_sample_weights, int layout) -> Tensor
  b200::tbe_backward_(Tensor(a!) dst, Tensor row_offsets, int dim, Tensor indices, Tensor offsets,
                      int batch, Tensor grad_out, int layout, float scale, int mode, int algo) -> Tensor(a!)
  b200::tbe_backward_fused_(Tensor(a!) weights, Tensor(b!)? state, Tensor row_offsets, int dim,
                            Tensor indices, Tensor offsets, int batch, Tensor grad_out, int layout,
                            int mode, Tensor? per_sample_weights, int optimizer, float lr, float eps,
                            bool stochastic_rounding, int sr_seed) -> Tensor(a!)
        (in-place ops return the mutated tensor: et_replay's IR builder cannot re-create an op with
        no output, et_replay_utils.py:171-200)
        the fused backward + optimizer a DLRM trace records as fbgemm's
        split_embedding_backward_codegen_*_exact ops; optimizer 1 = exact_sgd, 2 = exact_row_wise_adagrad;
        weights fp32 or fp16
  b200::regroup_sparse(Tensor lengths, Tensor indices, int world, int tables_local,
                       int local_batch) -> (Tensor, Tensor, Tensor)
Communication nodes (`record_param_comms` all_to_allv etc.) do not go through op names: et_replay
hands them to its comm backend (comm_replay.py:956-1106) — see param_b200/et/backend.py.
"""
from __future__ import annotations

import torch

from .. import ops as _ops

_MODES = {0: "sum", 1: "mean"}
_LAYOUT = {0: "BTD", 1: "TBD"}
_BWD = {0: "auto", 1: "atomic", 2: "sorted", 3: "exact"}
_OPT = {1: "exact_sgd", 2: "exact_row_wise_adagrad"}

_lib = torch.library.Library("b200", "DEF")
_lib.define("embedding_bag(Tensor weight, Tensor indices, Tensor offsets, int mode, "
            "Tensor? per_sample_weights, bool include_last_offset) -> Tensor")
_lib.define("tbe_forward(Tensor weights, Tensor row_offsets, int dim, Tensor indices, Tensor offsets, "
            "int batch, int mode, Tensor? per_sample_weights, int layout) -> Tensor")
_lib.define("tbe_backward_(Tensor(a!) dst, Tensor row_offsets, int dim, Tensor indices, Tensor offsets, "
            "int batch, Tensor grad_out, int layout, float scale, int mode, int algo) -> Tensor(a!)")
_lib.define("tbe_backward_fused_(Tensor(a!) weights, Tensor(b!)? state, Tensor row_offsets, int dim, "
            "Tensor indices, Tensor offsets, int batch, Tensor grad_out, int layout, int mode, "
            "Tensor? per_sample_weights, int optimizer, float lr, float eps, bool stochastic_rounding, "
            "int sr_seed) -> Tensor(a!)")
_lib.define("regroup_sparse(Tensor lengths, Tensor indices, int world, int tables_local, "
            "int local_batch) -> (Tensor, Tensor, Tensor)")


def _embedding_bag(weight, indices, offsets, mode, per_sample_weights, include_last_offset):
    return _ops.embedding_bag_forward(weight, indices, offsets, mode=_MODES[mode],
                                      per_sample_weights=per_sample_weights,
                                      include_last_offset=include_last_offset)


def _tbe_forward(weights, row_offsets, dim, indices, offsets, batch, mode, per_sample_weights, layout):
    T = row_offsets.numel() - 1
    arena = _ops.TableArena(weights, row_offsets, [0] * T, dim)
    return _ops.tbe_forward(arena, indices, offsets, batch, mode=_MODES[mode],
                            per_sample_weights=per_sample_weights, layout=_LAYOUT[layout])


def _tbe_backward_(dst, row_offsets, dim, indices, offsets, batch, grad_out, layout, scale, mode, algo):
    _ops.tbe_backward(dst, row_offsets, row_offsets.numel() - 1, dim, indices, offsets, batch, grad_out,
                      layout=_LAYOUT[layout], scale=scale, mode=_MODES[mode], algo=_BWD[algo])
    return dst


def _tbe_backward_fused_(weights, state, row_offsets, dim, indices, offsets, batch, grad_out, layout, mode,
                         per_sample_weights, optimizer, lr, eps, stochastic_rounding, sr_seed):
    _ops.tbe_backward_fused(weights, row_offsets, row_offsets.numel() - 1, dim, indices, offsets, batch,
                            grad_out, optimizer=_OPT[optimizer], lr=lr, eps=eps, state=state,
                            layout=_LAYOUT[layout], mode=_MODES[mode], per_sample_weights=per_sample_weights,
                            stochastic_rounding=stochastic_rounding, sr_seed=sr_seed)
    return weights


def _regroup_sparse(lengths, indices, world, tables_local, local_batch):
    lo, off, idx = _ops.regroup_sparse(lengths, indices, world, tables_local, local_batch)
    return lo, off, idx


for _name, _fn in (("embedding_bag", _embedding_bag), ("tbe_forward", _tbe_forward),
                   ("tbe_backward_", _tbe_backward_), ("tbe_backward_fused_", _tbe_backward_fused_),
                   ("regroup_sparse", _regroup_sparse)):
    _lib.impl(_name, _fn, "CUDA")
    _lib.impl(_name, _fn, "CPU")   # reaches ops._need_cuda -> raises PB200Error: no CPU fallback
