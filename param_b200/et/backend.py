"""Comm backend for et_replay (the fork of the PARAM backend boundary used by trace replay).

et_replay/tools/comm_replay.py:1734-1761 picks `customized_backend[commsParams.backend]` for an
unknown --backend name, calls initialize_backend(master_ip, master_port, backend=<name>) and
barrier_all_ranks(); replayed `record_param_comms` nodes reach the backend as
collectiveFunc[collName](collectiveArgs, retFlag=True) (comm_replay.py:956-1106), c10d
`alltoall_base` arriving as "all_to_allv" (et_replay/comm/comms_utils.py:373) with element-count
splits parsed from the trace (commsTraceParser.py:227-229).

  * with the et_replay package importable:  register_et_backend()  ->
        class B200ETBackend(B200CommsMixin, et_replay PyTorchDistBackend)
    registered through et_replay.comm.backend.base_backend.register_customized_backend("b200", ...);
  * stand-alone (no et_replay on the box):  B200ETStandalone — B200Backend plus the extra entries
    BaseBackend requires (base_backend.py:195-330): allgather_into_tensor_coalesced,
    allreduce_coalesced, reduce_scatter_tensor_coalesced, barrier_all_ranks, keyed wait.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..comms.pt.backend import B200Backend, B200CommsMixin


class _ETKeyedWaitMixin:
    """et_replay keys async work by (pg_id, seq_id, is_p2p) — pytorch_dist_backend.py:684-691."""

    def wait(self, collectiveArgs, retFlag=False):
        key = getattr(collectiveArgs, "wait_obj_key", None)
        ids = getattr(collectiveArgs, "waitObjIds", {})
        if key in ids:
            work = ids.pop(key)
            collectiveArgs.waitObj[:] = [w for w in collectiveArgs.waitObj if w is not work]
            if work is not None:
                work.wait()

    def complete_accel_ops(self, collectiveArgs, devSync=True):
        for req in collectiveArgs.waitObj:
            if req is not None:
                req.wait()
        collectiveArgs.waitObj.clear()
        if hasattr(collectiveArgs, "waitObjIds"):
            collectiveArgs.waitObjIds.clear()
        if devSync:
            self.device_sync(collectiveArgs)


class B200ETStandalone(_ETKeyedWaitMixin, B200Backend):
    def __init__(self, bootstrap_info, commsParams):
        super().__init__(bootstrap_info, commsParams)
        self.collectiveFunc.update({
            "allgather_into_tensor_coalesced": self.allgather_into_tensor_coalesced,
            "allreduce_coalesced": self.allreduce_coalesced,
            "reduce_scatter_tensor_coalesced": self.reduce_scatter_tensor_coalesced,
        })

    def initialize_backend(self, master_ip, master_port, backend="nccl", eager_mode=False):
        return super().initialize_backend(master_ip, master_port,
                                          backend="nccl" if backend == "b200" else backend,
                                          eager_mode=eager_mode)

    # coalesced forms: thin c10d pass-throughs (NCCL), not on the kernel path
    def allgather_into_tensor_coalesced(self, collectiveArgs, retFlag=False, pair=False):
        with dist._coalescing_manager(group=collectiveArgs.group, device=self.get_device(),
                                      async_ops=collectiveArgs.asyncOp) as cm:
            for o, i in zip(collectiveArgs.opTensor, collectiveArgs.ipTensor):
                dist.all_gather_into_tensor(o, i, group=collectiveArgs.group)
        return self._finish(collectiveArgs, cm if collectiveArgs.asyncOp else None, retFlag)

    def allreduce_coalesced(self, collectiveArgs, retFlag=False, pair=False):
        with dist._coalescing_manager(group=collectiveArgs.group, device=self.get_device(),
                                      async_ops=collectiveArgs.asyncOp) as cm:
            for t in collectiveArgs.ipTensor:
                dist.all_reduce(t, op=getattr(collectiveArgs, "op", dist.ReduceOp.SUM),
                                group=collectiveArgs.group)
        return self._finish(collectiveArgs, cm if collectiveArgs.asyncOp else None, retFlag)

    def reduce_scatter_tensor_coalesced(self, collectiveArgs, retFlag=False, pair=False):
        with dist._coalescing_manager(group=collectiveArgs.group, device=self.get_device(),
                                      async_ops=collectiveArgs.asyncOp) as cm:
            for o, i in zip(collectiveArgs.opTensor, collectiveArgs.ipTensor):
                dist.reduce_scatter_tensor(o, i, op=getattr(collectiveArgs, "op", dist.ReduceOp.SUM),
                                           group=collectiveArgs.group)
        return self._finish(collectiveArgs, cm if collectiveArgs.asyncOp else None, retFlag)


def register_et_backend(name: str = "b200"):
    """Register against the real et_replay package (must be importable)."""
    from et_replay.comm.backend.base_backend import register_customized_backend
    from et_replay.comm.backend.pytorch_dist_backend import PyTorchDistBackend

    class B200ETBackend(_ETKeyedWaitMixin, B200CommsMixin, PyTorchDistBackend):
        def __init__(self, bootstrap_info, commsParams):
            PyTorchDistBackend.__init__(self, bootstrap_info, commsParams)
            self.collectiveFunc["all_to_allv"] = self.all_to_allv
            self.collectiveFunc["all_to_all"] = self.all_to_all
            self.collectiveFunc["all_to_all_single"] = self.all_to_all_single
            self.computeFunc["emb_lookup"] = self.emb_lookup

        def initialize_backend(self, master_ip, master_port, backend="nccl", eager_mode=False):
            return PyTorchDistBackend.initialize_backend(
                self, master_ip, master_port, backend="nccl" if backend == "b200" else backend)

        def initialize_groups(self, backend="nccl"):
            return PyTorchDistBackend.initialize_groups(self, "nccl" if backend == "b200" else backend)

    register_customized_backend(name, B200ETBackend)
    return B200ETBackend
