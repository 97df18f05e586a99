"""B200 backend for PARAM's `backendFunctions` plugin boundary.

Surface mirrored (names, argument meaning, error behaviour):
  train/comms/pt/pytorch_backend_utils.py:156-411  backendFunctions ABC + collectiveFunc table
  train/comms/pt/pytorch_dist_backend.py            PyTorchDistBackend (the c10d implementation)
  et_replay/comm/backend/base_backend.py:136-330    BaseBackend (fork used by et_replay)

What changes: the all-to-all family (all_to_all_single :330-357, all_to_allv :262-328, all_to_all
:207-260) runs as the peer-push kernel over NVLink (libparam_b200, pb200_a2a_single); emb_lookup
(:832-857) and alloc_embedding_tables (:923-934) use the sm_100a EmbeddingBag kernels; comm buffers
come from the peer-mapped window (alloc_empty/alloc_random are the hook, exactly as
PyTorchNVShmemBackend.alloc_empty does, pytorch_nvshmem_backend.py:29-40).  Every other collective
is a thin c10d pass-through (NCCL), out of scope for kernels.

`B200CommsMixin` holds the overrides; `B200Backend` is a stand-alone class for boxes where the
reference is not installed; param_b200/integration/param_plugin.py combines the mixin with the
real PyTorchDistBackend and registers it with register_customized_backend("b200", ...).
"""
from __future__ import annotations

import logging
import os
from itertools import cycle
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from ..._cabi import PB200Error
from ...compute.pt.pytorch_emb import B200EmbeddingBag
from .peer_window import PeerWindow

logger = logging.getLogger(__name__)

DEFAULT_WINDOW_BYTES = int(os.environ.get("PB200_WINDOW_BYTES", str(5 << 30)))


class StreamOrderedWork:
    """What an async collective returns: the kernel was enqueued on the caller's current stream, so
    `wait()` only has to make later work on the *waiting* stream depend on it (c10d Work.wait()
    contract; the reference's DummyWork, pytorch_nvshmem_backend.py:14-19, does nothing at all)."""

    def __init__(self, device):
        self.event = torch.cuda.Event()
        self.event.record(torch.cuda.current_stream(device))
        self.device = device

    def wait(self, timeout=None):
        torch.cuda.current_stream(self.device).wait_event(self.event)
        return True

    def is_completed(self):
        return self.event.query()


def _param(commsParams, name, default=None):
    if isinstance(commsParams, dict):
        return commsParams.get(name, default)
    return getattr(commsParams, name, default)


class B200CommsMixin:
    """Overrides of the hot-path entries.  Host class must provide get_device()."""

    _window: Optional[PeerWindow] = None
    _window_group = None

    # ---- window management -------------------------------------------------------------------
    def _ensure_window(self, group=None) -> PeerWindow:
        """The peer window of `group`, mapped ONCE with PB200_WINDOW_BYTES (default 5 GB).  Mapping is a
        collective (rendezvous + barrier), so it is never triggered from a rank-local size: a request
        that does not fit raises on the rank that sees it (PeerWindow._staging_offset) with the knob to
        turn, instead of re-mapping under the feet of the peers."""
        group = group if group is not None else dist.group.WORLD
        if self._window is None or self._window_group is not group:
            if self._window is not None:
                self._window.close()
            self._window = PeerWindow.create(group, DEFAULT_WINDOW_BYTES, self.get_device())
            self._window_group = group
            logger.info("b200: mapped %d B peer window via %s", DEFAULT_WINDOW_BYTES,
                        getattr(self._window, "mapping", "?"))
        return self._window

    def _window_alloc(self, numel: int, dtype: torch.dtype) -> Optional[torch.Tensor]:
        """comm buffers live in the peer window so that all_to_all output is written in place"""
        if not dist.is_initialized() or not torch.cuda.is_available():
            return None
        try:
            win = self._ensure_window()
            return win.alloc(int(numel), dtype)[0]
        except PB200Error:
            return None  # larger than the window: plain allocation, result is staged + copied

    # ---- allocation hooks (pytorch_dist_backend.py:899-934) ---------------------------------------
    def alloc_empty(self, sizeArr, curRankDevice="cuda", dtype=torch.float32):
        numel = int(np.prod(sizeArr)) if not isinstance(sizeArr, int) else int(sizeArr)
        if str(curRankDevice).startswith("cuda"):
            t = self._window_alloc(numel, dtype)
            if t is not None:
                return t.view(*sizeArr) if not isinstance(sizeArr, int) else t
        return torch.empty(sizeArr, device=curRankDevice, dtype=dtype)

    def alloc_random(self, sizeArr, curRankDevice="cuda", dtype=torch.float32, scaleFactor=1.0):
        out = self.alloc_empty(sizeArr, curRankDevice, dtype)
        if dtype in (torch.int8, torch.uint8, torch.short, torch.int16, torch.int32, torch.long):
            out.copy_(torch.randint(0, 10, out.shape, device=out.device, dtype=dtype))
        elif dtype == torch.bool:
            out.copy_(torch.rand(out.shape, device=out.device) < 0.5)
        else:
            out.copy_(torch.rand(out.shape, device=out.device, dtype=dtype))
            if scaleFactor != 0:
                out.div_(scaleFactor)
        return out

    def alloc_ones(self, sizeArr, curRankDevice="cuda", dtype=torch.float32, scaleFactor=1.0):
        out = self.alloc_empty(sizeArr, curRankDevice, dtype)
        out.fill_(1)
        if scaleFactor != 1.0:
            out.mul_(scaleFactor)
        return out

    def alloc_embedding_tables(self, n, m, curRankDevice, dtype):
        """nn.EmbeddingBag(n, m, mode="sum", sparse=True) with W ~ U(+-sqrt(1/n))
        (pytorch_dist_backend.py:923-934), as the B200 module."""
        if dtype != torch.float32:
            raise PB200Error("B200 embedding tables are fp32")
        bound = float(np.sqrt(1.0 / n))
        w = torch.empty((n, m), dtype=torch.float32, device=curRankDevice).uniform_(-bound, bound)
        return B200EmbeddingBag(n, m, mode="sum", sparse=True, _weight=w)

    def clear_memory(self, collectiveArgs):
        for name in ("ipTensor", "opTensor"):
            if hasattr(collectiveArgs, name):
                try:
                    delattr(collectiveArgs, name)
                except AttributeError:
                    pass
        for lst in (getattr(collectiveArgs, "ipTensor_pair", None), getattr(collectiveArgs, "opTensor_pair", None)):
            if isinstance(lst, list):
                lst.clear()
        if self._window is not None:
            self._window.reset_alloc()
        torch.cuda.empty_cache()

    # ---- the all-to-all family ------------------------------------------------------------------
    def _a2a(self, out, inp, out_splits, in_splits, group, async_op):
        group = group if group is not None else dist.group.WORLD
        if not inp.is_cuda:
            raise PB200Error("the b200 backend moves CUDA tensors only (no CPU/gloo fallback)")
        win = self._ensure_window(group)
        if out.dtype != inp.dtype:
            raise PB200Error("all_to_all needs input and output of one dtype")
        if not out.is_contiguous():
            raise PB200Error("all_to_all needs a contiguous output tensor")
        win.all_to_all_single(out, inp.contiguous(),
                              list(out_splits) if out_splits is not None and len(out_splits) else None,
                              list(in_splits) if in_splits is not None and len(in_splits) else None)
        return StreamOrderedWork(inp.device) if async_op else None

    def all_to_all_single(self, collectiveArgs, retFlag=False, pair=False, pairIdx=0):
        if getattr(collectiveArgs, "all2all_qcomm", None):
            logger.warning("all_to_all_single does not support quantization")
            return
        op = collectiveArgs.opTensor if not pair else collectiveArgs.opTensor_pair[pairIdx]
        ip = collectiveArgs.ipTensor if not pair else collectiveArgs.ipTensor_pair[pairIdx]
        work = self._a2a(op, ip, collectiveArgs.opTensor_split, collectiveArgs.ipTensor_split,
                         collectiveArgs.group, collectiveArgs.asyncOp)
        if collectiveArgs.asyncOp:
            collectiveArgs.waitObj.append(work)
        if retFlag:
            return work

    def all_to_allv(self, collectiveArgs, retFlag=False, pair=False, pairIdx=0):
        op = collectiveArgs.opTensor if not pair else collectiveArgs.opTensor_pair[pairIdx]
        ip = collectiveArgs.ipTensor if not pair else collectiveArgs.ipTensor_pair[pairIdx]
        osp = collectiveArgs.opTensor_split if not pair else collectiveArgs.opTensor_split_pair[pairIdx]
        isp = collectiveArgs.ipTensor_split if not pair else collectiveArgs.ipTensor_split_pair[pairIdx]
        if not pair and op.dtype != ip.dtype:
            # et_replay re-types the holder's output tensor on a dtype mismatch
            # (et_replay/comm/backend/pytorch_dist_backend.py:350-356)
            logger.warning("all_to_allv: opTensor and ipTensor are not the same dtype")
            collectiveArgs.opTensor = op = op.to(ip.dtype)
        work = self._a2a(op, ip, osp, isp, collectiveArgs.group, collectiveArgs.asyncOp)
        if collectiveArgs.asyncOp:
            collectiveArgs.waitObj.append(work)
        if retFlag:
            return work

    def all_to_all(self, collectiveArgs, retFlag=False, pair=False, pairIdx=0):
        """list form (dist.all_to_all(list_out, list_in), pytorch_dist_backend.py:207-260): ONE push kernel
        with per-destination source pointers and per-source landing offsets (pb200_a2a_list) — no cat
        before, no split/copy after when the outputs live in the window"""
        ops_ = collectiveArgs.opTensor if not pair else collectiveArgs.opTensor_pair[pairIdx]
        ips_ = collectiveArgs.ipTensor if not pair else collectiveArgs.ipTensor_pair[pairIdx]
        if not isinstance(ips_, (list, tuple)):
            raise PB200Error("all_to_all expects lists of tensors")
        if not isinstance(ops_, (list, tuple)) or len(ops_) != len(ips_):
            raise PB200Error("all_to_all expects output and input lists of world_size tensors")
        group = collectiveArgs.group if collectiveArgs.group is not None else dist.group.WORLD
        win = self._ensure_window(group)
        # one push kernel: ips_[j] -> rank j, landing in that rank's ops_[me] (in place if it lives in
        # the window, which is where alloc_empty puts comm buffers)
        if any(not t.is_contiguous() for t in ops_):
            raise PB200Error("all_to_all needs contiguous output tensors")
        win.all_to_all(list(ops_), [t.contiguous() for t in ips_])
        work = StreamOrderedWork(ips_[0].device) if collectiveArgs.asyncOp else None
        if collectiveArgs.asyncOp:
            collectiveArgs.waitObj.append(work)
        if retFlag:
            return work

    # ---- compute function (pytorch_dist_backend.py:832-857) ---------------------------------------
    def emb_lookup(self, collectiveArgs):
        if collectiveArgs.direction == "forward":
            for i in range(len(collectiveArgs.embRequests)):
                indices, offsets, weights = collectiveArgs.embRequests[i]
                collectiveArgs.LookupOut = collectiveArgs.emb[i].forward(indices, offsets, weights)
        else:
            for _ in range(len(collectiveArgs.embRequests)):
                collectiveArgs.LookupOut.backward(collectiveArgs.grad_output,
                                                  retain_graph=collectiveArgs.reuseTensors)

    # ---- completion ---------------------------------------------------------------------------------
    def complete_accel_ops(self, collectiveArgs, devSync=True):
        for req in collectiveArgs.waitObj:
            if req is not None:
                req.wait()
        if devSync:
            self.device_sync(collectiveArgs)
        collectiveArgs.waitObj.clear()
        if hasattr(collectiveArgs, "waitObjIds"):
            collectiveArgs.waitObjIds.clear()

    def device_sync(self, collectiveArgs):
        torch.cuda.synchronize(getattr(collectiveArgs, "device", None) or self.get_device())


class B200Backend(B200CommsMixin):
    """Stand-alone implementation of the backendFunctions surface (used when the reference package
    is not importable, e.g. on the benchmark box).  Hot-path entries come from B200CommsMixin; the
    rest are c10d pass-throughs with the reference's call convention
    fn(collectiveArgs, retFlag=False, pair=False, pairIdx=0)."""

    def __init__(self, bootstrap_info, commsParams):
        self.bootstrap_info = bootstrap_info
        self.commsParams = commsParams
        self.use_ext_dist = False
        self.tcp_store = None
        self.groups, self.groupRanks, self.num_pgs = {}, {}, 0
        self.collectiveFunc = {
            "all_to_all_single": self.all_to_all_single, "all_to_all": self.all_to_all,
            "all_to_allv": self.all_to_allv, "all_reduce": self.all_reduce,
            "broadcast": self.broadcast, "all_gather": self.all_gather,
            "all_gather_base": self.all_gather_base, "reduce": self.reduce,
            "reduce_scatter_base": self.reduce_scatter_base, "barrier": self.barrier,
            "noop": self.noop, "wait": self.wait,
        }
        self.computeFunc = {"emb_lookup": self.emb_lookup, "gemm": self.gemm}

    # -- bootstrap (pytorch_dist_backend.py:1145-1251) --
    def initialize_tcpstore(self, master_ip, master_port):
        bi = self.bootstrap_info
        self.tcp_store = dist.TCPStore(master_ip, int(master_port), bi.world_size,
                                       is_master=(bi.global_rank == 0), use_libuv=True)

    def initialize_backend(self, master_ip, master_port, backend="nccl", eager_mode=False):
        bi = self.bootstrap_info
        if backend not in ("nccl",):
            raise PB200Error("the b200 backend bootstraps over nccl only (one process per GPU, one node)")
        self.set_device(bi.local_rank, bi.global_rank)
        if not dist.is_initialized():
            if self.tcp_store is None:
                self.initialize_tcpstore(master_ip, master_port)
            dist.init_process_group(backend, rank=bi.global_rank, world_size=bi.world_size,
                                    store=self.tcp_store,
                                    device_id=torch.device(f"cuda:{bi.local_rank}"))
        self.groups = {0: self.get_default_group()}
        self.num_pgs = 1
        self.round_robin_group = cycle(list(self.groups.values()))

    def initialize_groups(self, groupRanks=None, backend="nccl", force_new_group=False):
        if groupRanks is not None:
            self.groupRanks = groupRanks
        groups, world = {}, self.get_world_size()
        for pg_id, ranks in self.groupRanks.items():
            if len(ranks) > world:
                groups.clear()
                break
            groups[pg_id] = self.get_default_group() if len(ranks) == world and not force_new_group \
                else dist.new_group(ranks=ranks, backend=backend)
        if groups:
            self.groups = groups
        self.num_pgs = len(self.groups)
        self.round_robin_group = cycle(list(self.groups.values()))

    def sayHello(self, *unused):
        r, w = self.get_global_rank(), self.get_world_size()
        msg = (f"[Rank {r:3}] host {os.uname()[1]}, device: {self.get_device()}, "
               f"local_rank: {self.get_local_rank()} world_size: {w}, master_ip: {self.bootstrap_info.master_ip}")
        if self.tcp_store is None:
            print(msg)
            return
        self.store_set(f"hello_msg_{r}", msg)
        if r == 0:
            for k in range(w):
                print(f"Hello from Rank {k}: {self.store_get(f'hello_msg_{k}').decode()}")

    def store_get(self, key):
        return self.tcp_store.get(key)

    def store_set(self, key, val):
        self.tcp_store.set(key, val)

    def benchmark_comms(self, benchTime, commsParams):
        if getattr(commsParams, "init_only", False):
            return
        benchTime(0, commsParams, self)

    def set_up(self):
        return

    def tear_down(self):
        return

    # -- pass-through collectives (NCCL; out of scope for kernels) --
    def _finish(self, collectiveArgs, work, retFlag):
        if collectiveArgs.asyncOp:
            collectiveArgs.waitObj.append(work)
        if retFlag:
            return work

    def all_reduce(self, collectiveArgs, retFlag=False, pair=False, pairIdx=0):
        t = collectiveArgs.ipTensor if not pair else collectiveArgs.ipTensor_pair[pairIdx]
        w = dist.all_reduce(t, op=getattr(collectiveArgs, "op", dist.ReduceOp.SUM),
                            group=collectiveArgs.group, async_op=collectiveArgs.asyncOp)
        return self._finish(collectiveArgs, w, retFlag)

    def reduce(self, collectiveArgs, retFlag=False, pair=False, pairIdx=0):
        w = dist.reduce(collectiveArgs.ipTensor, dst=collectiveArgs.srcOrDst,
                        op=getattr(collectiveArgs, "op", dist.ReduceOp.SUM),
                        group=collectiveArgs.group, async_op=collectiveArgs.asyncOp)
        return self._finish(collectiveArgs, w, retFlag)

    def broadcast(self, collectiveArgs, retFlag=False, pair=False, pairIdx=0):
        w = dist.broadcast(collectiveArgs.opTensor, src=collectiveArgs.srcOrDst,
                           group=collectiveArgs.group, async_op=collectiveArgs.asyncOp)
        return self._finish(collectiveArgs, w, retFlag)

    def all_gather(self, collectiveArgs, retFlag=False, pair=False, pairIdx=0):
        w = dist.all_gather(collectiveArgs.opTensor, collectiveArgs.ipTensor,
                            group=collectiveArgs.group, async_op=collectiveArgs.asyncOp)
        return self._finish(collectiveArgs, w, retFlag)

    def all_gather_base(self, collectiveArgs, retFlag=False, pair=False, pairIdx=0):
        w = dist.all_gather_into_tensor(collectiveArgs.opTensor, collectiveArgs.ipTensor,
                                        group=collectiveArgs.group, async_op=collectiveArgs.asyncOp)
        return self._finish(collectiveArgs, w, retFlag)

    def reduce_scatter_base(self, collectiveArgs, retFlag=False, pair=False, pairIdx=0):
        w = dist.reduce_scatter_tensor(collectiveArgs.opTensor, collectiveArgs.ipTensor,
                                       op=getattr(collectiveArgs, "op", dist.ReduceOp.SUM),
                                       group=collectiveArgs.group, async_op=collectiveArgs.asyncOp)
        return self._finish(collectiveArgs, w, retFlag)

    def barrier(self, collectiveArgs, name="dummy", retFlag=False):
        w = dist.barrier(collectiveArgs.group, async_op=collectiveArgs.asyncOp,
                         device_ids=[self.get_device().index])
        return self._finish(collectiveArgs, w, retFlag)

    def barrier_all_ranks(self):
        dist.barrier(device_ids=[self.get_device().index])

    def sync_barrier(self, collectiveArgs, desc="dummy"):
        self.complete_accel_ops(collectiveArgs)
        self.barrier(collectiveArgs, name=desc)
        self.complete_accel_ops(collectiveArgs)

    def wait(self, collectiveArgs, retFlag=False):
        ids = getattr(collectiveArgs, "waitObjIds", {})
        key = getattr(collectiveArgs, "wait_obj_key", getattr(collectiveArgs, "collectiveId", None))
        if ids and key in ids:
            w = ids.pop(key)
            if w is not None:
                w.wait()
        elif collectiveArgs.waitObj:
            w = collectiveArgs.waitObj.pop(0)
            if w is not None:
                w.wait()
            self.device_sync(collectiveArgs)

    def noop(self, collectiveArgs=None, retFlag=False, pair=False):
        return None

    def gemm(self, collectiveArgs):
        raise PB200Error("gemm is dense tensor-core work, outside the B200 hot path (SURVEY §2 row 1)")

    def get_reduce_op(self, opName):
        return dist.ReduceOp.MAX if opName == "max" else dist.ReduceOp.SUM

    # -- memory / topology queries --
    def get_mem_size(self, collectiveArgs, pair=False, pairIdx=0):
        t = collectiveArgs.opTensor if not pair else collectiveArgs.opTensor_pair[pairIdx]
        if isinstance(t, list):
            return sum(x.nelement() * x.element_size() for x in t)
        return t.nelement() * t.element_size()

    def getBusBW(self, collective, algBW, collectiveArgs):
        """algBW * (W-1)/W for the all_to_all family (pytorch_backend_utils.py:200-247)."""
        w = getattr(collectiveArgs, "world_size", 0) or self.get_world_size()
        if collective == "all_reduce":
            return algBW * 2 * (w - 1) / w
        if "all_to_all" in collective or collective in ("gather", "all_gather", "reduce_scatter",
                                                        "reduce_scatter_base", "scatter", "all_gather_base"):
            return algBW * (w - 1) / w
        return algBW

    def tensor_list_to_numpy(self, tensorList):
        if isinstance(tensorList, list):
            tensorList = [t.cpu().detach().numpy() for t in tensorList]
        return np.array(tensorList)

    def get_local_rank(self):
        return self.bootstrap_info.local_rank

    def get_local_size(self):
        return self.bootstrap_info.local_size

    def get_global_rank(self):
        return dist.get_rank()

    def get_world_size(self):
        return dist.get_world_size()

    def get_group_rank(self, group):
        return dist.get_rank(group)

    def get_group_size(self, group):
        return dist.get_world_size(group)

    def get_device(self):
        if _param(self.commsParams, "device", "cuda") != "cuda":
            raise PB200Error("the b200 backend runs on CUDA devices only")
        ordinal = self.get_local_rank()
        return torch.device(f"cuda:{ordinal if ordinal >= 0 else 0}")

    def get_hw_device(self):
        return self.get_device()

    def get_default_group(self):
        return dist.GroupMember.WORLD

    def get_groups(self):
        return self.groups

    def get_num_pgs(self):
        return self.num_pgs

    def get_next_group(self):
        return next(self.round_robin_group)

    def set_device(self, local_rank, global_rank):
        if local_rank >= torch.cuda.device_count():
            raise ValueError(f"Insufficient #GPUs: available {torch.cuda.device_count()} requested {local_rank}")
        torch.cuda.set_device(local_rank)

    def get_new_stream(self):
        return torch.cuda.Stream(device=self.get_device(), priority=0)

    def get_new_event(self, enable_timing=False):
        return torch.cuda.Event(enable_timing)

    def get_current_stream(self, device=None):
        return torch.cuda.current_stream(device)

    def switch_stream(self, stream, device=None):
        if stream is None:
            return None
        cur = torch.cuda.current_stream(device or self.get_device())
        torch.cuda.set_stream(stream)
        return cur

    def sync_stream(self, stream=None, device=None):
        (stream or torch.cuda.current_stream(device or self.get_device())).synchronize()
