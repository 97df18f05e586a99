"""Peer-mapped HBM windows + the C-ABI all-to-all communicator.

Every rank allocates one buffer [signal pad | data window]; all ranks map each other's buffers and
hand the raw addresses to libparam_b200 (pb200_a2a_comm_create).  Two ways to get the mapping, tried
in this order:
  1. torch.distributed._symmetric_memory (CUDA VMM + fd passing) — the API the reference already
     imports at train/comms/pt/comms_utils.py:22 and uses in pytorch_nvshmem_backend.py:27-40;
  2. classic CUDA IPC handles (cudaIpcGetMemHandle through torch's storage sharing), exchanged with
     all_gather_object.
PyTorch is plumbing here (allocation, rendezvous, streams); the data path is the push kernel.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from ... import _cabi
from ..._cabi import PB200Error

SIGNAL_BYTES = _cabi.A2A_SIGNAL_BYTES
_ALIGN = 512


class _LocalGroup:
    """W virtual ranks inside ONE process on ONE GPU (each with its own window and stream).  Lets the
    single-GPU test box exercise the full ready/push/done protocol: the W kernels run concurrently
    on W streams and signal each other through ordinary device memory."""

    def __init__(self, world: int, window_bytes: int, device, max_ctas: int = 4,
                 spin_timeout_s: float = 5.0):
        self.world, self.device = world, device
        self.buffers = [torch.zeros(SIGNAL_BYTES + window_bytes, dtype=torch.uint8, device=device)
                        for _ in range(world)]
        self.streams = [torch.cuda.Stream(device=device) for _ in range(world)]
        torch.cuda.synchronize(device)
        self.windows = [PeerWindow._from_pointers(r, world, [b.data_ptr() for b in self.buffers],
                                                  window_bytes, device, keepalive=self.buffers)
                        for r in range(world)]
        for w in self.windows:
            w.configure(max_ctas=max_ctas, spin_timeout_s=spin_timeout_s)


class PeerWindow:
    """One rank's view of the W peer-mapped buffers."""

    def __init__(self):
        raise PB200Error("use PeerWindow.create(group, window_bytes) or PeerWindow.local_group(...)")

    # ---- construction -----------------------------------------------------------------------
    @classmethod
    def _from_pointers(cls, rank, world, base_ptrs: Sequence[int], window_bytes, device, keepalive):
        self = object.__new__(cls)
        self.rank, self.world, self.device = rank, world, device
        self.window_bytes = int(window_bytes)
        self._keepalive = keepalive
        self._base_ptrs = list(base_ptrs)
        data = (C.c_void_p * world)(*[p + SIGNAL_BYTES for p in base_ptrs])
        sig = (C.c_void_p * world)(*base_ptrs)
        comm = C.c_void_p()
        with torch.cuda.device(device):
            _cabi.check(_cabi.load().pb200_a2a_comm_create(C.byref(comm), rank, world, data, sig,
                                                           self.window_bytes), "pb200_a2a_comm_create")
        self._comm = comm
        self._bump = 0
        return self

    @classmethod
    def create(cls, group, window_bytes: int, device: Optional[torch.device] = None) -> "PeerWindow":
        """Collective over `group` (a c10d process group on one NVSwitch node)."""
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if world > _cabi.A2A_MAX_RANKS:
            raise PB200Error(f"at most {_cabi.A2A_MAX_RANKS} ranks per window")
        total = SIGNAL_BYTES + int(window_bytes)
        total = (total + _ALIGN - 1) // _ALIGN * _ALIGN
        mode = os.environ.get("PB200_PEER_MAP", "auto")
        errors = []
        if mode in ("auto", "symm_mem"):
            try:
                return cls._create_symm_mem(group, rank, world, total, window_bytes, device)
            except Exception as exc:  # noqa: BLE001
                errors.append(f"symmetric memory: {exc!r}")
                if mode == "symm_mem":
                    raise
        try:
            return cls._create_cuda_ipc(group, rank, world, total, window_bytes, device)
        except Exception as exc:  # noqa: BLE001
            errors.append(f"cuda ipc: {exc!r}")
        raise PB200Error("could not map peer memory: " + "; ".join(errors))

    @classmethod
    def _create_symm_mem(cls, group, rank, world, total, window_bytes, device):
        import torch.distributed._symmetric_memory as symm_mem
        buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        hdl = symm_mem.rendezvous(buf, group)
        buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        self = cls._from_pointers(rank, world, ptrs, window_bytes, device, keepalive=(buf, hdl))
        self.mapping = "symmetric_memory"
        return self

    @classmethod
    def _create_cuda_ipc(cls, group, rank, world, total, window_bytes, device):
        # a dedicated cudaMalloc allocation (not a slice of a caching-allocator block shared with
        # other tensors): allocate through a private pool so the IPC handle covers exactly this
        pool = torch.cuda.MemPool()
        with torch.cuda.use_mem_pool(pool, device=device):
            buf = torch.zeros(total, dtype=torch.uint8, device=device)
        torch.cuda.synchronize(device)
        handle = buf.untyped_storage()._share_cuda_()
        gathered = [None] * world
        dist.all_gather_object(gathered, handle, group=group)
        peers, ptrs = [], []
        for r in range(world):
            if r == rank:
                peers.append(buf)
                ptrs.append(buf.data_ptr())
            else:
                # the alias lives on the device ordinal the storage reports (the sender's), not on ours:
                # set_() refuses to cross devices; the kernels only need the address, and peer access
                # between the two ordinals is what makes it dereferenceable from this rank's GPU
                st = torch.UntypedStorage._new_shared_cuda(*gathered[r])
                t = torch.empty(0, dtype=torch.uint8, device=st.device).set_(st)
                if st.device != device and st.device.type == "cuda":
                    if not torch.cuda.can_device_access_peer(device.index, st.device.index):
                        raise PB200Error(f"no peer access from {device} to {st.device}")
                peers.append(t)
                ptrs.append(t.data_ptr())
        dist.barrier(group)
        self = cls._from_pointers(rank, world, ptrs, window_bytes, device, keepalive=(pool, peers))
        self.mapping = "cuda_ipc"
        return self

    @staticmethod
    def local_group(world: int, window_bytes: int, device, max_ctas: int = 4,
                    spin_timeout_s: float = 5.0) -> _LocalGroup:
        return _LocalGroup(world, window_bytes, device, max_ctas, spin_timeout_s)

    # ---- window carving ----------------------------------------------------------------------
    def local_ptr(self, offset: int = 0) -> int:
        return self._base_ptrs[self.rank] + SIGNAL_BYTES + offset

    def view(self, offset: int, numel: int, dtype: torch.dtype) -> torch.Tensor:
        """Tensor aliasing [offset, offset + numel*itemsize) of THIS rank's window."""
        nbytes = numel * torch.empty(0, dtype=dtype).element_size()
        if offset < 0 or offset + nbytes > self.window_bytes:
            raise PB200Error("view outside the window")
        base = self._local_tensor()
        return base[SIGNAL_BYTES + offset: SIGNAL_BYTES + offset + nbytes].view(dtype)

    def _local_tensor(self) -> torch.Tensor:
        k = self._keepalive
        if isinstance(k, list):            # _LocalGroup buffers
            return k[self.rank]
        if isinstance(k[0], torch.Tensor):  # symm_mem (buf, hdl)
            return k[0]
        return k[1][self.rank]             # cuda_ipc (pool, peers)

    def alloc(self, numel: int, dtype: torch.dtype):
        """Bump-allocate a tensor inside the window (same offset on every rank when called in the
        same order everywhere).  Returns (tensor, byte offset)."""
        es = torch.empty(0, dtype=dtype).element_size()
        off = (self._bump + _ALIGN - 1) // _ALIGN * _ALIGN
        if off + numel * es > self.window_bytes:
            raise PB200Error(f"window exhausted: need {numel * es} at {off}, have {self.window_bytes}")
        self._bump = off + numel * es
        return self.view(off, numel, dtype), off

    def reset_alloc(self) -> None:
        self._bump = 0

    def offset_of(self, t: torch.Tensor) -> Optional[int]:
        """Byte offset of tensor t inside this rank's window, or None if it lives elsewhere."""
        lo = self._base_ptrs[self.rank] + SIGNAL_BYTES
        p = t.data_ptr()
        if lo <= p and p + t.numel() * t.element_size() <= lo + self.window_bytes:
            return p - lo
        return None

    # ---- collectives ------------------------------------------------------------------------
    def configure(self, max_ctas: int = -1, spin_timeout_s: float = 0.0) -> None:
        _cabi.check(_cabi.load().pb200_a2a_comm_config(self._comm, max_ctas, spin_timeout_s))

    def error(self) -> int:
        v = C.c_int32(0)
        _cabi.check(_cabi.load().pb200_a2a_comm_error(self._comm, C.byref(v)))
        return int(v.value)

    def _stream(self, stream) -> int:
        return (stream or torch.cuda.current_stream(self.device)).cuda_stream

    def _staging_offset(self, nbytes: int) -> int:
        """Where a result that has no home in the window is staged: the TOP of the window, growing
        downwards, never overlapping what the bump allocator has handed out (live comm buffers grow
        upwards from 0).  Raises instead of overwriting a live tensor."""
        off = (self.window_bytes - int(nbytes)) // _ALIGN * _ALIGN
        if off < self._bump:
            raise PB200Error(
                f"all-to-all result of {nbytes} B does not fit the peer window beside the {self._bump} B of "
                f"buffers allocated in it (window {self.window_bytes} B): allocate the output with "
                "alloc()/alloc_empty so that peers write it in place, or size the window up front "
                "(PB200_WINDOW_BYTES)")
        return off

    @staticmethod
    def _dim0_elems(t: torch.Tensor, splits):
        """c10d splits count rows along dim 0; the kernel takes elements of the flattened tensor."""
        if t.dim() <= 1 or t.shape[0] == 0:
            return [int(v) for v in splits]
        row = t.numel() // t.shape[0]
        return [int(v) * row for v in splits]

    def all_to_all_single(self, out: Optional[torch.Tensor], inp: torch.Tensor,
                          out_splits: Optional[Sequence[int]] = None,
                          in_splits: Optional[Sequence[int]] = None,
                          out_window_off: Optional[int] = None, stream=None) -> torch.Tensor:
        """c10d all_to_all_single semantics (splits along dim 0, as c10d counts them).  If `out`
        already lives inside the window the peers write it in place (zero copy); otherwise the result is
        staged — at out_window_off if the caller carved the window itself, else at the top of the
        window, clear of every tensor alloc() has handed out — and copied into `out` on the stream.
        With out=None the staged view is returned; it is valid until the next staged call."""
        if not inp.is_cuda or not inp.is_contiguous():
            raise PB200Error("all_to_all_single needs a contiguous CUDA input")
        es = inp.element_size()
        W = self.world
        have_in = in_splits is not None and len(in_splits) > 0
        have_out = out_splits is not None and len(out_splits) > 0
        if have_in:
            in_splits = self._dim0_elems(inp, in_splits)
        if have_out:
            out_splits = self._dim0_elems(out if out is not None else inp, out_splits)
        if have_in != have_out:
            # c10d allows giving only one side; the missing side is the equal split
            n_in, n_out = inp.numel(), (out.numel() if out is not None else inp.numel())
            if (not have_in and n_in % W) or (not have_out and n_out % W):
                raise PB200Error("equal-split side of all_to_all_single needs numel % world_size == 0")
            in_splits = list(in_splits) if have_in else [n_in // W] * W
            out_splits = list(out_splits) if have_out else [n_out // W] * W
            have_in = have_out = True
        if have_in:
            if len(in_splits) != W or len(out_splits) != W:
                raise PB200Error("split lists must have world_size entries")
            if sum(in_splits) != inp.numel():
                raise PB200Error(f"input splits sum to {sum(in_splits)} elements, the input has {inp.numel()}")
            total_out = sum(out_splits)
        else:
            if inp.numel() % W:
                raise PB200Error("equal-split all_to_all_single needs numel % world_size == 0")
            total_out = inp.numel()
        copy_out = None
        if out is not None:
            if out.dtype != inp.dtype or not out.is_contiguous():
                raise PB200Error("out must be contiguous with the input dtype")
            if out.numel() < total_out:
                raise PB200Error(f"out has {out.numel()} elements, the splits deliver {total_out}")
            off = self.offset_of(out)
            if off is not None:
                out_window_off = off
            else:
                copy_out = out
        if out_window_off is None:
            out_window_off = self._staging_offset(total_out * es)
        if out_window_off + total_out * es > self.window_bytes:
            raise PB200Error("all-to-all result does not fit the window at the requested offset")
        lib = _cabi.load()
        isb = osb = None
        if have_in:
            isb = _cabi.i64_array([s_ * es for s_ in in_splits])
            osb = _cabi.i64_array([s_ * es for s_ in out_splits])
        rc = lib.pb200_a2a_single(self._comm, inp.data_ptr(), inp.numel() * es, isb, osb,
                                  int(out_window_off), None if copy_out is None else copy_out.data_ptr(),
                                  self._stream(stream))
        _cabi.check(rc, "pb200_a2a_single")
        if out is not None:
            return out
        return self.view(out_window_off, total_out, inp.dtype)

    def all_to_all(self, outs: Sequence[torch.Tensor], ins: Sequence[torch.Tensor], stream=None) -> None:
        """dist.all_to_all(output_tensor_list, input_tensor_list): ins[j] goes to rank j, outs[r] receives
        from rank r (pb200_a2a_list).  Output tensors that live in the window are written in place by
        the peers; the others are staged at the top of the window and copied out on the stream.  No
        cat before, no split after."""
        W = self.world
        if len(outs) != W or len(ins) != W:
            raise PB200Error("all_to_all needs world_size input and output tensors")
        for t in list(outs) + list(ins):
            if not t.is_cuda or not t.is_contiguous():
                raise PB200Error("all_to_all needs contiguous CUDA tensors")
        nb_out = [t.numel() * t.element_size() for t in outs]
        offs, copies = [], []
        staged = sum((b + _ALIGN - 1) // _ALIGN * _ALIGN for t, b in zip(outs, nb_out) if self.offset_of(t) is None)
        cursor = self._staging_offset(staged) if staged else 0
        for t, b in zip(outs, nb_out):
            off = self.offset_of(t)
            if off is None:
                off = cursor
                cursor += (b + _ALIGN - 1) // _ALIGN * _ALIGN
                copies.append(t.data_ptr())
            else:
                copies.append(None)
            offs.append(off)
        in_ptrs = (C.c_void_p * W)(*[t.data_ptr() if t.numel() else None for t in ins])
        out_copy = (C.c_void_p * W)(*copies) if any(c is not None for c in copies) else None
        rc = _cabi.load().pb200_a2a_list(self._comm, in_ptrs,
                                         _cabi.i64_array([t.numel() * t.element_size() for t in ins]),
                                         _cabi.i64_array(offs), _cabi.i64_array(nb_out), out_copy,
                                         self._stream(stream))
        _cabi.check(rc, "pb200_a2a_list")

    def pooled_forward(self, pooled: torch.Tensor, batch_split: Sequence[int],
                       tables_split: Sequence[int], emb_dim: int, layout: str = "BTD",
                       out_window_off: int = 0, stream=None) -> torch.Tensor:
        """Fused DLRM forward exchange + output permute (dlrm.py:86-134,157-177,1253).
        pooled: this rank's lookups for the GLOBAL batch, [N, T_local*E] ("BTD") or [T_local, N, E]
        ("TBD").  Returns this rank's [lN, T_global*E] tensor (a view of its window)."""
        T_l, N = int(tables_split[self.rank]), int(sum(batch_split))
        E = int(emb_dim)
        if pooled.dtype != torch.float32 or not pooled.is_contiguous() or pooled.numel() != T_l * N * E:
            raise PB200Error("pooled must be contiguous fp32 with T_local*N*E elements")
        st_t, st_n = (E, T_l * E) if layout == "BTD" else (N * E, E)
        rc = _cabi.load().pb200_a2a_pooled_fwd(self._comm, pooled.data_ptr(), st_t, st_n, E,
                                               _cabi.i64_array(batch_split), _cabi.i64_array(tables_split),
                                               int(out_window_off), self._stream(stream))
        _cabi.check(rc, "pb200_a2a_pooled_fwd")
        lN, Tg = int(batch_split[self.rank]), int(sum(tables_split))
        return self.view(out_window_off, lN * Tg * E, torch.float32).view(lN, Tg * E)

    def lookup_forward_fused(self, arena, indices: torch.Tensor, offsets: torch.Tensor,
                             batch_split: Sequence[int], tables_split: Sequence[int],
                             mode: str = "sum", out_window_off: int = 0, stream=None) -> torch.Tensor:
        """ONE kernel: batched lookup of this rank's tables for the global batch, pooled rows stored
        straight into every destination's [lN, T_global*E] tensor (pb200_tbe_fwd_a2a).  `arena` is an
        ops.TableArena with tables_split[rank] tables; offsets has T_local*N + 1 entries."""
        if not indices.is_cuda or indices.dtype != offsets.dtype:
            raise PB200Error("lookup_forward_fused needs CUDA indices/offsets of one integer dtype")
        it = {torch.int64: 0, torch.int32: 1}[indices.dtype]
        E, N = arena.dim, int(sum(batch_split))
        if offsets.numel() != arena.num_tables * N + 1:
            raise PB200Error("offsets must have T_local*N + 1 entries")
        rc = _cabi.load().pb200_tbe_fwd_a2a(
            self._comm, arena.weights.data_ptr(), arena.row_offsets.data_ptr(), arena.num_tables, E,
            indices.data_ptr(), indices.numel(), offsets.data_ptr(), it, {"sum": 0, "mean": 1}[mode],
            _cabi.i64_array(batch_split), _cabi.i64_array(tables_split), int(out_window_off),
            self._stream(stream))
        _cabi.check(rc, "pb200_tbe_fwd_a2a")
        lN, Tg = int(batch_split[self.rank]), int(sum(tables_split))
        return self.view(out_window_off, lN * Tg * E, torch.float32).view(lN, Tg * E)

    def pooled_backward(self, grad: torch.Tensor, batch_split: Sequence[int],
                        tables_split: Sequence[int], emb_dim: int, out_window_off: int = 0,
                        stream=None, part: int = 0, parts: int = 1) -> torch.Tensor:
        """Transpose exchange (dlrm.py:180-218,137-154): grad [lN, T_global*E] -> this rank's
        [N, T_local*E] gradient of its pooled lookups (a view of its window).  parts > 1: only the columns of
        every owner's tables of piece `part` move (pb200_a2a_pooled_bwd_part); the returned view is the same
        tensor, complete once all pieces have been exchanged."""
        lN, Tg, E = int(batch_split[self.rank]), int(sum(tables_split)), int(emb_dim)
        if grad.dtype != torch.float32 or not grad.is_contiguous() or grad.numel() != lN * Tg * E:
            raise PB200Error("grad must be contiguous fp32 [lN, T_global*E]")
        rc = _cabi.load().pb200_a2a_pooled_bwd_part(self._comm, grad.data_ptr(), E,
                                                    _cabi.i64_array(batch_split), _cabi.i64_array(tables_split),
                                                    int(out_window_off), int(part), int(parts), self._stream(stream))
        _cabi.check(rc, "pb200_a2a_pooled_bwd_part")
        N, T_l = int(sum(batch_split)), int(tables_split[self.rank])
        return self.view(out_window_off, N * T_l * E, torch.float32).view(N, T_l * E)

    def sparse_data_dist(self, lengths: torch.Tensor, indices: torch.Tensor, tables_split: Sequence[int],
                         local_batch: int, lengths_window_off: int, indices_window_off: int,
                         slot_elems: int, out=None, stream=None):
        """SparseDataDist (dlrm.py:744-855) as one asynchronous call with no host round trip
        (pb200_sparse_data_dist).  lengths int64 [T_global*b], indices int64 in the same order.
        Returns (lengths_out [T_l, W*b], offsets_out [T_l*W*b + 1], indices_out [W*slot_elems], of
        which the first offsets_out[-1] entries are valid).  `out`: optional preallocated triple."""
        if lengths.dtype != torch.int64 or indices.dtype != torch.int64 or not lengths.is_cuda:
            raise PB200Error("sparse_data_dist takes int64 CUDA lengths and indices")
        lengths, indices = lengths.contiguous().view(-1), indices.contiguous().view(-1)
        W, b, T_l = self.world, int(local_batch), int(tables_split[self.rank])
        if len(tables_split) != W or lengths.numel() != int(sum(tables_split)) * b:
            raise PB200Error("lengths must have T_global*b elements and tables_split one entry per rank")
        n = W * T_l * b
        if out is None:
            out = (torch.empty(n, dtype=torch.int64, device=self.device),
                   torch.empty(n + 1, dtype=torch.int64, device=self.device),
                   torch.empty(W * int(slot_elems), dtype=torch.int64, device=self.device))
        lengths_out, offsets_out, indices_out = out
        if lengths_out.numel() < n or offsets_out.numel() < n + 1 or indices_out.numel() < W * int(slot_elems):
            raise PB200Error("preallocated outputs are too small")
        lib = _cabi.load()
        sb = int(lib.pb200_regroup_scratch_bytes(W, T_l, b))
        # scratch is owned by the window (virtual ranks of a local group run concurrently on one GPU)
        scratch = getattr(self, "_dist_scratch", None)
        if scratch is None or scratch.numel() < sb:
            scratch = self._dist_scratch = torch.empty(sb, dtype=torch.uint8, device=self.device)
        rc = lib.pb200_sparse_data_dist(self._comm, lengths.data_ptr(), indices.data_ptr(), indices.numel(),
                                        _cabi.i64_array(tables_split), b, int(lengths_window_off),
                                        int(indices_window_off), int(slot_elems), lengths_out.data_ptr(),
                                        offsets_out.data_ptr(), indices_out.data_ptr(), scratch.data_ptr(),
                                        sb, self._stream(stream))
        _cabi.check(rc, "pb200_sparse_data_dist")
        return lengths_out[:n].view(T_l, W * b), offsets_out[:n + 1], indices_out

    def close(self) -> None:
        if getattr(self, "_comm", None):
            _cabi.load().pb200_a2a_comm_destroy(self._comm)
            self._comm = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
