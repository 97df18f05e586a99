"""DLRM communication pattern on the B200 kernels (mirror of train/comms/pt/dlrm.py).

Table-parallel embeddings + batch-parallel dense part, joined by an all-to-all:

  reference step (dlrm.py:1200-1323)                     here
  ------------------------------------------------------ ------------------------------------------
  SparseFeatures/calculateLengths (:226-277)             SparseBatch (lengths kept on the device)
  SparseDataDist: lengths a2a, .item() sync, indices     sparse_data_dist(): pb200_sparse_data_dist — lengths
    a2a, python splitPerTable (:744-855, :430-504)         push, index counts kept on the device, index push
                                                            into fixed slots, device regroup; no host sync
  apply_emb: T_l nn.EmbeddingBag launches + stack (:363) one batched TBE launch writing [N, T_l*E]
  All2Allv_Req/Wait + 2 torch.cat (:86-218, :1253)       pb200_a2a_pooled_fwd: pooled rows pushed straight
                                                            into the peers' final [lN, T_g*E] tensors
  tempB.backward(C) -> a2a bwd + cat/split +             pb200_a2a_pooled_bwd + pb200_tbe_bwd scatter-add
    EmbeddingBagBackward (:1296, :180-218, :137-154)
  MLP-gradient all_reduce (:1266-1281, :1303-1317)       the runner's two stages: async dist.all_reduce (NCCL) per
                                                            layer tensor, closed by a barrier — pass-through
  timers / report (:880-1009, :1011-1193)                 MARKS / REGIONS: the same 21 rows, same table layout

Run:  torchrun --nproc-per-node 8 -m param_b200.comms.pt.dlrm --mini-batch-size 8192 \
          --arch-embedding-size 1000000x512 --arch-sparse-feature-size 128 --num-indices-per-lookup 20
"""
from __future__ import annotations

import argparse
import json
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from ... import ops
from ..._cabi import PB200Error
from .peer_window import PeerWindow


def split_lengths(n: int, world: int) -> List[int]:
    """Contiguous block split with the remainder on the low ranks (get_split_lengths_by_len,
    dlrm.py:390-398)."""
    k, m = divmod(int(n), int(world))
    return [k + 1 if r < m else k for r in range(world)]


def owner_slice(rank: int, split: Sequence[int]) -> slice:
    """Tables owned by `rank` (get_slice_sparse, dlrm.py:424-428)."""
    lo = sum(split[:rank])
    return slice(lo, lo + split[rank])


def parse_embedding_sizes(spec: str) -> List[int]:
    """'1000-1000-2000' (reference syntax, dlrm.py:509) or the shorthand '1000000x512'."""
    if "x" in spec:
        rows, count = spec.split("x")
        return [int(rows)] * int(count)
    return [int(v) for v in spec.split("-") if v]


def lengths_exchange_splits(tables_split: Sequence[int], rank: int, local_batch: int):
    """all_to_all splits (in ELEMENTS) of the lengths exchange: rank r receives the lengths of ITS
    tables from everyone — in-splits T_j*b per destination j, out-splits T_rank*b per source
    (dlrm.py:768-785)."""
    in_splits = [int(t) * int(local_batch) for t in tables_split]
    out_splits = [int(tables_split[rank]) * int(local_batch)] * len(tables_split)
    return in_splits, out_splits


def indices_exchange_counts(lengths: torch.Tensor, lengths_out: torch.Tensor,
                            tables_split: Sequence[int], local_batch: int):
    """Index counts per destination (sum of my lengths over the tables each rank owns) and per
    source (sum of the received lengths of each source block) — the splits of the indices
    all-to-all (dlrm.py:801-818).  Device tensors int64 [W] each; no host sync here."""
    W = len(tables_split)
    per_table = lengths.view(-1, int(local_batch)).sum(dim=1)
    bounds = torch.tensor([0] + list(torch.tensor(list(tables_split)).cumsum(0).tolist()),
                          device=lengths.device)
    csum = torch.cat([per_table.new_zeros(1), per_table.cumsum(0)])
    send_counts = csum[bounds[1:]] - csum[bounds[:-1]]
    recv_counts = lengths_out.view(W, -1).sum(dim=1)
    return send_counts, recv_counts


@dataclass
class SparseBatch:
    """One rank's sparse inputs for its LOCAL batch and ALL global tables (table-major), the
    interface SparseDataDist expects (dlrm.py:254-277)."""
    count: int                 # T_global
    batch_size: int            # local batch b
    lengths: torch.Tensor      # int64 [T_global * b]
    indices: torch.Tensor      # int64 [sum(lengths)], table-major then sample order

    @staticmethod
    def from_offsets(offsets: Sequence[torch.Tensor], indices: Sequence[torch.Tensor], device) -> "SparseBatch":
        """Per-table nn.EmbeddingBag offsets/indices -> lengths (calculateLengths, dlrm.py:226-242:
        first differences with the last bag closed by len(indices))."""
        lens = []
        for off, idx in zip(offsets, indices):
            off = off.to(device=device, dtype=torch.int64)
            end = torch.tensor([idx.numel()], dtype=torch.int64, device=device)
            lens.append(torch.diff(torch.cat([off, end])))
        b = offsets[0].numel()
        return SparseBatch(len(offsets), b, torch.cat(lens),
                           torch.cat([i.to(device=device, dtype=torch.int64) for i in indices]))

    @staticmethod
    def synthetic(table_rows: Sequence[int], local_batch: int, bag: int, fixed: bool, seed: int,
                  device, alpha: float = 0.0) -> "SparseBatch":
        """Device-side generator.  fixed bag size (reference --num-indices-per-lookup-fixed) or ragged
        lengths uniform in [1, bag]; indices uniform (reference dlrm_data.py:182-183) or Zipf."""
        from ...compute.pt.pytorch_emb import zipf_cdf
        T, b = len(table_rows), int(local_batch)
        g = torch.Generator(device=device)
        g.manual_seed(int(seed))
        if fixed:
            lengths = torch.full((T * b,), int(bag), dtype=torch.int64, device=device)
        else:
            lengths = torch.randint(1, int(bag) + 1, (T * b,), generator=g, device=device, dtype=torch.int64)
        per_table = lengths.view(T, b).sum(dim=1).tolist()
        chunks, cdf_cache = [], {}
        for t, rows in enumerate(table_rows):
            n = int(per_table[t])
            if alpha > 0:
                if rows not in cdf_cache:
                    cdf_cache[rows] = torch.from_numpy(zipf_cdf(alpha, rows)).to(device)
                buf = torch.empty(n, dtype=torch.int64, device=device)
                ops.fill_zipf_indices_(buf, 1, cdf_cache[rows], seed * 1315423911 + t, dedupe=False)
                chunks.append(buf)
            else:
                chunks.append(torch.randint(0, int(rows), (n,), generator=g, device=device, dtype=torch.int64))
        return SparseBatch(T, b, lengths, torch.cat(chunks) if chunks else lengths.new_empty(0))


class DLRMParallelEmbedding:
    """The table-parallel half of one DLRM iteration on one rank."""

    def __init__(self, group, table_rows: Sequence[int], emb_dim: int, local_batch: int,
                 max_bag: int, device: torch.device, lr: float = 0.0, seed: int = 0,
                 window: Optional[PeerWindow] = None, bwd_algo: str = "sorted"):
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.table_rows = [int(r) for r in table_rows]
        self.T_global, self.E = len(self.table_rows), int(emb_dim)
        if self.T_global < self.world:
            raise PB200Error("need at least one table per rank (dlrm.py:511-514)")
        self.tables_split = split_lengths(self.T_global, self.world)
        self.batch_split = [int(local_batch)] * self.world       # a2ai.gNS: equal local batches
        self.b, self.N = int(local_batch), int(local_batch) * self.world
        self.my_tables = owner_slice(self.rank, self.tables_split)
        self.T_local = self.tables_split[self.rank]
        self.lr, self.bwd_algo = float(lr), bwd_algo
        rows = self.table_rows[self.my_tables]
        self.arena = ops.TableArena.allocate(rows, self.E, device)
        for t, n in enumerate(rows):  # U(+-sqrt(1/n)) init, alloc_embedding_tables pytorch_dist_backend.py:926-928
            ops.fill_uniform_(self.arena.table(t), -(1.0 / n) ** 0.5, (1.0 / n) ** 0.5,
                              seed=seed * 7919 + self.my_tables.start + t)
        # window carve-up (identical on every rank: sized with the max over ranks)
        T_max = max(self.tables_split)
        self.cap_lengths = self.world * T_max * self.b
        self.cap_indices = self.cap_lengths * int(max_bag)
        need = (self.cap_lengths + self.cap_indices) * 8 + self.b * self.T_global * self.E * 4 \
            + self.N * T_max * self.E * 4 + 4 * 512
        self.window = window or PeerWindow.create(group, need, device)
        if self.window.window_bytes < need:
            raise PB200Error(f"window too small: {self.window.window_bytes} < {need}")
        self.window.reset_alloc()
        _, self.off_lengths = self.window.alloc(self.cap_lengths, torch.int64)
        _, self.off_indices = self.window.alloc(self.cap_indices, torch.int64)
        _, self.off_pooled = self.window.alloc(self.b * self.T_global * self.E, torch.float32)
        _, self.off_grad = self.window.alloc(self.N * T_max * self.E, torch.float32)
        self._pooled_local = torch.empty((self.N, self.T_local * self.E), dtype=torch.float32, device=device)
        self._saved = None
        self.fused = True
        # device-side redistribution: every source gets a fixed slot of the index window
        self.slot_elems = self.cap_indices // self.world
        n_len = self.world * self.T_local * self.b
        self._dist_out = (torch.empty(n_len, dtype=torch.int64, device=device),
                          torch.empty(n_len + 1, dtype=torch.int64, device=device),
                          torch.empty(self.cap_indices, dtype=torch.int64, device=device))
        self.device_side_dist = True
        # the backward's sort needs the indices only: it is queued on a side stream at forward time
        # (ops.tbe_plan) and the backward after the gradient exchange is the segmented reduce alone
        self.presort = bwd_algo in ("sorted", "exact", "auto")
        self._side = torch.cuda.Stream(device=device) if self.presort else None
        self._plan, self._plan_buf = None, None
        self.max_table_rows = max(rows) if rows else 0
        # backward in pieces: the transpose exchange of table group g + 1 (comm stream, grid capped at 32 CTAs so that
        # it leaves the SMs to the reduce — csrc/a2a.cu, PB200_A2A_PART_CTAS) runs under the segmented reduce of group
        # g.  Measured (tools/pieces_bench.py, profiles/r02r_pieces_n*.log): N = 8 step 11.9 -> 10.7 ms with 4 parts
        # (10.6 with 8), N = 4 5.42 -> 4.96 ms with 4 parts; uncapped pushes gain nothing (5.49 ms: a full-grid push
        # leaves room for one reduce CTA per SM instead of three); parts that do not divide the tables evenly are
        # slower than no pieces at all (3 / 5 / 6 parts of 64 tables: 6.9 / 6.0 / 6.0 ms).  At N = 2 the exchange
        # (0.48 ms) is shorter than what a second epoch + launch costs (3.01 vs 2.86 ms) -> one piece there.
        # PB200_DLRM_BWD_PARTS overrides.
        t_min = min(self.tables_split)
        default_parts = 1
        if self.world >= 4:
            # a part must keep the push kernel's row loop busy: it copies one gradient row segment of
            # (tables of the part) x E x 4 bytes per pass of 512 threads x 16 B = 8 KB (8 parts of 64 tables = 4 KB
            # segments were slower than 4 parts at N = 4)
            for cand in (4, 2):
                if all(t % cand == 0 and (t // cand) * self.E * 4 >= 8192 for t in self.tables_split):
                    default_parts = cand
                    break
        self.bwd_parts = max(1, min(int(os.environ.get("PB200_DLRM_BWD_PARTS", default_parts)), t_min))
        self._comm_stream = torch.cuda.Stream(device=device) if self.bwd_parts > 1 else None

    # ---- step 2: SparseDataDist ---------------------------------------------------------------
    def sparse_data_dist(self, batch: SparseBatch, device_side: Optional[bool] = None, mark=None):
        """batch-parallel -> table-parallel redistribution of (lengths, indices); returns the TBE
        request (offsets int64[T_l*N+1], indices) for this rank's tables over the GLOBAL batch.
        device_side=True (default): one asynchronous call, index counts never leave the device
        (the returned indices tensor then has the window's capacity; offsets[-1] entries are valid).
        device_side=False: the two-collective form with ONE [2, W]-count D2H in between; `mark(name)` is
        then called at the reference's offset_xchg_end / idx_xchg_start / idx_xchg_end stamps (dlrm.py:786-830)."""
        mark = mark or (lambda name: None)
        W, b, win = self.world, self.b, self.window
        if batch.count != self.T_global or batch.batch_size != b:
            raise PB200Error("SparseBatch does not match the configured tables / local batch")
        if self.device_side_dist if device_side is None else device_side:
            if batch.indices.numel() > self.T_global * b * (self.cap_indices // self.cap_lengths):
                raise PB200Error("more indices than the window was sized for (raise max_bag)")
            _, offsets, indices = win.sparse_data_dist(batch.lengths, batch.indices, self.tables_split, b,
                                                       self.off_lengths, self.off_indices, self.slot_elems,
                                                       out=self._dist_out)
            return offsets, indices
        in_splits, out_splits = lengths_exchange_splits(self.tables_split, self.rank, b)
        lengths_out = win.all_to_all_single(None, batch.lengths, out_splits, in_splits,
                                            out_window_off=self.off_lengths)
        mark("offset_xchg_end")
        # element counts per destination / per source: the one host round trip (the reference has
        # .item() + two .numpy() here, dlrm.py:801-818)
        send_counts, recv_counts = indices_exchange_counts(batch.lengths, lengths_out,
                                                           self.tables_split, b)
        counts = torch.stack([send_counts, recv_counts]).cpu()
        idx_in, idx_out = counts[0].tolist(), counts[1].tolist()
        n_recv = int(sum(idx_out))
        if n_recv > self.cap_indices:
            raise PB200Error("received more indices than the window was sized for (raise max_bag)")
        mark("idx_xchg_start")
        indices_out = win.all_to_all_single(None, batch.indices, idx_out, idx_in,
                                            out_window_off=self.off_indices)
        mark("idx_xchg_end")
        # per-table regroup + offsets (splitPerTable / lengthsToOffsets, dlrm.py:430-504, 245-251)
        _, offsets, indices = ops.regroup_sparse(lengths_out, indices_out[:n_recv], W, self.T_local, b)
        return offsets, indices

    # ---- steps 3+4: apply_emb + forward all-to-all ------------------------------------------------
    def forward(self, offsets: torch.Tensor, indices: torch.Tensor, fused: Optional[bool] = None,
                before_exchange=None) -> torch.Tensor:
        """lookup for the global batch + exchange; returns [lN, T_global*E].  fused=True (default):
        ONE kernel whose epilogue stores pooled rows into the peers' windows (pb200_tbe_fwd_a2a);
        fused=False: lookup kernel to a local [N, T_l*E] buffer, then the push kernel.
        before_exchange: called where the reference stamps fwd_a2a_start (after apply_emb is queued, before the
        all-to-all, dlrm.py:1239-1253) — with the fused kernel that is before the one launch."""
        self._saved = (offsets, indices)
        self._plan = None
        if self.presort:
            self._plan = ops.tbe_plan(self.arena.row_offsets, self.T_local, self.E, indices, offsets, self.N,
                                      self.max_table_rows, layout="BTD", exact=self.bwd_algo == "exact",
                                      stream=self._side, buf=self._plan_buf)
            self._plan_buf = self._plan.buf
        if self.fused if fused is None else fused:
            if before_exchange is not None:
                before_exchange()
            return self.window.lookup_forward_fused(self.arena, indices, offsets, self.batch_split,
                                                    self.tables_split, out_window_off=self.off_pooled)
        ops.tbe_forward(self.arena, indices, offsets, self.N, layout="BTD", out=self._pooled_local)
        if before_exchange is not None:
            before_exchange()
        return self.window.pooled_forward(self._pooled_local, self.batch_split, self.tables_split, self.E,
                                          layout="BTD", out_window_off=self.off_pooled)

    # ---- step 6: backward all-to-all + scatter-add ---------------------------------------------------
    def backward(self, grad: torch.Tensor) -> None:
        offsets, indices = self._saved
        grad = grad.contiguous()
        if self.bwd_parts > 1 and self._plan is not None and not self._plan.exact:
            # piece g of the exchange on the comm stream, its reduce on the caller's stream as soon as it has
            # landed: only the first piece of the exchange is exposed
            cur = torch.cuda.current_stream(self.device)
            self._comm_stream.wait_stream(cur)                    # the gradient is ready on `cur`
            grad.record_stream(self._comm_stream)
            landed = []
            for g in range(self.bwd_parts):
                g_local = self.window.pooled_backward(grad, self.batch_split, self.tables_split, self.E,
                                                      out_window_off=self.off_grad, stream=self._comm_stream,
                                                      part=g, parts=self.bwd_parts)
                ev = torch.cuda.Event()
                ev.record(self._comm_stream)
                landed.append(ev)
            for g in range(self.bwd_parts):
                cur.wait_event(landed[g])
                lo, hi = ops.part_range(self.T_local, g, self.bwd_parts)
                ops.tbe_backward_tables(self.arena.weights, self.arena.row_offsets, self.T_local, self.E, indices,
                                        offsets, self.N, g_local, self._plan, lo, hi, layout="BTD", scale=-self.lr)
            return
        g_local = self.window.pooled_backward(grad, self.batch_split, self.tables_split,
                                              self.E, out_window_off=self.off_grad)
        ops.tbe_backward(self.arena.weights, self.arena.row_offsets, self.T_local, self.E, indices,
                         offsets, self.N, g_local, layout="BTD", scale=-self.lr, algo=self.bwd_algo,
                         max_table_rows=self.max_table_rows, plan=self._plan)


class NCCLReferenceEmbedding:
    """Comparator: the same step the way the reference does it on the same box — per-table
    torch.nn.EmbeddingBag, torch.stack, cat + dist.all_to_all_single (NCCL) + split/cat
    (dlrm.py:363-388, 86-218, 1253).  Used only to print a baseline beside the B200 numbers."""

    def __init__(self, group, peer: DLRMParallelEmbedding):
        self.group, self.p = group, peer
        import torch.nn as nn
        self.embs = []
        for t in range(peer.T_local):
            e = nn.EmbeddingBag(peer.arena.rows[t], peer.E, mode="sum", sparse=True,
                                _weight=peer.arena.table(t).clone())
            self.embs.append(e.to(peer.device))

    def forward(self, offsets, indices):
        p = self.p
        N, E, W = p.N, p.E, p.world
        ly = []
        for t, e in enumerate(self.embs):
            lo, hi = offsets[t * N], offsets[(t + 1) * N]
            ly.append(e(indices[lo:hi], offsets[t * N:(t + 1) * N] - lo))
        ly = torch.stack(ly)                                            # [T_l, N, E]
        inp = torch.cat(list(ly), dim=1).view(-1)                       # [N, T_l*E] flattened
        in_splits = [m * p.T_local * E for m in p.batch_split]
        out_splits = [p.b * t * E for t in p.tables_split]
        out = inp.new_empty(sum(out_splits))
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=self.group)
        parts = [o.view(p.b, -1) for o in out.split(out_splits)]
        return torch.cat(parts, dim=1)                                  # [lN, T_g*E]


# ------------------------------------------------------------------------------------------------
# runner: the reference's iteration, marks and report table (dlrm.py:880-1009, 1011-1193, 1200-1323)
# ------------------------------------------------------------------------------------------------
# Host-clock marks of one iteration, in the order they are taken.  As in the reference they are wall-clock
# stamps; the device is drained only where the reference calls sync_barrier (before bef_emb_lookup, after
# grad_push_start, after each all-reduce stage) — or at every mark with --perf-debug.
MARKS = ("iter_start", "length_calc_end", "mem_push_idx_end", "offset_xchg_start", "offset_xchg_end",
         "idx_xchg_start", "idx_xchg_end", "bef_emb_lookup", "fwd_a2a_start", "fwd_a2a_end", "grad_push_start",
         "bwd_top_ar_start", "bwd_top_ar_end", "bwd_a2a_start", "bwd_a2a_end", "bwd_bot_ar_start", "bwd_bot_ar_end")

# (region, first mark, last mark) — the 21 rows of the reference's table, same names and order
# (initTimers dlrm.py:961-1009, all_timers dlrm.py:1015-1037)
REGIONS = (
    ("intermed_calc_length", "iter_start", "length_calc_end"),
    ("mem_push_idx", "length_calc_end", "mem_push_idx_end"),
    ("intermed_bef_offset_xchg", "mem_push_idx_end", "offset_xchg_start"),
    ("offset_xchg", "offset_xchg_start", "offset_xchg_end"),
    ("intermed_btw_offset_idx_xchg", "offset_xchg_end", "idx_xchg_start"),
    ("idx_xchg", "idx_xchg_start", "idx_xchg_end"),
    ("intermed_post_idx_xchg_sparse_dist", "idx_xchg_end", "bef_emb_lookup"),
    ("intermed_emb_lookup_to_a2a_start", "bef_emb_lookup", "fwd_a2a_start"),
    ("fwd_a2a", "fwd_a2a_start", "fwd_a2a_end"),
    ("intermed_fwd_a2a_grad_push", "fwd_a2a_end", "grad_push_start"),
    ("mem_push_gradients", "grad_push_start", "bwd_top_ar_start"),
    ("bwd_top_ar", "bwd_top_ar_start", "bwd_top_ar_end"),
    ("intermed_top_ar_end_to_bwd_a2a_start", "bwd_top_ar_end", "bwd_a2a_start"),
    ("bwd_a2a", "bwd_a2a_start", "bwd_a2a_end"),
    ("intermed_bwd_a2a_bot_ar", "bwd_a2a_end", "bwd_bot_ar_start"),
    ("bwd_bot_ar", "bwd_bot_ar_start", "bwd_bot_ar_end"),
    ("iter_time", "iter_start", "bwd_bot_ar_end"),
    ("iter_data_prep", "iter_start", "bef_emb_lookup"),
    ("iter_fwd_a2a", "iter_start", "grad_push_start"),
    ("iter_bwd_top_ar", "iter_start", "bwd_top_ar_end"),
    ("iter_bwd_a2a", "iter_start", "bwd_bot_ar_start"),
)
# regions that carry a message size in the table; every other row reports 0 (intermed_region_memory, dlrm.py:912-934)
SIZED_REGIONS = ("offset_xchg", "idx_xchg", "fwd_a2a", "bwd_top_ar", "bwd_a2a", "bwd_bot_ar")


def mlp_layer_shapes(spec: Sequence[int]) -> List[List[int]]:
    """'a-b-c' -> [[b, a], [c, b]]: one [out, in] weight per consecutive pair (create_mlp, dlrm.py:400-411)."""
    dims = [int(v) for v in spec]
    return [[dims[i + 1], dims[i]] for i in range(len(dims) - 1)]


def top_mlp_dims(n_tables: int, bot: Sequence[int], top: Sequence[int], interaction_op: str = "dot",
                 interaction_itself: bool = False, project_size: int = 0) -> List[int]:
    """Input width of the top MLP from the interaction (dlrm.py:575-604): with F = tables + 1 features and
    d = last bottom width, 'dot' gives F(F-1)/2 + d pairs (F(F+1)/2 + d with the diagonal), 'cat' gives F*d;
    a projection replaces it by F*project_size + d."""
    F, d = int(n_tables) + 1, int(bot[-1])
    if interaction_op == "dot":
        n_int = (F * (F + 1)) // 2 + d if interaction_itself else (F * (F - 1)) // 2 + d
    elif interaction_op == "cat":
        n_int = F * d
    else:
        raise PB200Error(f"--arch-interaction-op={interaction_op} is not supported")
    if project_size > 0:
        n_int = F * int(project_size) + d
    return [n_int] + [int(v) for v in top]


def region_times_us(marks: dict) -> dict:
    """One iteration's region latencies in microseconds from its marks (computeTimes, dlrm.py:952-959; the
    reference labels the same quantity 'nanoseconds' and prints it under 'Latency(us)')."""
    return {name: (marks[b] - marks[a]) * 1e6 for name, a, b in REGIONS}


def percentile_rows(lat: "torch.Tensor", mem: "torch.Tensor"):
    """lat, mem: [W, R, iters] gathered samples.  Returns (per_sample_rows, per_rank_mean_rows), each a list of
    (region, mem_p50, min, p50, p75, p95, running sum of p50 over the non-'iter' regions) — the two tables the
    reference prints (dlrm.py:1086-1160): percentiles over all W*iters samples, and over the W per-rank means.
    Single precision throughout, as there: the reference gathers its samples as float32 tensors and numpy keeps
    float32 through np.percentile and the running sums, which decides the last printed digit."""
    import numpy as np
    lat = lat.to(torch.float32)
    W, R, _ = lat.shape
    rows_all, rows_mean = [], []
    run_all, run_mean = np.float32(0.0), np.float32(0.0)
    for r in range(R):
        name = REGIONS[r][0]
        samples = lat[:, r, :].reshape(-1).numpy()
        means = np.array([lat[w, r, :].mean() for w in range(W)], dtype=np.float32)
        mem_p50 = float(np.percentile(mem[:, r, :].reshape(-1).numpy(), 50))
        pa = [np.percentile(samples, q) for q in (50, 75, 95)]
        pm = [np.percentile(means, q) for q in (50, 75, 95)]
        if "iter" not in name:
            run_all = np.float32(run_all + pa[0])
            run_mean = np.float32(run_mean + pm[0])
        rows_all.append((name, mem_p50, float(samples.min()), *(float(v) for v in pa), float(run_all)))
        rows_mean.append((name, mem_p50, float(means.min()), *(float(v) for v in pm), float(run_mean)))
    return rows_all, rows_mean


def format_report(iters: int, rows, header: bool = True) -> str:
    """The reference's table layout (dlrm.py:1069-1081, 1133-1177): tab-separated, a blank line before the
    iter_* rows, a total_time footer.  The reference prints the column header once, above the first table."""
    out = []
    if header:
        out.append("\t{}\t{:>36}\t{:>12}\t{:>12}\t{:>12}\t{:>12}\t{:>12}\t{:>12}".format(
            "iters", "region", "memory (B)", "Latency(us):min", "p50", "p75", "p95", "sum(p50)"))
    total = 0.0
    for name, mem, lo, p50, p75, p95, run in rows:
        if name == "iter_time":
            out.append("\n")
        out.append("\t%d\t%36s\t%12s\t%12s\t%12s\t%12s\t%12s\t%12s"
                   % (iters, name, "%d" % mem, "%.3f" % lo, "%.3f" % p50, "%.3f" % p75, "%.3f" % p95, "%.3f" % run))
        total = run
    out.append("\t%d\t%36s\t%12s\t%12s\t%12s" % (iters, "total_time", "N/A", "N/A", "%.3f" % total))
    return "\n".join(out)


def _parse(argv=None):
    ap = argparse.ArgumentParser(description="DLRM comm pattern on B200 peer-push all-to-all")
    ap.add_argument("--mini-batch-size", type=int, default=8192, help="per-rank batch")
    ap.add_argument("--num-batches", type=int, default=10)
    ap.add_argument("--warmup-batches", type=int, default=3)
    ap.add_argument("--arch-embedding-size", type=str, default="100000x16")
    ap.add_argument("--arch-sparse-feature-size", type=int, default=128)
    ap.add_argument("--arch-mlp-bot", type=str, default="4-3-2")
    ap.add_argument("--arch-mlp-top", type=str, default="4-2-1")
    ap.add_argument("--arch-interaction-op", type=str, default="dot")
    ap.add_argument("--arch-interaction-itself", action="store_true")
    ap.add_argument("--arch-project-size", type=int, default=0)
    ap.add_argument("--num-indices-per-lookup", type=int, default=20)
    ap.add_argument("--num-indices-per-lookup-fixed", type=lambda s: str(s).lower() in ("1", "true"), default=True)
    ap.add_argument("--alpha", type=float, default=0.0)
    ap.add_argument("--lr", type=float, default=0.0)
    ap.add_argument("--perf-debug", action="store_true", help="drain the device at every mark (dlrm.py:1261-1299)")
    ap.add_argument("--two-collective-dist", action="store_true",
                    help="lengths exchange and index exchange as two collectives (separate offset_xchg / idx_xchg "
                         "rows) instead of the one device-side call")
    ap.add_argument("--unfused-forward", action="store_true",
                    help="lookup kernel, then the push kernel (separate emb_lookup / fwd_a2a rows)")
    ap.add_argument("--compare-nccl", action="store_true")
    ap.add_argument("--json", action="store_true")
    return ap.parse_args(argv)


def run(argv=None):
    import time
    args = _parse(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    rows = parse_embedding_sizes(args.arch_embedding_size)
    E, b, L = args.arch_sparse_feature_size, args.mini_batch_size, args.num_indices_per_lookup
    model = DLRMParallelEmbedding(group, rows, E, b, L, dev, lr=args.lr, seed=1)
    # dense part: one random [out, in] tensor per MLP layer, all-reduced every iteration — the data-parallel
    # gradient exchange of the reference (initializeData dlrm.py:307-360, stages :1266-1281 / :1303-1317).
    # Not on the a2a path: plain NCCL all_reduce, async within a stage, the stage closed by a barrier.
    bot = [int(v) for v in args.arch_mlp_bot.split("-") if v]
    top = top_mlp_dims(len(rows), bot, [int(v) for v in args.arch_mlp_top.split("-") if v],
                       args.arch_interaction_op, args.arch_interaction_itself, args.arch_project_size)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    top_layers = [torch.rand(s, generator=g, device=dev) for s in mlp_layer_shapes(top)]
    bot_layers = [torch.rand(s, generator=g, device=dev) for s in mlp_layer_shapes(bot)]
    mem_top = sum(t.numel() * t.element_size() for t in top_layers)
    mem_bot = sum(t.numel() * t.element_size() for t in bot_layers)

    def drain():
        torch.cuda.synchronize(dev)
        dist.barrier(group)

    def all_reduce_stage(layers):
        works = [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True) for t in layers]
        for w in works:
            w.wait()
        drain()

    lat, mem, dev_ms = [], [], {"offset_idx_xchg": [], "emb_lookup_fwd_a2a": [], "bwd_a2a_emb_update": [],
                                "iter_time": []}
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    grad = None
    for it in range(args.warmup_batches + args.num_batches):
        batch = SparseBatch.synthetic(rows, b, L, args.num_indices_per_lookup_fixed, seed=rank * 1000 + it,
                                      device=dev, alpha=args.alpha)
        drain()
        m, sizes = {}, dict.fromkeys(SIZED_REGIONS, 0)

        def mark(name, sync=False):
            if sync or args.perf_debug:
                torch.cuda.synchronize(dev)
            m[name] = time.monotonic()

        e = [ev() for _ in range(4)]
        mark("iter_start")
        e[0].record()
        # lengths are the native form of SparseBatch and the inputs are generated on the device: the reference's
        # calculateLengths and host-to-device push (dlrm.py:254-277) have no counterpart — both rows read ~0
        mark("length_calc_end")
        mark("mem_push_idx_end")
        sizes["offset_xchg"] = batch.lengths.numel() * 8
        sizes["idx_xchg"] = batch.indices.numel() * 8
        mark("offset_xchg_start")
        if args.two_collective_dist:
            offsets, indices = model.sparse_data_dist(batch, device_side=False, mark=mark)
        else:
            offsets, indices = model.sparse_data_dist(batch, device_side=True)
            mark("offset_xchg_end")
            mark("idx_xchg_start")       # one call does both exchanges: the index row is empty by construction
            mark("idx_xchg_end")
        e[1].record()
        drain()
        mark("bef_emb_lookup")
        sizes["fwd_a2a"] = sizes["bwd_a2a"] = model.N * model.T_local * E * 4
        # fused (default): lookup and exchange are one kernel, so the gap row reads ~0 and fwd_a2a holds both
        out = model.forward(offsets, indices, fused=not args.unfused_forward,
                            before_exchange=lambda: mark("fwd_a2a_start"))
        mark("fwd_a2a_end")
        e[2].record()
        if grad is None:
            grad = torch.ones_like(out)
        mark("grad_push_start")
        drain()
        mark("bwd_top_ar_start")
        sizes["bwd_top_ar"] = mem_top
        all_reduce_stage(top_layers)
        mark("bwd_top_ar_end")
        e_b0 = ev()
        e_b0.record()
        mark("bwd_a2a_start")
        model.backward(grad)
        mark("bwd_a2a_end", sync=True)   # tempB.backward(C) returns with the update queued; drained here so the
        e[3].record()                    # row holds the exchange + update, not the enqueue time
        mark("bwd_bot_ar_start")
        sizes["bwd_bot_ar"] = mem_bot
        all_reduce_stage(bot_layers)
        mark("bwd_bot_ar_end")
        torch.cuda.synchronize(dev)
        if it >= args.warmup_batches:
            t = region_times_us(m)
            lat.append([t[name] for name, _, _ in REGIONS])
            mem.append([float(sizes.get(name, 0)) for name, _, _ in REGIONS])
            dev_ms["offset_idx_xchg"].append(e[0].elapsed_time(e[1]))
            dev_ms["emb_lookup_fwd_a2a"].append(e[1].elapsed_time(e[2]))
            dev_ms["bwd_a2a_emb_update"].append(e_b0.elapsed_time(e[3]))
            dev_ms["iter_time"].append(e[0].elapsed_time(e[3]))
    if model.window.error():
        raise PB200Error("a peer wait timed out during the run")
    stats = {}
    for k, v in dev_ms.items():
        t = torch.tensor(v, device=dev).median().view(1)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        stats[k + "_ms_p50_max_rank"] = float(t.item())
    # the reference's report: every rank's samples gathered, percentiles over all of them (dlrm.py:1039-1066)
    lat_t = torch.tensor(lat, dtype=torch.float64, device=dev).t().contiguous()       # [R, iters]
    mem_t = torch.tensor(mem, dtype=torch.float64, device=dev).t().contiguous()
    lat_all = [torch.empty_like(lat_t) for _ in range(world)]
    mem_all = [torch.empty_like(mem_t) for _ in range(world)]
    dist.all_gather(lat_all, lat_t, group=group)
    dist.all_gather(mem_all, mem_t, group=group)
    if rank == 0:
        rows_all, rows_mean = percentile_rows(torch.stack(lat_all).cpu(), torch.stack(mem_all).cpu())
        stats["regions_us_p50"] = {r[0]: round(r[3], 3) for r in rows_all}
        if args.json:
            print(json.dumps(stats))
        else:
            print(format_report(args.num_batches, rows_all))
            print("\n\n " + "-" * 125 + "\n\n")
            print(format_report(args.num_batches, rows_mean, header=False))
            print()
            for k, v in stats.items():
                if k.endswith("_ms_p50_max_rank"):
                    print(f"device time {k:40s} {v:10.3f} ms")
    dist.barrier(group)
    return stats


if __name__ == "__main__":
    run()
    if dist.is_initialized():
        dist.destroy_process_group()
