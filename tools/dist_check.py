#!/usr/bin/env python3
"""Multi-GPU parity check of the peer-window path (run under torchrun on N GPUs of one box):

    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py

Checks, each against c10d/NCCL on the same tensors (bit-exact) and, at small sizes, against the CPU
oracle gathered on rank 0:
  1. peer mapping (symmetric memory, CUDA-IPC fallback) comes up;
  2. all_to_all_single: equal + uneven splits, int64 / fp32 / uint8, zero-copy and staged outputs,
     repeated epochs, inside a CUDA graph;
  3. the fused pooled exchange forward / backward vs the reference's cat + all_to_all_single + cat;
  4. the whole DLRM step (sparse_data_dist -> lookup -> exchange -> backward) vs an oracle run of
     the same global problem.
Prints one line per check and exits non-zero on any mismatch.
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from param_b200 import ops  # noqa: E402
from param_b200.comms.pt.dlrm import (DLRMParallelEmbedding, SparseBatch, owner_slice,  # noqa: E402
                                      split_lengths)
from param_b200.comms.pt.peer_window import PeerWindow  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
fails = 0


def report(name, ok):
    global fails
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"[dist_check] {name}: {'PASS' if int(t) else 'FAIL'}", flush=True)
    if not int(t):
        fails += 1


win = PeerWindow.create(dist.group.WORLD, 256 << 20, dev)
win.configure(spin_timeout_s=5.0)
report(f"peer mapping via {win.mapping}", True)

# ---- 2. all_to_all_single ---------------------------------------------------------------------
rng = np.random.default_rng(7)          # same stream on every rank -> same split matrix
for it, dtype in enumerate((torch.float32, torch.int64, torch.uint8)):
    splits = rng.integers(0, 5000, size=(world, world))
    splits[0, world - 1] = 0
    if dtype == torch.float32:
        splits *= 4                      # 16 B-aligned blocks -> vector path; others exercise 4 B / 1 B
    in_s = [int(x) for x in splits[rank]]
    out_s = [int(splits[s][rank]) for s in range(world)]
    x = ((torch.arange(sum(in_s), device=dev) + 7919 * rank) % 251).to(dtype)
    ref = torch.empty(sum(out_s), dtype=dtype, device=dev)
    dist.all_to_all_single(ref, x, out_s, in_s)
    got = torch.empty_like(ref)
    for _ in range(3):                   # repeated epochs
        got.zero_()
        win.all_to_all_single(got, x, out_s, in_s)
    report(f"all_to_all_single uneven {dtype}", torch.equal(got, ref))
n = 1 << 20
x = torch.arange(n * world, device=dev, dtype=torch.float32) + rank * 0.5
ref = torch.empty_like(x)
dist.all_to_all_single(ref, x)
win.reset_alloc()
zc, _ = win.alloc(n * world, torch.float32)
win.all_to_all_single(zc, x)
report("all_to_all_single equal split, zero-copy output in window", torch.equal(zc, ref))
# CUDA graph capture + replay (comms.py run_coll_cuda_graph, comms.py:375-450)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    zc.zero_()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(4):
            win.all_to_all_single(zc, x)
    g.replay()
    g.replay()
st.synchronize()
report("all_to_all_single inside a replayed CUDA graph", torch.equal(zc, ref))

# staged output next to LIVE window tensors (the staging area used to be window offset 0, where the
# bump allocator puts ipTensor): the input lives in the window, the output does not, three iterations
win.reset_alloc()
xin_w, _ = win.alloc(n * world, torch.float32)
xin_w.copy_(x)
got = torch.empty_like(ref)
ok = True
for _ in range(3):
    got.zero_()
    win.all_to_all_single(got, xin_w)
    ok &= torch.equal(got, ref) and torch.equal(xin_w, x)
report("all_to_all_single staged output beside a live in-window input (3 iterations)", ok)
# list form: dist.all_to_all(list, list) as ONE push kernel, ragged blocks, outputs outside / inside the window
sizes = rng.integers(0, 3000, size=(world, world))
ins_l = [torch.arange(int(sizes[rank][d]), device=dev, dtype=torch.float32) + 1000 * rank + d for d in range(world)]
ref_l = [torch.empty(int(sizes[s][rank]), device=dev) for s in range(world)]
dist.all_to_all(ref_l, ins_l)
got_l = [torch.empty_like(t) for t in ref_l]
win.all_to_all(got_l, ins_l)
report("all_to_all (list form) staged outputs == c10d", all(torch.equal(a, b_) for a, b_ in zip(got_l, ref_l)))
got_w = [win.alloc(int(sizes[s][rank]), torch.float32)[0] for s in range(world)]
for _ in range(2):
    win.all_to_all(got_w, ins_l)
report("all_to_all (list form) outputs in the window, written in place == c10d",
       all(torch.equal(a, b_) for a, b_ in zip(got_w, ref_l)) and torch.equal(xin_w, x))

# ---- 3. fused pooled exchange vs the reference's cat + a2a + cat --------------------------------
T_g, b, E = 3 * world + 1, 64, 128
ts, bs = split_lengths(T_g, world), [b] * world
T_l, N = ts[rank], b * world
pooled = torch.randn(N, T_l * E, device=dev)
win.reset_alloc()
out = win.pooled_forward(pooled, bs, ts, E, layout="BTD", out_window_off=0)
ly = pooled.view(N, T_l, E).permute(1, 0, 2).contiguous()                       # [T_l, N, E] (dlrm.py:387)
inp = torch.cat(list(ly), dim=1).view(-1)                                       # dlrm.py:97
in_s, out_s = [m * T_l * E for m in bs], [b * t * E for t in ts]                # dlrm.py:95-96
o = inp.new_empty(sum(out_s))
dist.all_to_all_single(o, inp, out_s, in_s)
ref = torch.cat([p.view(b, -1) for p in o.split(out_s)], dim=1)                 # dlrm.py:173-175, 1253
report("pooled forward (fused permute) == reference cat/a2a/cat", torch.equal(out, ref))
out_tbd = win.pooled_forward(ly, bs, ts, E, layout="TBD", out_window_off=64 << 20)
report("pooled forward from [T,N,E] (multi-round) == reference", torch.equal(out_tbd, ref))
grad = torch.randn(b, T_g * E, device=dev)
gin = win.pooled_backward(grad, bs, ts, E, out_window_off=128 << 20)             # [N, T_l*E]
go = torch.cat([g_.contiguous().view(-1) for g_ in grad.split([t * E for t in ts], dim=1)])   # dlrm.py:188-189
gi = go.new_empty(N * T_l * E)
dist.all_to_all_single(gi, go, in_s, out_s)
report("pooled backward (fused transpose) == reference", torch.equal(gin, gi.view(N, T_l * E)))

# ---- 4. whole DLRM step vs the oracle on the gathered global problem -----------------------------
from oracle import oracle  # checker only  # noqa: E402

rows, L, lr = 500, 6, 0.25
T_g, b, E = 2 * world + 1, 16, 64
model = DLRMParallelEmbedding(dist.group.WORLD, [rows] * T_g, E, b, L, dev, lr=lr, seed=5, bwd_algo="sorted")
batch = SparseBatch.synthetic([rows] * T_g, b, L, False, seed=100 + rank, device=dev)
w0 = model.arena.weights.clone()
offsets, indices = model.sparse_data_dist(batch)
out2 = model.forward(offsets, indices, fused=False).clone()   # lookup kernel + push kernel
out = model.forward(offsets, indices)                         # ONE fused kernel (default)
gsum = out * 0.5 + 1.0                                       # some dOut that depends on the data
model.backward(gsum)
torch.cuda.synchronize()
# gather the global problem on every rank (small) and replay it on the CPU oracle
gl = [torch.empty_like(batch.lengths) for _ in range(world)]
dist.all_gather(gl, batch.lengths)
cnt = torch.tensor([batch.indices.numel()], device=dev)
cnts = [torch.zeros_like(cnt) for _ in range(world)]
dist.all_gather(cnts, cnt)
mx = int(max(int(c) for c in cnts))
pad = torch.zeros(mx, dtype=torch.int64, device=dev)
pad[:batch.indices.numel()] = batch.indices
gi_ = [torch.empty_like(pad) for _ in range(world)]
dist.all_gather(gi_, pad)
ts = split_lengths(T_g, world)
sl = owner_slice(rank, ts)
N = b * world
lens_np = [g_.cpu().numpy().reshape(T_g, b) for g_ in gl]
idx_np = [gi_[r][:int(cnts[r])].cpu().numpy() for r in range(world)]
# per owned table: bags of all ranks in rank order
tbl_idx, tbl_len = [], []
for t in range(sl.start, sl.stop):
    ii, ll = [], []
    for r in range(world):
        starts = np.concatenate([[0], np.cumsum(lens_np[r].reshape(-1))])
        lo, hi = starts[t * b], starts[(t + 1) * b]
        ii.append(idx_np[r][lo:hi])
        ll.append(lens_np[r][t])
    tbl_idx.append(np.concatenate(ii))
    tbl_len.append(np.concatenate(ll))
want_idx = np.concatenate(tbl_idx)
want_off = np.concatenate([[0], np.cumsum(np.concatenate(tbl_len))]).astype(np.int64)
n_valid = int(want_off[-1])
report("sparse_data_dist (ONE call, device-resident counts, slot layout) == global regroup",
       np.array_equal(indices.cpu().numpy()[:n_valid], want_idx) and np.array_equal(offsets.cpu().numpy(), want_off))
off_h, idx_h = model.sparse_data_dist(batch, device_side=False)
report("sparse_data_dist (2 peer a2a + one count D2H + device regroup) == global regroup",
       np.array_equal(idx_h.cpu().numpy(), want_idx) and np.array_equal(off_h.cpu().numpy(), want_off))
T_l = ts[rank]
tro = np.arange(T_l + 1, dtype=np.int64) * rows
pooled_ref = oracle.tbe_fwd(w0.cpu().numpy(), tro, E, want_idx, want_off, N, layout="TBD")     # [T_l, N, E]
allp = [torch.empty(ts[r], N, E, device=dev) for r in range(world)]
# all_gather with uneven shapes: pad to max tables
Tm = max(ts)
mine = torch.zeros(Tm, N, E, device=dev)
mine[:T_l] = torch.from_numpy(pooled_ref).to(dev)
gath = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(gath, mine)
pooled_all = [gath[r][:ts[r]].cpu().numpy() for r in range(world)]
want_out = oracle.pooled_a2a_fwd(pooled_all, [b] * world, ts, E)[rank]
report("DLRM forward (ONE fused lookup+exchange kernel) == oracle (bit-exact)", np.array_equal(out.cpu().numpy(), want_out))
report("DLRM forward (lookup kernel + push kernel) == oracle (bit-exact)", np.array_equal(out2.cpu().numpy(), want_out))
gs = [torch.empty_like(gsum) for _ in range(world)]
dist.all_gather(gs, gsum)
gin_ref = oracle.pooled_a2a_bwd([g_.cpu().numpy() for g_ in gs], [b] * world, ts, E)[rank]     # [T_l, N, E]
want_w = w0.cpu().numpy().astype(np.float64) + oracle.tbe_bwd(T_l * rows, tro, E, want_idx, want_off, N, gin_ref,
                                                               layout="TBD", scale=-lr, dtype=np.float64)
err = np.abs(model.arena.weights.cpu().numpy() - want_w).max() / np.abs(want_w).max()
report(f"DLRM backward (fused transpose + scatter-add) rel err {err:.2e} <= 1e-5", err <= 1e-5)
report("no peer-wait timeouts", win.error() == 0 and model.window.error() == 0)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if fails else 0)
