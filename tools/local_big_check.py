#!/usr/bin/env python3
"""Single-GPU stress of the peer kernels at cfg4-like sizes with W virtual ranks (local group):
sparse input exchange (lengths + indices all_to_all_single + regroup), fused lookup+exchange,
push-kernel exchange fwd/bwd.  Compared against the oracle-free invariants / torch reference ops.
Meant to be run plainly and under `compute-sanitizer --tool memcheck`."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from param_b200 import ops  # noqa: E402
from param_b200.comms.pt.dlrm import indices_exchange_counts, lengths_exchange_splits, split_lengths  # noqa: E402
from param_b200.comms.pt.peer_window import PeerWindow  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 2
T_l = int(sys.argv[2]) if len(sys.argv) > 2 else 64
b = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
L, E, rows = 20, 128, 50_000
dev = torch.device("cuda:0")
T_g, N = T_l * W, b * W
ts, bs = [T_l] * W, [b] * W
cap_len, cap_idx = W * T_l * b, W * T_l * b * L
need = (cap_len + cap_idx) * 8 + b * T_g * E * 4 + N * T_l * E * 4 + 4096
grp = PeerWindow.local_group(W, need, dev, max_ctas=max(1, 128 // W), spin_timeout_s=20.0)
off_len, off_idx = 0, cap_len * 8
off_pool, off_grad = off_idx + cap_idx * 8, off_idx + cap_idx * 8 + b * T_g * E * 4


def run_all(fn):
    outs = []
    for r, (w, st) in enumerate(zip(grp.windows, grp.streams)):
        st.wait_stream(torch.cuda.current_stream())   # inputs were produced on the current stream
        with torch.cuda.stream(st):
            outs.append(fn(r, w, st))
    torch.cuda.synchronize()
    assert all(w.error() == 0 for w in grp.windows), "timeout"
    return outs


g = torch.Generator(device=dev).manual_seed(1)
lengths = [torch.full((T_g * b,), L, dtype=torch.int64, device=dev) for _ in range(W)]
indices = [torch.randint(0, rows, (T_g * b * L,), generator=g, device=dev) for _ in range(W)]
# 1. lengths exchange
lo = run_all(lambda r, w, st: w.all_to_all_single(None, lengths[r], *lengths_exchange_splits(ts, r, b)[::-1],
                                                  out_window_off=off_len, stream=st))
print("lengths a2a ok", [int(x.sum()) for x in lo])
cnt = [torch.stack(indices_exchange_counts(lengths[r], lo[r], ts, b)).cpu() for r in range(W)]
io = run_all(lambda r, w, st: w.all_to_all_single(None, indices[r], cnt[r][1].tolist(), cnt[r][0].tolist(),
                                                  out_window_off=off_idx, stream=st))
# reference: what rank r must receive = cat over sources s of s's block for r
for r in range(W):
    want = torch.cat([indices[s][r * T_l * b * L:(r + 1) * T_l * b * L] for s in range(W)])
    assert torch.equal(io[r], want), f"indices a2a rank {r}"
print("indices a2a ok")
reqs = [ops.regroup_sparse(lo[r], io[r], W, T_l, b) for r in range(W)]
torch.cuda.synchronize()
for r in range(W):
    _, off, idx = reqs[r]
    assert int(off[-1]) == idx.numel() and bool((off[1:] >= off[:-1]).all())
    want = torch.cat([io[r].view(W, T_l, b * L)[:, t] for t in range(T_l)]).view(-1)
    assert torch.equal(idx, want), f"regroup rank {r}"
print("regroup ok")
arenas = []
for r in range(W):
    a = ops.TableArena.allocate([rows] * T_l, E, dev)
    ops.fill_uniform_(a.weights, -0.01, 0.01, seed=r)
    arenas.append(a)
pooled = [ops.tbe_forward(arenas[r], reqs[r][2], reqs[r][1], N, layout="BTD") for r in range(W)]
f1 = [t.clone() for t in run_all(lambda r, w, st: w.pooled_forward(pooled[r], bs, ts, E, out_window_off=off_pool, stream=st))]
f2 = run_all(lambda r, w, st: w.lookup_forward_fused(arenas[r], reqs[r][2], reqs[r][1], bs, ts,
                                                     out_window_off=off_pool, stream=st))
def diag(name, got, want):
    if torch.equal(got, want):
        return True
    bad = (got != want)
    nz = bad.nonzero()
    print(f"MISMATCH {name}: {int(bad.sum())} of {bad.numel()} elements; first at {nz[0].tolist()}, last at {nz[-1].tolist()}; "
          f"rows affected {int(bad.any(dim=1).sum())} of {bad.shape[0]}; cols affected {int(bad.any(dim=0).sum())} of {bad.shape[1]}; "
          f"got there {got[tuple(nz[0].tolist())].item()} want {want[tuple(nz[0].tolist())].item()}", flush=True)
    cols = bad.any(dim=0).nonzero().view(-1)
    rws = bad.any(dim=1).nonzero().view(-1)
    print("  bad col range", int(cols.min()), int(cols.max()), " bad row range", int(rws.min()), int(rws.max()), flush=True)
    return False


okall = True
for r in range(W):
    want = torch.cat([pooled[s][r * b:(r + 1) * b] for s in range(W)], dim=1)
    okall &= diag(f"push fwd rank {r}", f1[r], want)
    okall &= diag(f"fused fwd rank {r}", f2[r], want)
assert okall
print("pooled forward (push) and fused lookup+exchange ok")
grads = [torch.randn(b, T_g * E, device=dev) for _ in range(W)]
gb = run_all(lambda r, w, st: w.pooled_backward(grads[r], bs, ts, E, out_window_off=off_grad, stream=st))
for r in range(W):
    want = torch.cat([grads[s][:, r * T_l * E:(r + 1) * T_l * E] for s in range(W)], dim=0)
    assert torch.equal(gb[r], want), f"push bwd rank {r}"
print("pooled backward ok")
print("ALL OK")
