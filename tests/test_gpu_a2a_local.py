"""The peer-push all-to-all protocol exercised on ONE GPU: W virtual ranks in one process, each
with its own window, pad and stream (PeerWindow.local_group).  The kernels are the same ones that
run across NVLink; only the pointers differ.  Golden vectors come from the reference's c10d/gloo
run and its own All2Allv_Req/All2Allv_Wait Functions (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _group(W, window_bytes, dev):
    from param_b200.comms.pt.peer_window import PeerWindow
    return PeerWindow.local_group(W, window_bytes, dev, max_ctas=4, spin_timeout_s=5.0)


def _run_all(grp, fn):
    """launch fn(rank, window, stream) for every virtual rank, then synchronise and check errors"""
    outs = []
    for r, (w, st) in enumerate(zip(grp.windows, grp.streams)):
        st.wait_stream(torch.cuda.current_stream())   # inputs were produced on the current stream
        with torch.cuda.stream(st):
            outs.append(fn(r, w, st))
    torch.cuda.synchronize()
    for w in grp.windows:
        assert w.error() == 0, "a2a kernel timed out waiting for a peer"
    return outs


def test_all_to_all_single_matches_c10d_golden(cuda_device, golden_dir):
    d = np.load(golden_dir / "a2a_gloo_ref.npz")
    W = int(d["world"])
    splits = d["splits"]
    grp = _group(W, 1 << 16, cuda_device)
    for tag, dt in (("i64", torch.int64), ("f32", torch.float32)):
        ins = [torch.from_numpy(d[f"r{r}_raw_{tag}_in"]).to(cuda_device) for r in range(W)]
        outs = [torch.empty(int(splits[:, r].sum()), dtype=dt, device=cuda_device) for r in range(W)]
        _run_all(grp, lambda r, w, st: w.all_to_all_single(
            outs[r], ins[r], [int(splits[s][r]) for s in range(W)], [int(x) for x in splits[r]], stream=st))
        for r in range(W):
            assert np.array_equal(outs[r].cpu().numpy(), d[f"r{r}_raw_{tag}_out"]), (tag, r)


def test_pooled_exchange_matches_reference_functions(cuda_device, golden_dir):
    d = np.load(golden_dir / "a2a_gloo_ref.npz")
    W = int(d["world"])
    N, E, Tg = (int(x) for x in d["pool_dims"])
    ts = [int(x) for x in d["r0_pool_tables_split"]]
    bs = [int(x) for x in d["r0_pool_batch_split"]]
    grp = _group(W, 1 << 16, cuda_device)
    for layout in ("TBD", "BTD"):
        lys = []
        for r in range(W):
            ly = torch.from_numpy(d[f"r{r}_pool_ly"]).to(cuda_device)          # [T_r, N, E]
            if layout == "BTD":
                ly = ly.permute(1, 0, 2).contiguous().view(N, -1)              # [N, T_r*E]
            lys.append(ly)
        outs = _run_all(grp, lambda r, w, st: w.pooled_forward(lys[r], bs, ts, E, layout=layout, stream=st))
        for r in range(W):
            assert np.array_equal(outs[r].cpu().numpy(), d[f"r{r}_pool_out"]), (layout, r)
    grads = [torch.from_numpy(d[f"r{r}_pool_gradout"]).to(cuda_device) for r in range(W)]
    gins = _run_all(grp, lambda r, w, st: w.pooled_backward(grads[r], bs, ts, E, stream=st))
    for r in range(W):
        got = gins[r].view(N, ts[r], E).permute(1, 0, 2).cpu().numpy()          # -> [T_r, N, E]
        assert np.array_equal(got, d[f"r{r}_pool_gradin"]), r


@pytest.mark.parametrize("W", [2, 4, 8])
@pytest.mark.parametrize("dtype", [torch.float32, torch.int64, torch.uint8])
def test_all_to_all_single_vs_oracle_repeated(cuda_device, oracle, W, dtype):
    """position-coded payloads (a constant payload cannot detect a wrong permutation),
    uneven splits incl. zero-length blocks, several epochs on the same communicator."""
    rng = np.random.default_rng(W * 10 + dtype.itemsize)
    grp = _group(W, 1 << 20, cuda_device)
    np_dt = {torch.float32: np.float32, torch.int64: np.int64, torch.uint8: np.uint8}[dtype]
    for it in range(4):
        splits = rng.integers(0, 300, size=(W, W))
        splits[rng.integers(0, W), rng.integers(0, W)] = 0
        if it == 3:                                       # equal-split form (no split lists)
            splits[:] = 64
        ins_h = [((np.arange(splits[r].sum()) + 1000 * r + 7 * it) % 251).astype(np_dt) for r in range(W)]
        want = oracle.all_to_all_single(ins_h, splits)
        ins = [torch.from_numpy(x).to(cuda_device) for x in ins_h]
        outs = [torch.empty(int(splits[:, r].sum()), dtype=dtype, device=cuda_device) for r in range(W)]
        if it == 3:
            _run_all(grp, lambda r, w, st: w.all_to_all_single(outs[r], ins[r], stream=st))
        else:
            _run_all(grp, lambda r, w, st: w.all_to_all_single(
                outs[r], ins[r], [int(splits[s][r]) for s in range(W)], [int(x) for x in splits[r]], stream=st))
        for r in range(W):
            assert np.array_equal(outs[r].cpu().numpy(), want[r]), (W, dtype, it, r)


def test_zero_copy_output_in_window(cuda_device, oracle):
    W = 4
    grp = _group(W, 1 << 20, cuda_device)
    n = 4096
    outs = [w.alloc(n, torch.float32)[0] for w in grp.windows]
    ins_h = [(np.arange(n) + 10000 * r).astype(np.float32) for r in range(W)]
    ins = [torch.from_numpy(x).to(cuda_device) for x in ins_h]
    _run_all(grp, lambda r, w, st: w.all_to_all_single(outs[r], ins[r], stream=st))
    want = oracle.all_to_all_single(ins_h, np.full((W, W), n // W))
    for r in range(W):
        assert grp.windows[r].offset_of(outs[r]) is not None
        assert np.array_equal(outs[r].cpu().numpy(), want[r])


@pytest.mark.parametrize("W,T,E", [(2, 4, 32), (4, 6, 32), (8, 8, 32), (3, 5, 32), (2, 128, 128), (2, 32, 128), (4, 96, 128)])
def test_pooled_exchange_vs_oracle(cuda_device, oracle, W, T, E):
    from param_b200.comms.pt.dlrm import split_lengths
    N = 50
    ts = split_lengths(T, W)
    bs = split_lengths(N, W)
    rng = np.random.default_rng(W + T)
    pooled_h = [rng.standard_normal((ts[r], N, E)).astype(np.float32) for r in range(W)]
    want = oracle.pooled_a2a_fwd(pooled_h, bs, ts, E)
    grp = _group(W, 16 << 20, cuda_device)
    btd = [torch.from_numpy(p).to(cuda_device).permute(1, 0, 2).contiguous().view(N, -1) for p in pooled_h]
    outs = _run_all(grp, lambda r, w, st: w.pooled_forward(btd[r], bs, ts, E, layout="BTD", stream=st))
    for r in range(W):
        assert np.array_equal(outs[r].cpu().numpy(), want[r])
    grads_h = [rng.standard_normal((bs[r], T * E)).astype(np.float32) for r in range(W)]
    want_b = oracle.pooled_a2a_bwd(grads_h, bs, ts, E)
    grads = [torch.from_numpy(g).to(cuda_device) for g in grads_h]
    gins = _run_all(grp, lambda r, w, st: w.pooled_backward(grads[r], bs, ts, E, out_window_off=8 << 20, stream=st))
    for r in range(W):
        got = gins[r].view(N, ts[r], E).permute(1, 0, 2).cpu().numpy()
        assert np.array_equal(got, want_b[r])


def test_missing_peer_times_out_instead_of_hanging(cuda_device):
    from param_b200.comms.pt.peer_window import PeerWindow
    grp = PeerWindow.local_group(2, 1 << 12, cuda_device, max_ctas=1, spin_timeout_s=0.2)
    x = torch.arange(64, dtype=torch.float32, device=cuda_device)
    grp.windows[0].all_to_all_single(None, x)          # rank 1 never calls
    torch.cuda.synchronize()
    assert grp.windows[0].error() == 1


@pytest.mark.parametrize("W,T,dim", [(2, 4, 128), (4, 6, 64), (8, 9, 128), (3, 5, 56)])
def test_fused_lookup_exchange_vs_oracle(cuda_device, oracle, W, T, dim):
    """pb200_tbe_fwd_a2a: lookup + exchange + permute in ONE kernel per rank, bit-exact against
    oracle lookup -> oracle pooled exchange; then again on the same communicator (epochs), mixed
    with a plain all_to_all_single."""
    from param_b200 import ops
    from param_b200.comms.pt.dlrm import split_lengths
    rng = np.random.default_rng(W * 31 + T)
    ts, b = split_lengths(T, W), 24
    bs = [b] * W
    N = b * W
    grp = _group(W, 4 << 20, cuda_device)
    for rep in range(2):
        arenas, reqs, pooled_ref = [], [], []
        for r in range(W):
            rows = rng.integers(30, 200, size=ts[r])
            tro = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
            w = rng.standard_normal((int(tro[-1]), dim)).astype(np.float32)
            lens = rng.integers(0, 12, size=ts[r] * N)
            off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            idx = np.concatenate([rng.integers(0, rows[t], size=int(lens[t * N:(t + 1) * N].sum()))
                                  for t in range(ts[r])]).astype(np.int64)
            arenas.append(ops.TableArena(torch.from_numpy(w).to(cuda_device), torch.from_numpy(tro).to(cuda_device),
                                         list(rows), dim))
            reqs.append((torch.from_numpy(idx).to(cuda_device), torch.from_numpy(off).to(cuda_device)))
            pooled_ref.append(oracle.tbe_fwd(w, tro, dim, idx, off, N, layout="TBD"))
        want = oracle.pooled_a2a_fwd(pooled_ref, bs, ts, dim)
        outs = _run_all(grp, lambda r, w_, st: w_.lookup_forward_fused(arenas[r], reqs[r][0], reqs[r][1], bs, ts, stream=st))
        for r in range(W):
            assert np.array_equal(outs[r].cpu().numpy(), want[r]), (rep, r)
        xs = [torch.full((W * 8,), float(r), device=cuda_device) for r in range(W)]
        ys = _run_all(grp, lambda r, w_, st: w_.all_to_all_single(None, xs[r], out_window_off=2 << 20, stream=st).clone())
        for r in range(W):
            assert ys[r].view(W, 8)[:, 0].tolist() == [float(s) for s in range(W)]


@pytest.mark.parametrize("W,T,b,max_len", [(2, 4, 16, 6), (3, 7, 9, 5), (4, 4, 32, 20), (8, 19, 8, 3)])
def test_sparse_data_dist_device_side_vs_oracle(cuda_device, oracle, W, T, b, max_len):
    """pb200_sparse_data_dist (lengths push -> device-resident index counts -> index push into fixed
    slots -> regroup, no host round trip) against the oracle's c10d all_to_all_single of the lengths
    and indices followed by splitPerTable — int64, bit-exact; run twice on the same communicator with
    different data (epochs, stale slot tails from the first round)."""
    from param_b200.comms.pt.dlrm import split_lengths
    rng = np.random.default_rng(W * 17 + T)
    ts = split_lengths(T, W)
    T_max = max(ts)
    slot = T_max * b * max_len
    n_len_max = W * T_max * b
    win_bytes = (n_len_max + W * slot) * 8 + 2048
    grp = _group(W, win_bytes, cuda_device)
    off_len = 0
    off_idx = (n_len_max * 8 + 511) // 512 * 512
    bases = np.concatenate([[0], np.cumsum(ts)])
    for rep in range(2):
        lens_h = [rng.integers(0, max_len + 1, size=T * b).astype(np.int64) for _ in range(W)]
        if rep == 1:
            lens_h[0][:] = 0                                  # a rank that sends nothing at all
        idx_h = [((np.arange(int(l.sum())) * 7 + 100000 * r + rep) % (1 << 40)).astype(np.int64)
                 for r, l in enumerate(lens_h)]
        # oracle: the two all_to_all_single of SparseDataDist, then splitPerTable per rank
        len_splits = np.array([[ts[d] * b for d in range(W)] for _ in range(W)])
        lens_recv = oracle.all_to_all_single(lens_h, len_splits)
        idx_splits = np.array([[int(lens_h[s][bases[d] * b:bases[d + 1] * b].sum()) for d in range(W)] for s in range(W)])
        idx_recv = oracle.all_to_all_single(idx_h, idx_splits)
        lens_d = [torch.from_numpy(x).to(cuda_device) for x in lens_h]
        idx_d = [torch.from_numpy(x).to(cuda_device) for x in idx_h]
        outs = _run_all(grp, lambda r, w, st: w.sparse_data_dist(lens_d[r], idx_d[r], ts, b, off_len, off_idx, slot,
                                                                 stream=st))
        for r in range(W):
            want_len, want_off, want_idx = oracle.split_per_table(lens_recv[r], idx_recv[r], W, ts[r], b)
            got_len, got_off, got_idx = (x.cpu().numpy() for x in outs[r])
            assert np.array_equal(got_len, want_len), (rep, r)
            assert np.array_equal(got_off, want_off), (rep, r)
            assert np.array_equal(got_idx[:want_off[-1]], want_idx), (rep, r)


def test_sparse_data_dist_slot_overflow_is_reported(cuda_device):
    """a block larger than its slot is truncated (never written past the slot) and flagged"""
    W, T, b = 2, 2, 4
    grp = _group(W, 1 << 16, cuda_device)
    lens = [torch.full((T * b,), 5, dtype=torch.int64, device=cuda_device) for _ in range(W)]
    idx = [torch.arange(T * b * 5, dtype=torch.int64, device=cuda_device) for _ in range(W)]
    slot = 8                                                  # needs 1 table * 4 bags * 5 = 20
    for r, (w, st) in enumerate(zip(grp.windows, grp.streams)):
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            w.sparse_data_dist(lens[r], idx[r], [1, 1], b, 0, 4096, slot, stream=st)
    torch.cuda.synchronize()
    assert all(w.error() == 2 for w in grp.windows)


@pytest.mark.parametrize("W", [2, 4])
def test_staged_output_never_lands_on_live_window_tensors(cuda_device, oracle, W):
    """an output that lives outside the window used to be staged at window offset 0,
    on top of the comm buffers the bump allocator hands out from there.  Now it is staged at the top of
    the window: the live in-window tensors (here: the INPUT of the very same collective) are intact
    after several iterations, and a result that cannot fit beside them raises instead of overwriting."""
    from param_b200._cabi import PB200Error
    grp = _group(W, 1 << 20, cuda_device)
    n = 8192
    rng = np.random.default_rng(W)
    ins_h = [rng.integers(0, 1 << 20, size=n).astype(np.int64) for _ in range(W)]
    ins = []
    for r, w in enumerate(grp.windows):
        t, off = w.alloc(n, torch.int64)        # what alloc_random does for ipTensor: offset 0 of the window
        assert off == 0
        t.copy_(torch.from_numpy(ins_h[r]))
        ins.append(t)
    splits = np.zeros((W, W), np.int64)
    for r in range(W):          # every rank sends all n elements, cut at random places
        cuts = np.sort(rng.choice(np.arange(1, n), size=W - 1, replace=False))
        splits[r] = np.diff(np.concatenate([[0], cuts, [n]]))
    want = oracle.all_to_all_single(ins_h, splits)
    outs = [torch.empty(int(splits[:, r].sum()), dtype=torch.int64, device=cuda_device) for r in range(W)]
    for _ in range(3):      # iteration 2 used to re-send what iteration 1 had received
        _run_all(grp, lambda r, w, st: w.all_to_all_single(
            outs[r], ins[r], [int(splits[s][r]) for s in range(W)], [int(x) for x in splits[r]], stream=st))
        for r in range(W):
            assert np.array_equal(outs[r].cpu().numpy(), want[r])
            assert np.array_equal(ins[r].cpu().numpy(), ins_h[r]), "the live input buffer was overwritten"
    # no room left beside the live buffers: refuse
    for w in grp.windows:
        w.alloc((1 << 20) // 8 - n - 64, torch.int64)
    with pytest.raises(PB200Error):
        grp.windows[0].all_to_all_single(outs[0], ins[0], [int(splits[s][0]) for s in range(W)],
                                         [int(x) for x in splits[0]])


@pytest.mark.parametrize("W", [2, 3, 8])
@pytest.mark.parametrize("in_window", [False, True])
def test_list_form_all_to_all(cuda_device, W, in_window):
    """dist.all_to_all(output_tensor_list, input_tensor_list) (pytorch_dist_backend.py:207-260) as ONE push
    kernel with per-destination sources and per-source landing places: ragged block sizes, outputs
    inside the window (written in place by the peers) or outside (staged + copied out)."""
    grp = _group(W, 1 << 20, cuda_device)
    rng = np.random.default_rng(W + 50)
    sizes = rng.integers(0, 500, size=(W, W))          # sizes[s][d]: elements rank s sends to rank d
    sizes[0, W - 1] = 0
    ins = [[(torch.arange(int(sizes[s][d]), dtype=torch.float32, device=cuda_device) + 1000 * s + 10 * d)
            for d in range(W)] for s in range(W)]
    outs = []
    for r, w in enumerate(grp.windows):
        if in_window:
            w.alloc(77, torch.float32)                  # something live at offset 0
            outs.append([w.alloc(int(sizes[s][r]), torch.float32)[0] for s in range(W)])
        else:
            outs.append([torch.empty(int(sizes[s][r]), dtype=torch.float32, device=cuda_device) for s in range(W)])
    for _ in range(2):
        _run_all(grp, lambda r, w, st: w.all_to_all(outs[r], ins[r], stream=st))
        for r in range(W):
            for s in range(W):
                assert torch.equal(outs[r][s], ins[s][r]), (r, s)


def test_all_to_all_single_counts_splits_along_dim0(cuda_device, oracle):
    """c10d counts split sizes in rows of dim 0: a [rows, 6] tensor with splits in rows"""
    W = 2
    grp = _group(W, 1 << 18, cuda_device)
    rows = np.array([[3, 5], [2, 7]])
    ins_h = [np.arange(rows[r].sum() * 6, dtype=np.float32).reshape(-1, 6) + 100 * r for r in range(W)]
    want = oracle.all_to_all_single([x.reshape(-1) for x in ins_h], rows * 6)
    ins = [torch.from_numpy(x).to(cuda_device) for x in ins_h]
    outs = [torch.empty((int(rows[:, r].sum()), 6), device=cuda_device) for r in range(W)]
    _run_all(grp, lambda r, w, st: w.all_to_all_single(
        outs[r], ins[r], [int(rows[s][r]) for s in range(W)], [int(x) for x in rows[r]], stream=st))
    for r in range(W):
        assert np.array_equal(outs[r].cpu().numpy().reshape(-1), want[r])
    from param_b200._cabi import PB200Error
    with pytest.raises(PB200Error):     # output too small for what the splits deliver
        grp.windows[0].all_to_all_single(torch.empty((2, 6), device=cuda_device), ins[0], [3, 2], [3, 5])


@pytest.mark.parametrize("W,T,parts", [(2, 7, 2), (4, 9, 3), (3, 4, 4)])
def test_pooled_backward_in_pieces_equals_whole(cuda_device, oracle, W, T, parts):
    """pb200_a2a_pooled_bwd_part: the transpose exchange cut into table groups (so that the reduce of group g can
    run under the exchange of group g + 1) lands exactly what the single exchange lands — also when a rank owns
    fewer tables than there are pieces (empty pieces)."""
    from param_b200.comms.pt.dlrm import split_lengths
    N, E = 40, 32
    ts, bs = split_lengths(T, W), split_lengths(N, W)
    rng = np.random.default_rng(W * 100 + T)
    grads_h = [rng.standard_normal((bs[r], T * E)).astype(np.float32) for r in range(W)]
    want = oracle.pooled_a2a_bwd(grads_h, bs, ts, E)
    grads = [torch.from_numpy(g).to(cuda_device) for g in grads_h]
    grp = _group(W, 8 << 20, cuda_device)
    for w in grp.windows:
        w.view(0, 1 << 20, torch.float32).fill_(float("nan"))
    for g in range(parts):
        gins = _run_all(grp, lambda r, w, st: w.pooled_backward(grads[r], bs, ts, E, stream=st, part=g, parts=parts))
    for r in range(W):
        got = gins[r].view(N, ts[r], E).permute(1, 0, 2).cpu().numpy()
        assert np.array_equal(got, want[r]), r
