"""CPU model of the run decomposition used by the EXACT backward (param_b200/csrc/emb_bwd_exact.cu):
E1 sums runs of equal sorted keys per segment and classifies each run as final / head partial /
tail partial; E2 stitches the runs that cross segment boundaries.  The model follows the kernels'
rules line by line and must update every distinct key exactly once with the full sum — the property
a non-linear fused optimizer (rowwise Adagrad) depends on.  CPU only, no oracle involved."""
import numpy as np
import pytest


def _model(keys, vals, seg_len):
    n = len(keys)
    n_seg = (n + seg_len - 1) // seg_len
    updates = {}                       # key -> list of sums applied (must end up with exactly one)
    head = [None] * n_seg
    tail = [None] * n_seg

    def apply(k, s):
        updates.setdefault(int(k), []).append(s)

    # E1
    for seg in range(n_seg):
        s0, s1 = seg * seg_len, min((seg + 1) * seg_len, n)
        head_cont = s0 > 0 and keys[s0 - 1] == keys[s0]
        tail_cont = s1 < n and keys[s1] == keys[s1 - 1]
        first_run, cur, acc = True, None, 0.0

        def flush(last):
            nonlocal first_run, acc
            if cur is None:
                return
            if first_run and head_cont:
                head[seg] = acc
            elif last and tail_cont:
                tail[seg] = acc
            else:
                apply(cur, acc)
            first_run, acc = False, 0.0

        for i in range(s0, s1):
            if keys[i] != cur:
                flush(False)
                cur = keys[i]
            acc += vals[i]
        flush(True)
    # E2
    for seg in range(n_seg):
        s0, s1 = seg * seg_len, min((seg + 1) * seg_len, n)
        if s1 >= n:
            continue
        kl = keys[s1 - 1]
        if keys[s1] != kl:
            continue
        if keys[s0] == kl and s0 > 0 and keys[s0 - 1] == kl:
            continue
        acc = tail[seg]
        assert acc is not None, "E2 expects a tail partial that E1 did not write"
        j = seg + 1
        while j * seg_len < n and keys[j * seg_len] == kl:
            assert head[j] is not None, "E2 expects a head partial that E1 did not write"
            acc += head[j]
            head[j] = "used"
            j += 1
        tail[seg] = "used"
        apply(kl, acc)
    assert all(h in (None, "used") for h in head), "a head partial was written but never consumed"
    assert all(t in (None, "used") for t in tail), "a tail partial was written but never consumed"
    return updates


@pytest.mark.parametrize("seg_len", [1, 2, 3, 8, 128])
@pytest.mark.parametrize("n_keys,n", [(1, 1), (1, 700), (3, 50), (5, 1000), (400, 1000), (1000, 300), (7, 128), (2, 256)])
def test_every_key_is_updated_once_with_its_full_sum(seg_len, n_keys, n):
    rng = np.random.default_rng(seg_len * 1000 + n_keys + n)
    # skewed draws so that some keys span many segments and others sit inside one
    keys = np.sort(np.minimum((rng.pareto(1.2, size=n)).astype(np.int64), n_keys - 1))
    vals = rng.integers(1, 100, size=n).astype(np.float64)     # integers: sums are exact in any order
    updates = _model(keys, vals, seg_len)
    want = {int(k): float(vals[keys == k].sum()) for k in np.unique(keys)}
    assert set(updates) == set(want)
    for k, sums in updates.items():
        assert len(sums) == 1, f"key {k} updated {len(sums)} times"
        assert sums[0] == want[k]


def test_runs_aligned_to_segment_boundaries():
    # runs that start and end exactly on segment boundaries, incl. a run of exactly k segments
    seg = 4
    keys = np.array([0] * 4 + [1] * 8 + [2] * 4 + [3] * 2 + [4] * 2 + [5] * 12 + [6])
    vals = np.arange(1, len(keys) + 1, dtype=np.float64)
    updates = _model(keys, vals, seg)
    for k in np.unique(keys):
        assert updates[int(k)] == [float(vals[keys == k].sum())]


# ---------------------------------------------------------------------------------------------------
# slot layout of the device-side sparse redistribution (param_b200/csrc/sparse_dist.cu)
# ---------------------------------------------------------------------------------------------------
def _regroup_slots_model(lengths_in, window, W, T, b, slot):
    """seg_sum_permute + seg_starts (slot form) + seg_copy, rule for rule"""
    seg_sum = lengths_in.reshape(W * T, b).sum(axis=1)
    in_start = np.zeros(W * T, np.int64)
    in_end = np.zeros(W * T, np.int64)
    for r in range(W):
        acc = r * slot
        for t in range(T):
            in_start[r * T + t] = acc
            acc += seg_sum[r * T + t]
            in_end[r * T + t] = acc
    out_start = np.zeros(W * T, np.int64)
    acc = 0
    for t in range(T):
        for r in range(W):
            out_start[r * T + t] = acc
            acc += seg_sum[r * T + t]
    out = np.full(W * slot, -1, np.int64)
    for i in range(W * slot):
        lo, hi = 0, W * T - 1
        while lo < hi:
            mid = (lo + hi + 1) >> 1
            if in_start[mid] <= i:
                lo = mid
            else:
                hi = mid - 1
        if in_start[lo] <= i < in_end[lo]:
            o = out_start[lo] + (i - in_start[lo])
            if o < W * slot:
                out[o] = window[i]
    return out


@pytest.mark.parametrize("W,T,b,max_len", [(2, 3, 4, 5), (4, 2, 3, 2), (3, 1, 7, 9), (8, 4, 2, 1)])
def test_slot_layout_regroup_model_matches_oracle(oracle, W, T, b, max_len):
    rng = np.random.default_rng(W * 100 + T * 10 + b)
    lengths = rng.integers(0, max_len + 1, size=W * T * b).astype(np.int64)
    lengths[rng.integers(0, W * T * b, size=3)] = 0
    if W > 2:
        lengths[T * b:2 * T * b] = 0                      # a source that sends nothing
    slot = T * b * max_len
    window = np.full(W * slot, -7, np.int64)              # stale content in the slot tails
    packed = []
    for r in range(W):
        n = int(lengths[r * T * b:(r + 1) * T * b].sum())
        blk = rng.integers(0, 1 << 40, size=n).astype(np.int64)
        window[r * slot:r * slot + n] = blk
        packed.append(blk)
    packed = np.concatenate(packed)
    _, want_off, want_idx = oracle.split_per_table(lengths, packed, W, T, b)
    got = _regroup_slots_model(lengths, window, W, T, b, slot)
    assert np.array_equal(got[:want_off[-1]], want_idx)


# ---------------------------------------------------------------------------------------------------
# prefetch bookkeeping of E1 (state / old weights are loaded with the batch, not at the flush)
# ---------------------------------------------------------------------------------------------------
def _e1_prefetch_model(keys, G, U):
    """One lane group walking one segment in G-blocks and U-batches like exact_reduce_kernel.  Every
    flush must find cur_state == the prefetch of the row it is about to update."""
    my_n = len(keys)
    cur_key, cur_state, flushed = None, None, []

    def flush():
        nonlocal cur_key
        if cur_key is not None:
            assert cur_state == ("state-of", cur_key), (cur_key, cur_state)
            flushed.append(cur_key)

    for base in range(0, my_n, G):
        valid = my_n - base
        cnt = min(G, my_n - base)
        for j0 in range(0, cnt, U):
            kk = [keys[base + j0 + u] if (j0 + u < valid and j0 + u < G) else 0 for u in range(U)]
            st = []
            for u in range(U):
                ok_u = (j0 + u < valid) and (j0 + u < G)
                ok_n = (u + 1 < U) and (j0 + u + 1 < valid) and (j0 + u + 1 < G)
                ends = ok_u and (not ok_n or kk[u + 1 if u + 1 < U else u] != kk[u])
                st.append(("state-of", kk[u]) if ends else None)
            all_valid = (j0 + U <= valid) and (j0 + U <= G)
            same = kk[U - 1] == kk[0]
            if all_valid and same and (kk[0] == cur_key or cur_key is None):
                cur_key = kk[0]
                cur_state = st[U - 1]
            else:
                for u in range(U):
                    if j0 + u < valid and j0 + u < G:
                        if kk[u] != cur_key:
                            flush()
                            cur_key = kk[u]
                        cur_state = st[u]
    flush()
    return flushed


@pytest.mark.parametrize("G,U", [(32, 8), (32, 4), (32, 2), (16, 8), (8, 8), (4, 8)])
def test_prefetched_state_always_belongs_to_the_flushed_row(G, U):
    rng = np.random.default_rng(G * 10 + U)
    for n in (1, 2, 7, 8, 9, 31, 32, 33, 100, 128):
        for n_keys in (1, 2, 5, 200):
            keys = np.sort(rng.integers(0, n_keys, size=n)).tolist()
            flushed = _e1_prefetch_model(keys, G, U)
            assert flushed == sorted(set(keys))
