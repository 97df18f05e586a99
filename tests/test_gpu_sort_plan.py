"""The sort plan of the backward (pb200_tbe_plan_build, param_b200/csrc/radix_sort.cu) against numpy:
the hand-written per-table radix sort must reproduce a STABLE sort of every table's lookups by row,
bit for bit — keys (arena rows), values (gradient-row offsets or positions) and the weighted side arrays —
for 1 to 4 radix passes, ragged and empty bags, tiles that span several sub-tiles, int32 / int64, and a
request whose offsets do not start at 0.  Then: a backward that consumes a plan built on a side stream
gives the same bits as the one that sorts inline."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _t(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(dev)


def _request(rng, rows, B, lmax, empty_frac=0.1, alpha=0.0):
    T = len(rows)
    lens = rng.integers(0, lmax + 1, size=T * B)
    lens[rng.random(T * B) < empty_frac] = 0
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.empty(int(offsets[-1]), np.int64)
    for t in range(T):
        lo, hi = offsets[t * B], offsets[(t + 1) * B]
        if alpha > 0:
            p = 1.0 / np.arange(1, rows[t] + 1) ** alpha
            idx[lo:hi] = rng.choice(rows[t], size=hi - lo, p=p / p.sum())
        else:
            idx[lo:hi] = rng.integers(0, rows[t], size=hi - lo)
    return offsets, idx


def _expected(rows, B, D, offsets, idx, layout):
    T = len(rows)
    tro = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
    n = idx.size
    bag_of = np.repeat(np.arange(T * B), np.diff(offsets))
    t_of, b_of = bag_of // B, bag_of % B
    st_t, st_b = (D, T * D) if layout == "BTD" else (B * D, D)
    goff = ((t_of * st_t + b_of * st_b) >> 2).astype(np.uint32)
    keys = np.empty(n, np.uint32)
    order = np.empty(n, np.int64)
    for t in range(T):
        lo, hi = offsets[t * B], offsets[(t + 1) * B]
        o = np.argsort(idx[lo:hi], kind="stable") + lo
        order[lo:hi] = o
        keys[lo:hi] = (idx[o] + tro[t]).astype(np.uint32)
    return tro, keys, goff, order


def _plan_arrays(plan, n):
    raw = plan.buf.cpu().numpy()
    arr = (n * 4 + 255) & ~255
    keys = raw[0:n * 4].view(np.uint32)
    vals = raw[arr:arr + n * 4].view(np.uint32)
    goff_of = raw[2 * arr:2 * arr + n * 4].view(np.uint32)
    w_of = raw[3 * arr:3 * arr + n * 4].view(np.float32)
    return keys, vals, goff_of, w_of


CASES = [
    # rows per table, batch, max bag, max_table_rows hint (0 = unknown -> 32-bit keys), Zipf alpha
    ([700, 1000, 13], 300, 9, 1000, 0.0),            # 10 bits: one pass
    ([50_000, 100_000, 7], 2500, 30, 100_000, 0.0),  # 17 bits: two passes, tiles of > 1 sub-tile
    ([100_000, 30_000], 4096, 24, 100_000, 1.15),    # two passes, heavy duplicates
    ([3_000_000, 5_000_000], 3000, 12, 5_000_000, 0.0),   # 23 bits: three passes
    ([40_000, 90_000, 90_000], 1111, 20, 0, 0.0),    # no hint: four passes over 32 bits
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
@pytest.mark.parametrize("layout", ["BTD", "TBD"])
def test_plan_is_a_stable_sort_by_row(cuda_device, case, idx_dtype, layout):
    from param_b200 import ops
    rows, B, lmax, hint, alpha = CASES[case]
    rng = np.random.default_rng(100 + case)
    D = 16
    offsets, idx = _request(rng, rows, B, lmax, alpha=alpha)
    tro, keys, goff, order = _expected(rows, B, D, offsets, idx, layout)
    n = idx.size
    plan = ops.tbe_plan(_t(tro, cuda_device), len(rows), D, _t(idx, cuda_device, idx_dtype),
                        _t(offsets, cuda_device, idx_dtype), B, hint, layout=layout)
    torch.cuda.synchronize()
    k, v, _, _ = _plan_arrays(plan, n)
    assert np.array_equal(k, keys), "keys are not the per-table sorted arena rows"
    assert np.array_equal(v, goff[order]), "values do not follow a stable sort"
    assert np.all(np.diff(k.astype(np.int64)) >= 0), "the whole array must be sorted"


@pytest.mark.parametrize("mode", ["sum", "mean"])
def test_plan_side_arrays_weighted_and_mean(cuda_device, mode):
    from param_b200 import ops
    rows, B, D = [5000, 800], 700, 32
    rng = np.random.default_rng(5)
    offsets, idx = _request(rng, rows, B, 17)
    tro, keys, goff, order = _expected(rows, B, D, offsets, idx, "BTD")
    n = idx.size
    psw = rng.standard_normal(n).astype(np.float32)
    plan = ops.tbe_plan(_t(tro, cuda_device), 2, D, _t(idx, cuda_device), _t(offsets, cuda_device), B, 5000,
                        mode=mode, per_sample_weights=_t(psw, cuda_device))
    torch.cuda.synchronize()
    k, v, goff_of, w_of = _plan_arrays(plan, n)
    assert np.array_equal(k, keys)
    assert np.array_equal(v.astype(np.int64), order)          # values are positions
    assert np.array_equal(goff_of, goff)
    lens = np.repeat(np.diff(offsets), np.diff(offsets)).astype(np.float32)
    want = psw * (np.float32(1.0) / lens) if mode == "mean" else psw
    np.testing.assert_array_equal(w_of, want.astype(np.float32))


def test_plan_with_offsets_not_starting_at_zero(cuda_device, oracle):
    """the host-buffer entry hands the kernels a shifted indices pointer and absolute offsets"""
    from param_b200 import ops
    rows, B, D = [3000, 3000], 400, 64
    rng = np.random.default_rng(8)
    offsets, idx = _request(rng, rows, B, 10)
    tro = np.array([0, 3000, 6000], np.int64)
    g = rng.standard_normal((B, 2 * D)).astype(np.float32)
    ref64 = oracle.tbe_bwd(6000, tro, D, idx, offsets, B, g, dtype=np.float64)
    shift = 12345
    big = torch.zeros(shift + idx.size, dtype=torch.int64, device=cuda_device)
    big[shift:] = _t(idx, cuda_device)
    # ops works on whole tensors: emulate the shifted view with a tensor that starts `shift` early
    from param_b200 import _cabi
    lib = _cabi.load()
    off = _t(offsets + shift, cuda_device)
    dst = torch.zeros((6000, D), device=cuda_device)
    sb = int(lib.pb200_tbe_bwd_scratch_bytes(idx.size, 2, B, 6000, _cabi.BWD_SORTED))
    scratch = torch.empty(sb, dtype=torch.uint8, device=cuda_device)
    gt = _t(g, cuda_device)
    rc = lib.pb200_tbe_bwd(dst.data_ptr(), _t(tro, cuda_device).data_ptr(), 2, D, big.data_ptr(), idx.size,
                           off.data_ptr(), B, _cabi.IDX_I64, None, 0, gt.data_ptr(), D, 2 * D, 1.0,
                           _cabi.BWD_SORTED, 3000, scratch.data_ptr(), sb, 0,
                           torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert np.abs(dst.cpu().numpy() - ref64).max() <= 1e-5 * np.abs(ref64).max()


@pytest.mark.parametrize("exact", [False, True])
def test_backward_with_a_plan_built_on_a_side_stream(cuda_device, exact):
    from param_b200 import ops
    rows, B, D = [20_000] * 6, 2048, 128
    rng = np.random.default_rng(21)
    offsets, idx = _request(rng, rows, B, 20, alpha=1.05)
    tro = _t(np.concatenate([[0], np.cumsum(rows)]).astype(np.int64), cuda_device)
    i_d, o_d = _t(idx, cuda_device), _t(offsets, cuda_device)
    g = torch.randn(B, 6 * D, device=cuda_device)
    algo = "exact" if exact else "sorted"
    inline = torch.zeros((sum(rows), D), device=cuda_device)
    ops.tbe_backward(inline, tro, 6, D, i_d, o_d, B, g, algo=algo, max_table_rows=20_000)
    side = torch.cuda.Stream(device=cuda_device)
    plan = ops.tbe_plan(tro, 6, D, i_d, o_d, B, 20_000, exact=exact, stream=side)
    planned = torch.zeros((sum(rows), D), device=cuda_device)
    ops.tbe_backward(planned, tro, 6, D, i_d, o_d, B, g, algo=algo, max_table_rows=20_000, plan=plan)
    torch.cuda.synchronize()
    if exact:
        assert torch.equal(inline, planned)          # deterministic: identical bits
    else:
        assert (inline - planned).abs().max() <= 1e-5 * inline.abs().max()
    # a plan of another request is refused
    from param_b200._cabi import PB200Error
    with pytest.raises(PB200Error):
        ops.tbe_backward(planned, tro, 6, D, i_d.clone(), o_d, B, g, algo=algo, plan=plan)


@pytest.mark.parametrize("bag", [1, 2, 20, 33])
def test_plan_fixed_size_bags(cuda_device, bag):
    """equal bag lengths take the multiply-high path for the bag of a position (the benchmark / DLRM case)"""
    from param_b200 import ops
    rows, B, D = [70_000, 9_000], 5000, 16
    rng = np.random.default_rng(bag)
    T = len(rows)
    offsets = (np.arange(T * B + 1) * bag).astype(np.int64)
    idx = np.concatenate([rng.integers(0, rows[t], size=B * bag) for t in range(T)]).astype(np.int64)
    tro, keys, goff, order = _expected(rows, B, D, offsets, idx, "BTD")
    plan = ops.tbe_plan(_t(tro, cuda_device), T, D, _t(idx, cuda_device), _t(offsets, cuda_device), B, 70_000)
    torch.cuda.synchronize()
    k, v, _, _ = _plan_arrays(plan, idx.size)
    assert np.array_equal(k, keys)
    assert np.array_equal(v, goff[order])


def test_backward_by_table_groups_equals_whole(cuda_device, oracle):
    """pb200_tbe_bwd_tables: the segmented reduce restricted to table groups (position ranges read from the
    offsets on the device; groups cut segments anywhere) — all groups together equal the whole backward and the
    float64 oracle, each group alone touches only its tables' rows."""
    from param_b200 import ops
    rows, B, D = [3000, 50, 7000, 900, 12], 700, 64
    rng = np.random.default_rng(77)
    offsets, idx = _request(rng, rows, B, 15, alpha=1.1)
    T = len(rows)
    tro_h = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
    tro, i_d, o_d = _t(tro_h, cuda_device), _t(idx, cuda_device), _t(offsets, cuda_device)
    g = torch.randn(B, T * D, device=cuda_device)
    ref64 = oracle.tbe_bwd(int(tro_h[-1]), tro_h, D, idx, offsets, B, g.cpu().numpy(), dtype=np.float64)
    plan = ops.tbe_plan(tro, T, D, i_d, o_d, B, max(rows))
    dst = torch.zeros((int(tro_h[-1]), D), device=cuda_device)
    for parts in (1, 2, 3):
        dst.zero_()
        for p_ in range(parts):
            lo, hi = ops.part_range(T, p_, parts)
            before = dst.clone()
            ops.tbe_backward_tables(dst, tro, T, D, i_d, o_d, B, g, plan, lo, hi)
            changed = (dst != before).any(dim=1).nonzero().view(-1)
            if changed.numel():
                assert int(changed.min()) >= tro_h[lo] and int(changed.max()) < tro_h[hi], (parts, p_)
        assert np.abs(dst.cpu().numpy() - ref64).max() <= 1e-5 * np.abs(ref64).max(), parts
