"""Parity of the sm_100a EmbeddingBag kernels (through the C ABI) against the CPU oracle and the
golden vectors generated from the reference.  Bit-exact for SUM pooling and all integer work;
fp32 tolerance 1e-5 relative (BASELINE.json north_star) elsewhere."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _golden_cases(golden_dir):
    d = np.load(golden_dir / "embbag_torch_cpu.npz")
    for k in range(int(d["n_cases"])):
        p = f"c{k}_"
        psw = d[p + "psw"]
        yield k, dict(weight=d[p + "weight"], indices=d[p + "indices"], offsets=d[p + "offsets"],
                      mode=str(d[p + "mode"]), psw=psw if psw.size else None, out=d[p + "out"],
                      grad_out=d[p + "grad_out"], grad_weight=d[p + "grad_weight"])


def _t(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(dev)


@pytest.mark.parametrize("algo", ["direct", "staged", "pipelined"])
@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
def test_forward_golden(cuda_device, golden_dir, algo, idx_dtype):
    from param_b200 import ops
    for k, c in _golden_cases(golden_dir):
        w = _t(c["weight"], cuda_device)
        out = ops.embedding_bag_forward(
            w, _t(c["indices"], cuda_device, idx_dtype), _t(c["offsets"], cuda_device, idx_dtype),
            mode=c["mode"], per_sample_weights=None if c["psw"] is None else _t(c["psw"], cuda_device),
            algo=algo).cpu().numpy()
        if c["mode"] == "sum" and c["psw"] is None:
            assert np.array_equal(out, c["out"]), f"case {k} not bit-exact"
        else:
            np.testing.assert_allclose(out, c["out"], rtol=RTOL, atol=1e-6, err_msg=f"case {k}")


@pytest.mark.parametrize("algo", ["atomic", "sorted"])
def test_backward_golden(cuda_device, golden_dir, algo):
    from param_b200 import ops
    for k, c in _golden_cases(golden_dir):
        rows, dim = c["weight"].shape
        idx = _t(c["indices"], cuda_device)
        off = _t(np.concatenate([c["offsets"], [c["indices"].size]]), cuda_device)
        n_bags = c["offsets"].size
        grad_w = torch.zeros((rows, dim), device=cuda_device)
        ro = torch.tensor([0, rows], dtype=torch.int64, device=cuda_device)
        ops.tbe_backward(grad_w, ro, 1, dim, idx, off, n_bags, _t(c["grad_out"], cuda_device),
                         layout="TBD", scale=1.0, mode=c["mode"],
                         per_sample_weights=None if c["psw"] is None else _t(c["psw"], cuda_device),
                         algo=algo)
        np.testing.assert_allclose(grad_w.cpu().numpy(), c["grad_weight"], rtol=RTOL, atol=1e-5,
                                   err_msg=f"case {k} ({algo})")


def _random_tbe(rng, T, B, dim, max_len, rows_lo=50, rows_hi=400, fixed_len=None):
    rows = rng.integers(rows_lo, rows_hi, size=T)
    tro = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
    arena = rng.standard_normal((int(tro[-1]), dim)).astype(np.float32)
    if fixed_len is None:
        lens = rng.integers(0, max_len + 1, size=T * B)
        lens[rng.integers(0, T * B, size=max(1, T * B // 10))] = 0
    else:
        lens = np.full(T * B, fixed_len)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.concatenate([rng.integers(0, rows[t], size=int(lens[t * B:(t + 1) * B].sum()))
                          for t in range(T)]).astype(np.int64) if offsets[-1] else np.zeros(0, np.int64)
    return rows, tro, arena, offsets, idx


@pytest.mark.parametrize("dim", [128, 64, 56, 256, 8])
@pytest.mark.parametrize("algo", ["direct", "staged", "pipelined"])
@pytest.mark.parametrize("layout", ["BTD", "TBD"])
def test_tbe_forward_vs_oracle_ragged(cuda_device, oracle, dim, algo, layout):
    from param_b200 import ops
    rng = np.random.default_rng(dim * 7 + len(algo))
    T, B = 5, 777
    rows, tro, arena, offsets, idx = _random_tbe(rng, T, B, dim, 45)
    ref = oracle.tbe_fwd(arena, tro, dim, idx, offsets, B, layout=layout)
    ar = ops.TableArena(_t(arena, cuda_device), _t(tro, cuda_device), list(rows), dim)
    out = ops.tbe_forward(ar, _t(idx, cuda_device), _t(offsets, cuda_device), B, layout=layout, algo=algo)
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("algo", ["direct", "staged", "pipelined"])
def test_tbe_forward_odd_alignment_and_tail(cuda_device, oracle, algo):
    """odd total index count, odd bag starts, int32 indices: exercises the 16 B alignment logic of
    the bulk-copy staging and its direct fallback on the last tile."""
    from param_b200 import ops
    rng = np.random.default_rng(5)
    for T, B, L in [(1, 33, 3), (3, 65, 7), (2, 64, 1), (1, 1, 5), (4, 31, 9)]:
        rows, tro, arena, offsets, idx = _random_tbe(rng, T, B, 128, 0, fixed_len=L)
        ref = oracle.tbe_fwd(arena, tro, 128, idx, offsets, B)
        ar = ops.TableArena(_t(arena, cuda_device), _t(tro, cuda_device), list(rows), 128)
        for dt in (torch.int64, torch.int32):
            out = ops.tbe_forward(ar, _t(idx, cuda_device, dt), _t(offsets, cuda_device, dt), B, algo=algo)
            assert np.array_equal(out.cpu().numpy(), ref), (T, B, L, dt)


def test_tbe_forward_mean_and_weighted(cuda_device, oracle):
    from param_b200 import ops
    rng = np.random.default_rng(11)
    T, B, dim = 3, 200, 64
    rows, tro, arena, offsets, idx = _random_tbe(rng, T, B, dim, 30)
    psw = rng.random(idx.size).astype(np.float32) + 0.5
    ar = ops.TableArena(_t(arena, cuda_device), _t(tro, cuda_device), list(rows), dim)
    for algo in ("direct", "staged", "pipelined"):
        out = ops.tbe_forward(ar, _t(idx, cuda_device), _t(offsets, cuda_device), B, mode="mean", algo=algo)
        np.testing.assert_allclose(out.cpu().numpy(), oracle.tbe_fwd(arena, tro, dim, idx, offsets, B, mode="mean"),
                                   rtol=RTOL, atol=1e-6)
        out = ops.tbe_forward(ar, _t(idx, cuda_device), _t(offsets, cuda_device), B,
                              per_sample_weights=_t(psw, cuda_device), algo=algo)
        np.testing.assert_allclose(out.cpu().numpy(), oracle.tbe_fwd(arena, tro, dim, idx, offsets, B, psw=psw),
                                   rtol=RTOL, atol=1e-5)


def test_empty_inputs(cuda_device):
    from param_b200 import ops
    w = torch.randn(10, 16, device=cuda_device)
    e = torch.zeros(0, dtype=torch.int64, device=cuda_device)
    assert ops.embedding_bag_forward(w, e, e).shape == (0, 16)
    off = torch.zeros(4, dtype=torch.int64, device=cuda_device)
    out = ops.embedding_bag_forward(w, e, off)
    assert out.shape == (4, 16) and float(out.abs().sum()) == 0.0


@pytest.mark.parametrize("algo", ["atomic", "sorted"])
@pytest.mark.parametrize("dim", [128, 64, 56])
def test_tbe_backward_vs_oracle(cuda_device, oracle, algo, dim):
    from param_b200 import ops
    rng = np.random.default_rng(dim + 3)
    T, B = 4, 500
    rows, tro, arena, offsets, idx = _random_tbe(rng, T, B, dim, 25, rows_lo=20, rows_hi=80)
    g = rng.standard_normal((B, T * dim)).astype(np.float32)
    ref64 = oracle.tbe_bwd(int(tro[-1]), tro, dim, idx, offsets, B, g, dtype=np.float64)
    dst = torch.zeros((int(tro[-1]), dim), device=cuda_device)
    ops.tbe_backward(dst, _t(tro, cuda_device), T, dim, _t(idx, cuda_device), _t(offsets, cuda_device),
                     B, _t(g, cuda_device), layout="BTD", scale=1.0, algo=algo)
    got = dst.cpu().numpy()
    scale = np.abs(ref64).max()
    assert np.abs(got - ref64).max() <= RTOL * scale
    # fused SGD form: W -= lr * dW, in place on the arena
    lr = 0.05
    w = _t(arena, cuda_device).clone()
    ops.tbe_backward(w, _t(tro, cuda_device), T, dim, _t(idx, cuda_device), _t(offsets, cuda_device),
                     B, _t(g, cuda_device), layout="BTD", scale=-lr, algo=algo)
    want = arena.astype(np.float64) - lr * ref64
    assert np.abs(w.cpu().numpy() - want).max() <= RTOL * np.abs(want).max()


def test_backward_zipf_hot_rows(cuda_device, oracle):
    """heavy duplication (the Zipf regime): many bags hit the same few rows."""
    from param_b200 import ops
    rng = np.random.default_rng(77)
    T, B, L, dim, rows = 2, 4096, 20, 128, 1000
    tro = np.array([0, rows, 2 * rows], np.int64)
    p = 1.0 / np.arange(1, rows + 1) ** 1.15
    p /= p.sum()
    idx = rng.choice(rows, size=T * B * L, p=p).astype(np.int64)
    offsets = np.arange(T * B + 1, dtype=np.int64) * L
    g = rng.standard_normal((T, B, dim)).astype(np.float32)
    ref64 = oracle.tbe_bwd(2 * rows, tro, dim, idx, offsets, B, g, layout="TBD", dtype=np.float64)
    for algo in ("atomic", "sorted"):
        dst = torch.zeros((2 * rows, dim), device=cuda_device)
        ops.tbe_backward(dst, _t(tro, cuda_device), T, dim, _t(idx, cuda_device), _t(offsets, cuda_device),
                         B, _t(g, cuda_device), layout="TBD", algo=algo)
        err = np.abs(dst.cpu().numpy() - ref64).max() / np.abs(ref64).max()
        assert err <= RTOL, (algo, err)


def test_check_indices(cuda_device):
    from param_b200 import ops
    tro = torch.tensor([0, 10, 30], dtype=torch.int64, device=cuda_device)
    idx = torch.tensor([0, 9, 10, 5, 19, 30, -1], dtype=torch.int64, device=cuda_device)
    off = torch.tensor([0, 2, 3, 5, 7], dtype=torch.int64, device=cuda_device)  # T=2, B=2
    assert ops.check_indices(tro, 2, idx, off, 2) == 3  # 10 (table 0), 30 and -1 (table 1)


def test_module_matches_reference_call_contract(cuda_device, oracle):
    """B200EmbeddingBag(features, embdim, mode="sum")(indices, offsets) — pytorch_emb.py:179,61 —
    plus autograd to a dense weight.grad."""
    from param_b200.compute.pt.pytorch_emb import B200EmbeddingBag, init_indices, make_offsets
    torch.manual_seed(0)
    feats, dim, nnz, batch = 5000, 64, 20, 512
    idx = init_indices(0.0, feats, batch, nnz)
    off = make_offsets(batch, nnz)
    emb = B200EmbeddingBag(feats, dim, mode="sum").to(cuda_device)
    out = emb(idx.to(cuda_device), off.to(cuda_device))
    ref = oracle.embbag_fwd(emb.weight.detach().cpu().numpy(), idx.numpy(), off.numpy())
    assert np.array_equal(out.detach().cpu().numpy(), ref)
    g = torch.randn_like(out)
    out.backward(g)
    ref_g = oracle.embbag_bwd(feats, dim, idx.numpy(), off.numpy(), g.cpu().numpy(), dtype=np.float64)
    assert np.abs(emb.weight.grad.cpu().numpy() - ref_g).max() <= RTOL * np.abs(ref_g).max()


def test_cpu_tensor_is_refused():
    from param_b200 import ops
    from param_b200._cabi import PB200Error
    with pytest.raises(PB200Error):
        ops.embedding_bag_forward(torch.randn(4, 4), torch.zeros(1, dtype=torch.int64),
                                  torch.zeros(1, dtype=torch.int64))


# ---- size-independent properties at a large shape (oracle would take too long) -----------------
def test_large_shape_properties(cuda_device):
    from param_b200 import ops
    from param_b200.compute.pt.pytorch_emb import init_indices
    T, B, L, dim, rows = 16, 65536, 20, 128, 200_000
    ar = ops.TableArena.allocate([rows] * T, dim, cuda_device)
    ops.fill_uniform_(ar.weights, -0.01, 0.01, seed=3)
    idx = torch.cat([init_indices(1.15, rows, B, L, compat=False, seed=100 + t, device=cuda_device)
                     for t in range(T)])
    off = torch.arange(T * B + 1, dtype=torch.int64, device=cuda_device) * L
    assert ops.check_indices(ar.row_offsets, T, idx, off, B) == 0
    # distinct inside every bag (reference generator's per-bag dedupe)
    srt = idx.view(T * B, L).sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    out_d = ops.tbe_forward(ar, idx, off, B, algo="direct")
    out_s = ops.tbe_forward(ar, idx, off, B, algo="staged")
    out_p = ops.tbe_forward(ar, idx, off, B, algo="pipelined")
    assert torch.equal(out_d, out_s) and torch.equal(out_d, out_p)       # variants agree bit for bit
    # linearity: lookup(2W) == 2 lookup(W) exactly (power-of-two scaling commutes with rounding)
    ar.weights.mul_(2.0)
    assert torch.equal(ops.tbe_forward(ar, idx, off, B), out_d * 2.0)
    ar.weights.mul_(0.5)
    # all-ones table: every pooled vector equals the bag length
    ones = ops.TableArena(torch.ones_like(ar.weights), ar.row_offsets, ar.rows, dim)
    assert bool((ops.tbe_forward(ones, idx, off, B) == float(L)).all())
    # checksum of checksums: sum of outputs == sum over lookups of row sums (float64)
    row_sum = ar.weights.double().sum(dim=1)
    base = ar.row_offsets[:-1].repeat_interleave(B * L)
    want = row_sum[idx + base].sum()
    got = out_d.double().sum()
    assert abs(float(got - want)) <= 1e-6 * float(row_sum[idx + base].abs().sum())
    # backward: round trip of counts — dW from an all-ones grad is the per-row hit count
    for algo in ("atomic", "sorted", "exact"):
        dst = torch.zeros_like(ar.weights)
        ops.tbe_backward(dst, ar.row_offsets, T, dim, idx, off, B, torch.ones(B, T * dim, device=cuda_device),
                         layout="BTD", algo=algo)
        counts = torch.bincount(idx + base, minlength=ar.total_rows).float()
        assert torch.equal(dst[:, 0], counts) and torch.equal(dst[:, dim - 1], counts), algo


def test_fill_uniform_is_counter_based(cuda_device):
    from param_b200 import ops
    a = ops.fill_uniform_(torch.empty(1001, device=cuda_device), -1.0, 1.0, seed=9)
    b = ops.fill_uniform_(torch.empty(4004, device=cuda_device), -1.0, 1.0, seed=9)
    assert torch.equal(a, b[:1001]) and float(a.min()) >= -1.0 and float(a.max()) <= 1.0
    assert abs(float(b.mean())) < 0.1


def test_host_buffer_entry(cuda_device, oracle):
    """C-ABI §7: host pointers in, host pointers out, H2D/D2H inside the call."""
    import ctypes as C
    from param_b200 import _cabi, ops
    rng = np.random.default_rng(21)
    T, B, L, dim = 6, 256, 5, 128
    rows, tro, arena, offsets, idx = _random_tbe(rng, T, B, dim, 0, fixed_len=L)
    ar = ops.TableArena(_t(arena, cuda_device), _t(tro, cuda_device), list(rows), dim)
    lib = _cabi.load()
    ctx = C.c_void_p()
    _cabi.check(lib.pb200_host_ctx_create(C.byref(ctx), 4 * B * L + 16, 4 * B, dim))
    h_idx = torch.from_numpy(idx).pin_memory()
    h_off = torch.from_numpy(offsets).pin_memory()
    tro_h = torch.from_numpy(tro)
    for layout, name in ((0, "BTD"), (1, "TBD")):
        h_out = torch.empty((B, T * dim) if layout == 0 else (T, B, dim)).pin_memory()
        _cabi.check(lib.pb200_tbe_fwd_host(ctx, ar.weights.data_ptr(), ar.row_offsets.data_ptr(),
                                           tro_h.data_ptr(), T, dim, h_idx.data_ptr(), idx.size,
                                           h_off.data_ptr(), B, 0, h_out.data_ptr(), layout, 4))
        assert np.array_equal(h_out.numpy(), oracle.tbe_fwd(arena, tro, dim, idx, offsets, B, layout=name))
    lib.pb200_host_ctx_destroy(ctx)


def test_host_buffer_step_loss_form(cuda_device, oracle):
    """C-ABI §7, loss form: indices/offsets from HOST memory, per-table sums of the pooled vectors back to HOST
    memory, the scatter-add applied to the resident arena; the full-output form gives the same arena."""
    import ctypes as C
    from param_b200 import _cabi, ops
    rng = np.random.default_rng(22)
    T, B, L, dim = 7, 192, 6, 64
    rows, tro, arena, offsets, idx = _random_tbe(rng, T, B, dim, 0, fixed_len=L)
    lib = _cabi.load()
    ctx = C.c_void_p()
    _cabi.check(lib.pb200_host_ctx_create(C.byref(ctx), 3 * B * L + 16, 3 * B, dim))
    h_idx = torch.from_numpy(idx).pin_memory()
    h_off = torch.from_numpy(offsets).pin_memory()
    tro_h = torch.from_numpy(tro)
    pooled = oracle.tbe_fwd(arena, tro, dim, idx, offsets, B, layout="TBD")          # [T, B, dim]
    want_loss = pooled.astype(np.float64).sum(axis=(1, 2))
    scale = -0.05
    want_w = oracle.tbe_bwd(arena.shape[0], tro, dim, idx, offsets, B, pooled, layout="TBD", scale=scale,
                            dtype=np.float64, dst=arena.astype(np.float64).copy())
    results = []
    for form in ("loss", "full"):
        ar = ops.TableArena(_t(arena, cuda_device), _t(tro, cuda_device), list(rows), dim)
        h_loss = torch.zeros(T, dtype=torch.float64).pin_memory()
        h_out = torch.empty((T, B, dim)).pin_memory()
        for do_bwd in (0, 1):      # forward only, then the step: the lookup precedes the update, so both see the same arena
            if form == "loss":
                _cabi.check(lib.pb200_tbe_step_host_loss(ctx, ar.weights.data_ptr(), ar.row_offsets.data_ptr(),
                                                         tro_h.data_ptr(), T, dim, h_idx.data_ptr(), idx.size,
                                                         h_off.data_ptr(), B, 0, h_loss.data_ptr(), 3, do_bwd,
                                                         C.c_float(scale)))
                assert np.allclose(h_loss.numpy(), want_loss, rtol=1e-12, atol=1e-9)
            else:
                _cabi.check(lib.pb200_tbe_step_host(ctx, ar.weights.data_ptr(), ar.row_offsets.data_ptr(),
                                                    tro_h.data_ptr(), T, dim, h_idx.data_ptr(), idx.size,
                                                    h_off.data_ptr(), B, 0, h_out.data_ptr(), 1, 3, do_bwd,
                                                    C.c_float(scale)))
                assert np.array_equal(h_out.numpy(), pooled)
        got = ar.weights.cpu().numpy()
        assert np.abs(got - want_w).max() <= 1e-5 * max(1.0, np.abs(want_w).max())
        results.append(got)
    assert np.array_equal(results[0], results[1])          # same kernels, same order: identical bits
    # argument errors: both or neither result pointer, dim not a multiple of 4
    assert lib.pb200_tbe_step_host_loss(ctx, ar.weights.data_ptr(), ar.row_offsets.data_ptr(), tro_h.data_ptr(), T,
                                        dim, h_idx.data_ptr(), idx.size, h_off.data_ptr(), B, 0, None, 3, 0,
                                        C.c_float(0.0)) == -1
    lib.pb200_host_ctx_destroy(ctx)


@pytest.mark.parametrize("shape,blocks", [((5, 33, 64), 5), ((257, 1000), 257), ((3, 4), 1), ((70000, 8), 70000)])
def test_pooled_sum_matches_float64_numpy(cuda_device, shape, blocks):
    from param_b200 import ops
    from param_b200._cabi import PB200Error
    rng = np.random.default_rng(blocks)
    x = rng.standard_normal(shape).astype(np.float32)
    got = ops.pooled_sum(_t(x, cuda_device), blocks).cpu().numpy()
    want = x.astype(np.float64).reshape(blocks, -1).sum(axis=1)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-10)
    again = ops.pooled_sum(_t(x, cuda_device), blocks).cpu().numpy()
    assert np.array_equal(got, again)                       # fixed order: run to run identical
    with pytest.raises(PB200Error):
        ops.pooled_sum(_t(x, cuda_device), blocks + 1 if x.size % (blocks + 1) else x.size + 1)
    with pytest.raises(PB200Error):
        ops.pooled_sum(torch.from_numpy(x), blocks)


def test_tbe_module_fused_sgd(cuda_device, oracle):
    """B200TBE: forward(indices, offsets, per_sample_weights) + out.backward(grad) with the optimizer
    fused into the backward (how pytorch_dist_backend.py:832-857 drives the TBE op)."""
    from param_b200.compute.tbe import B200TBE
    rng = np.random.default_rng(3)
    T, B, dim = 3, 128, 64
    specs = [(200, dim), (300, dim), (150, dim)]
    op = B200TBE(specs, lr=0.1, device=cuda_device)
    lens = rng.integers(0, 9, size=T * B)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.concatenate([rng.integers(0, specs[t][0], size=int(lens[t * B:(t + 1) * B].sum())) for t in range(T)]).astype(np.int64)
    w0 = op.weights.detach().cpu().numpy().copy()
    tro = op.arena.row_offsets.cpu().numpy()
    out = op.forward(_t(idx, cuda_device), _t(offsets, cuda_device), None)
    assert np.array_equal(out.detach().cpu().numpy(), oracle.tbe_fwd(w0, tro, dim, idx, offsets, B))
    g = torch.randn_like(out)
    out.backward(g)
    want = w0.astype(np.float64) + oracle.tbe_bwd(int(tro[-1]), tro, dim, idx, offsets, B, g.cpu().numpy(),
                                                  scale=-0.1, dtype=np.float64)
    assert np.abs(op.weights.detach().cpu().numpy() - want).max() <= RTOL * np.abs(want).max()


def test_et_replay_style_dispatch(cuda_device, oracle):
    """et_replay rebuilds an op from node.name + node.op_schema through TorchScript IR
    (et_replay/et_replay_utils.py:171-212); the b200:: ops must be callable that way."""
    import param_b200.et  # noqa: F401  (what the replay config's "import modules" does)
    ir = """
graph(%0: Tensor, %1: Tensor, %2: Tensor, %3: int, %4: Tensor?, %5: bool):
    %output: Tensor = b200::embedding_bag(%0, %1, %2, %3, %4, %5)
    return (%output)
"""
    fn = torch._C.CompilationUnit().create_function("b200::embedding_bag", torch._C.parse_ir(ir))
    rng = np.random.default_rng(8)
    w = rng.standard_normal((100, 128)).astype(np.float32)
    idx = rng.integers(0, 100, size=60).astype(np.int64)
    off = (np.arange(12) * 5).astype(np.int64)
    out = fn(_t(w, cuda_device), _t(idx, cuda_device), _t(off, cuda_device), 0, None, False)
    assert np.array_equal(out.cpu().numpy(), oracle.embbag_fwd(w, idx, off))
    tro = torch.tensor([0, 100], dtype=torch.int64, device=cuda_device)
    off1 = _t(np.concatenate([off, [60]]), cuda_device)
    out2 = torch.ops.b200.tbe_forward(_t(w, cuda_device), tro, 128, _t(idx, cuda_device), off1, 12, 0, None, 0)
    assert torch.equal(out2, out)
    dst = torch.zeros(100, 128, device=cuda_device)
    torch.ops.b200.tbe_backward_(dst, tro, 128, _t(idx, cuda_device), off1, 12, torch.ones(12, 128, device=cuda_device), 0, 1.0, 0, 1)
    assert torch.equal(dst[:, 0].cpu(), torch.bincount(torch.from_numpy(idx), minlength=100).float())


def test_emb_lookup_compute_function(cuda_device, oracle):
    """backendFuncs.emb_lookup over requests prepared by init_emb_lookup (the comms-side compute
    kernel, pytorch_dist_backend.py:832-857 + comms_utils.py:1956-2039), forward and backward."""
    import types
    from param_b200.comms.pt.backend import B200CommsMixin
    from param_b200.comms.pt.emb_lookup import init_emb_lookup

    class _Be(B200CommsMixin):
        def get_device(self):
            return cuda_device

    be = _Be()
    params = types.SimpleNamespace(direction="backward", emb_dim=64, num_embs=500, batch_size=32,
                                   num_emb_tables_per_device=4, num_emb_tables_batched=2, bag_size=5,
                                   device="cuda", emb_lr=0.5, emb_optimizer="exact_sgd")
    ca = types.SimpleNamespace(reuseTensors=True)    # retain_graph: the reference loops one backward per request
    init_emb_lookup(ca, params, be)
    assert ca.num_emb_ops == 2 and len(ca.emb) == 2 and len(ca.embRequests) == 2
    op, (idx, off, _) = ca.emb[1], ca.embRequests[1]
    first_w0 = ca.emb[0].weights.detach().cpu().numpy().copy()
    w0 = op.weights.detach().cpu().numpy().copy()
    tro = op.arena.row_offsets.cpu().numpy()
    want_out = oracle.tbe_fwd(w0, tro, 64, idx.cpu().numpy(), off.cpu().numpy(), 32)
    assert np.array_equal(ca.LookupOut.detach().cpu().numpy(), want_out)        # last op's forward
    be.emb_lookup(ca)                                                            # direction == backward
    want_w = w0.astype(np.float64) + oracle.tbe_bwd(int(tro[-1]), tro, 64, idx.cpu().numpy(), off.cpu().numpy(), 32,
                                                    ca.grad_output.cpu().numpy(), scale=-0.5, dtype=np.float64)
    got = op.weights.detach().cpu().numpy()
    # the reference loops the SAME LookupOut.backward once per request (pytorch_dist_backend.py:849-857, with
    # retain_graph): with two requests, two identical SGD steps land on the LAST op — one answer, w0 - 2 lr dW
    want_w2 = want_w + (want_w - w0.astype(np.float64))
    assert np.abs(got - want_w2).max() <= RTOL * np.abs(want_w2).max()
    first = ca.emb[0].weights.detach().cpu().numpy()     # ... and none on the first op
    assert np.array_equal(first, first_w0)
    ca.direction = "forward"
    be.emb_lookup(ca)
    assert ca.LookupOut.shape == (32, 2 * 64)


def test_operator_plugin_runs(cuda_device, oracle):
    """train/compute/python OperatorInterface protocol: build -> forward -> create_grad -> backward"""
    from param_b200.comms.pt.emb_lookup import generate_requests
    from param_b200.compute.operator import B200BatchedEmbeddingBagOp
    op = B200BatchedEmbeddingBagOp()
    op.device = str(cuda_device)
    op.build(3, 400, 128, pooling=0, weighted=False, weights_precision="fp32", optimizer="exact_sgd", lr=0.1)
    (idx, off, w), = generate_requests(1, 64, 3, 10, 400, alpha=1.15, seed=2, device=cuda_device)
    w0 = op.op.weights.detach().cpu().numpy().copy()
    tro = op.op.arena.row_offsets.cpu().numpy()
    out = op.forward(idx, off, w)
    assert np.array_equal(out.detach().cpu().numpy(), oracle.tbe_fwd(w0, tro, 128, idx.cpu().numpy(), off.cpu().numpy(), 64))
    op.create_grad()
    op.backward()
    want = w0.astype(np.float64) + oracle.tbe_bwd(int(tro[-1]), tro, 128, idx.cpu().numpy(), off.cpu().numpy(), 64,
                                                  np.ones((64, 3 * 128), np.float32), scale=-0.1, dtype=np.float64)
    assert np.abs(op.op.weights.detach().cpu().numpy() - want).max() <= RTOL * np.abs(want).max()
    op.cleanup()                    # outputs go, the built op stays (several inputs may run against one build)
    assert op.fwd_out is None and op.grad_in is None and op.op is not None
    out2 = op.forward(idx, off, w)
    assert out2.shape == out.shape


def test_sparse_gradient_matches_golden_and_allocates_no_dense_buffer(cuda_device, golden_dir):
    """nn.EmbeddingBag(sparse=True) semantics (what the reference's alloc_embedding_tables builds,
    pytorch_dist_backend.py:923-934): an uncoalesced COO gradient with one entry per lookup whose dense form is
    the golden dW; the module returns it when constructed with sparse=True."""
    from param_b200 import ops
    from param_b200.compute.pt.pytorch_emb import B200EmbeddingBag
    for k, c in _golden_cases(golden_dir):
        rows, dim = c["weight"].shape
        psw = None if c["psw"] is None else _t(c["psw"], cuda_device)
        g = ops.embedding_bag_backward_sparse(_t(c["grad_out"], cuda_device), _t(c["indices"], cuda_device),
                                              _t(c["offsets"], cuda_device), rows, mode=c["mode"],
                                              per_sample_weights=psw)
        assert g.is_sparse and g._nnz() == c["indices"].size and tuple(g.shape) == (rows, dim)
        np.testing.assert_allclose(g.to_dense().cpu().numpy(), c["grad_weight"], rtol=RTOL, atol=1e-5,
                                   err_msg=f"case {k}")
        emb = B200EmbeddingBag(rows, dim, mode=c["mode"], sparse=True, _weight=_t(c["weight"], cuda_device))
        out = emb(_t(c["indices"], cuda_device), _t(c["offsets"], cuda_device), psw)
        out.backward(_t(c["grad_out"], cuda_device))
        gw = emb.weight.grad          # COO when the table dwarfs the batch (rows > 8 x lookups), dense otherwise
        assert gw.is_sparse == (rows > 8 * c["indices"].size)
        gw = gw.to_dense() if gw.is_sparse else gw
        np.testing.assert_allclose(gw.cpu().numpy(), c["grad_weight"], rtol=RTOL, atol=1e-5)
    big = B200EmbeddingBag(100_000, 8, mode="sum", sparse=True, device=cuda_device)
    big(torch.tensor([5, 7, 99_999], device=cuda_device), torch.tensor([0, 1], device=cuda_device)).sum().backward()
    assert big.weight.grad.is_sparse and big.weight.grad._nnz() == 3
    # int32 indices, a dim that is not a multiple of 4 (scalar path), mean pooling
    rng = np.random.default_rng(3)
    idx = rng.integers(0, 50, size=200).astype(np.int32)
    off = np.sort(rng.integers(0, 200, size=30)).astype(np.int32)
    off[0] = 0
    go = rng.standard_normal((30, 6)).astype(np.float32)
    g = ops.embedding_bag_backward_sparse(_t(go, cuda_device), _t(idx, cuda_device), _t(off, cuda_device), 50, mode="mean")
    ref = torch.nn.functional.embedding_bag(torch.from_numpy(idx.astype(np.int64)), w := torch.zeros(50, 6, requires_grad=True),
                                            torch.from_numpy(off.astype(np.int64)), mode="mean")
    ref.backward(torch.from_numpy(go))
    np.testing.assert_allclose(g.to_dense().cpu().numpy(), w.grad.numpy(), rtol=RTOL, atol=1e-6)
