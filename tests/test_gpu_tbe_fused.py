"""Parity of the fused-optimizer ("exact") backward and of the fp16-table kernels, through the C ABI,
against the CPU oracle.  Tolerances: fp32 tables 1e-5 relative (north_star) against a float64 oracle;
fp16 tables: the forward is bit-exact against the oracle run on the table converted to fp32, the
updated weights are within one fp16 ulp (2^-10 relative) of the float64 result."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5
F16_ULP = 2.0 ** -10


def _t(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(dev)


def _request(rng, T, B, dim, rows, max_len=12, zipf=None, fixed_len=None):
    rows = np.asarray(rows)
    tro = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
    arena = rng.uniform(-1, 1, size=(int(tro[-1]), dim)).astype(np.float32)
    lens = np.full(T * B, fixed_len) if fixed_len is not None else rng.integers(0, max_len + 1, size=T * B)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    parts = []
    for t in range(T):
        n = int(lens[t * B:(t + 1) * B].sum())
        if zipf:
            p = 1.0 / np.arange(1, rows[t] + 1) ** zipf
            parts.append(rng.choice(rows[t], size=n, p=p / p.sum()))
        else:
            parts.append(rng.integers(0, rows[t], size=n))
    idx = np.concatenate(parts).astype(np.int64)
    return tro, arena, offsets, idx


def _rel(got, want):
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)


@pytest.mark.parametrize("dim", [128, 64, 56, 256, 8, 512])
@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
def test_exact_sgd_vs_oracle(cuda_device, oracle, dim, idx_dtype):
    """few rows, many lookups: every row's run crosses several 128-entry segments"""
    from param_b200 import ops
    rng = np.random.default_rng(dim)
    T, B = 3, 96
    tro, arena, offsets, idx = _request(rng, T, B, dim, [2, 300, 40], max_len=14)
    g = rng.standard_normal((B, T * dim)).astype(np.float32)
    want, _ = oracle.fused_optimizer_step(
        arena, oracle.tbe_bwd(int(tro[-1]), tro, dim, idx, offsets, B, g, dtype=np.float64), "exact_sgd", lr=0.3)
    w = _t(arena, cuda_device)
    args = (_t(tro, cuda_device), T, dim, _t(idx, cuda_device, idx_dtype), _t(offsets, cuda_device, idx_dtype), B,
            _t(g, cuda_device))
    ops.tbe_backward_fused(w, *args, optimizer="exact_sgd", lr=0.3)
    assert _rel(w.cpu().numpy(), want) <= RTOL
    # deterministic: a second run from the same start gives the same bits
    w2 = _t(arena, cuda_device)
    ops.tbe_backward_fused(w2, *args, optimizer="exact_sgd", lr=0.3)
    assert torch.equal(w, w2)


def test_exact_algo_of_tbe_backward_matches_golden_and_sorted(cuda_device, golden_dir, oracle):
    """pb200_tbe_bwd(algo = EXACT): the dense gradient of the torch-CPU goldens, and == SORTED to 1e-5"""
    from param_b200 import ops
    d = np.load(golden_dir / "embbag_torch_cpu.npz")
    for k in range(int(d["n_cases"])):
        p = f"c{k}_"
        weight, indices, offsets = d[p + "weight"], d[p + "indices"], d[p + "offsets"]
        psw, mode = d[p + "psw"], str(d[p + "mode"])
        rows, dim = weight.shape
        if dim % 4:
            continue
        off = _t(np.concatenate([offsets, [indices.size]]), cuda_device)
        ro = torch.tensor([0, rows], dtype=torch.int64, device=cuda_device)
        outs = {}
        for algo in ("exact", "sorted"):
            gw = torch.zeros((rows, dim), device=cuda_device)
            ops.tbe_backward(gw, ro, 1, dim, _t(indices, cuda_device), off, offsets.size,
                             _t(d[p + "grad_out"], cuda_device), layout="TBD", scale=1.0, mode=mode,
                             per_sample_weights=_t(psw, cuda_device) if psw.size else None, algo=algo)
            outs[algo] = gw.cpu().numpy()
        np.testing.assert_allclose(outs["exact"], d[p + "grad_weight"], rtol=RTOL, atol=1e-5, err_msg=f"case {k}")
        np.testing.assert_allclose(outs["exact"], outs["sorted"], rtol=RTOL, atol=1e-5, err_msg=f"case {k}")


@pytest.mark.parametrize("mode,weighted", [("sum", False), ("mean", False), ("sum", True)])
@pytest.mark.parametrize("dim", [128, 64])
def test_rowwise_adagrad_two_steps_vs_oracle(cuda_device, oracle, mode, weighted, dim):
    from param_b200 import ops
    rng = np.random.default_rng(11 + dim)
    T, B, lr, eps = 4, 128, 0.05, 1e-8
    rows = [50, 900, 200, 13]
    tro = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
    w_want = rng.uniform(-1, 1, size=(int(tro[-1]), dim))
    m_want = None
    w = _t(w_want.astype(np.float32), cuda_device)
    w_want = w.cpu().numpy().astype(np.float64)
    state = torch.zeros(int(tro[-1]), device=cuda_device)
    for step in range(2):
        _, _, offsets, idx = _request(rng, T, B, dim, rows, max_len=10, zipf=1.15 if step else None)
        psw = rng.uniform(0.5, 1.5, size=idx.size).astype(np.float32) if weighted else None
        g = rng.standard_normal((B, T * dim)).astype(np.float32)
        dense = oracle.tbe_bwd(int(tro[-1]), tro, dim, idx, offsets, B, g, mode=mode, psw=psw, dtype=np.float64)
        w_want, m_want = oracle.fused_optimizer_step(w_want, dense, "exact_row_wise_adagrad", lr=lr, eps=eps,
                                                     state=m_want)
        ops.tbe_backward_fused(w, _t(tro, cuda_device), T, dim, _t(idx, cuda_device), _t(offsets, cuda_device), B,
                               _t(g, cuda_device), optimizer="exact_row_wise_adagrad", lr=lr, eps=eps, state=state,
                               mode=mode, per_sample_weights=None if psw is None else _t(psw, cuda_device))
        assert _rel(state.cpu().numpy(), m_want) <= RTOL, f"state, step {step}"
        assert _rel(w.cpu().numpy(), w_want) <= RTOL, f"weights, step {step}"


def test_exact_zipf_hot_rows_span_hundreds_of_segments(cuda_device, oracle):
    """cfg2-shaped skew at a size the oracle finishes in seconds: the hottest row of each table holds
    ~15 % of 80 K lookups = ~100 segments; TBD gradient layout; multiple chunks are not needed here"""
    from param_b200 import ops
    rng = np.random.default_rng(21)
    T, B, L, dim = 2, 4096, 20, 128
    rows = [20000, 5000]
    tro, arena, offsets, idx = _request(rng, T, B, dim, rows, zipf=1.15, fixed_len=L)
    g = rng.standard_normal((T, B, dim)).astype(np.float32)
    dense = oracle.tbe_bwd(int(tro[-1]), tro, dim, idx, offsets, B, g, layout="TBD", dtype=np.float64)
    for optimizer in ("exact_sgd", "exact_row_wise_adagrad"):
        want, m_want = oracle.fused_optimizer_step(arena, dense, optimizer, lr=0.01, eps=1e-8)
        w = _t(arena, cuda_device)
        state = torch.zeros(int(tro[-1]), device=cuda_device)
        ops.tbe_backward_fused(w, _t(tro, cuda_device), T, dim, _t(idx, cuda_device), _t(offsets, cuda_device), B,
                               _t(g, cuda_device), optimizer=optimizer, lr=0.01, eps=1e-8, state=state, layout="TBD")
        assert _rel(w.cpu().numpy(), want) <= RTOL, optimizer
        if m_want is not None:
            assert _rel(state.cpu().numpy(), m_want) <= RTOL
            # rows that no lookup touched keep a zero state and their weights, bit for bit
            untouched = np.ones(int(tro[-1]), bool)
            untouched[np.concatenate([idx[offsets[t * B]:offsets[(t + 1) * B]] + tro[t] for t in range(T)])] = False
            assert not state.cpu().numpy()[untouched].any()
            assert np.array_equal(w.cpu().numpy()[untouched], arena[untouched])


def test_multi_chunk_backward_of_ones_counts_hits_exactly(cuda_device):
    """size-independent property at a size that needs SEVERAL chunks (tables are grouped into chunks
    of <= 2^24 rows once a chunk holds >= 8 M lookups; chunk i+1 is built and sorted on a side stream
    under the reduce of chunk i): grad = ones into a zeroed fp32 buffer -> every element of a row
    equals its hit count, exactly, for the EXACT and the SORTED variant."""
    from param_b200 import ops
    T, B, L, dim, rows = 4, 1 << 19, 8, 32, 9_000_000
    tro = torch.arange(T + 1, dtype=torch.int64, device=cuda_device) * rows
    g = torch.Generator(device=cuda_device)
    g.manual_seed(3)
    n = T * B * L
    # a quarter of the lookups hit 50 hot rows per table (runs of ~20 K entries = ~160 segments),
    # the rest are spread over the table
    idx = torch.randint(0, rows, (n,), generator=g, device=cuda_device)
    hot = torch.rand(n, generator=g, device=cuda_device) < 0.25
    idx[hot] = idx[hot] % 50
    off = torch.arange(T * B + 1, dtype=torch.int64, device=cuda_device) * L
    table_of = torch.arange(n, device=cuda_device) // (B * L)
    counts = torch.bincount(idx + table_of * rows, minlength=T * rows).to(torch.float32)
    ones = torch.ones((B, T * dim), device=cuda_device)
    for algo in ("exact", "sorted"):
        w = torch.zeros((T * rows, dim), device=cuda_device)
        ops.tbe_backward(w, tro, T, dim, idx, off, B, ones, scale=1.0, algo=algo)
        assert torch.equal(w[:, 0], counts) and torch.equal(w[:, dim - 1], counts), algo
        del w
    # rowwise Adagrad on the same request: state = mean_d(count^2) = count^2 exactly (counts < 2^10: every partial sum k*count^2 fits 24 bits)
    w = torch.zeros((T * rows, dim), device=cuda_device)
    state = torch.zeros(T * rows, device=cuda_device)
    ops.tbe_backward_fused(w, tro, T, dim, idx, off, B, ones, optimizer="exact_row_wise_adagrad", lr=1.0,
                           eps=0.0, state=state)
    small = counts < 1024
    assert torch.equal(state[small], (counts * counts)[small])
    touched = counts > 0
    # w = -lr * g / sqrt(g^2) = -1 on every touched row, 0 elsewhere
    assert torch.allclose(w[touched][:, 0], torch.full((int(touched.sum()),), -1.0, device=cuda_device), rtol=1e-5)
    assert not w[~touched].any()


@pytest.mark.parametrize("dim", [128, 64, 8])
def test_fp16_forward_is_bit_exact_vs_oracle_on_converted_table(cuda_device, oracle, dim):
    from param_b200 import ops
    rng = np.random.default_rng(31 + dim)
    T, B = 3, 200
    rows = [100, 333, 57]
    tro, arena, offsets, idx = _request(rng, T, B, dim, rows, max_len=25)
    a16 = arena.astype(np.float16)
    ar = ops.TableArena(_t(a16, cuda_device), _t(tro, cuda_device), rows, dim)
    psw = rng.uniform(0.5, 1.5, size=idx.size).astype(np.float32)
    for layout in ("BTD", "TBD"):
        out = ops.tbe_forward(ar, _t(idx, cuda_device), _t(offsets, cuda_device), B, layout=layout)
        assert np.array_equal(out.cpu().numpy(),
                              oracle.tbe_fwd(a16.astype(np.float32), tro, dim, idx, offsets, B, layout=layout))
    out = ops.tbe_forward(ar, _t(idx, cuda_device, torch.int32), _t(offsets, cuda_device, torch.int32), B,
                          mode="mean", per_sample_weights=None)
    np.testing.assert_allclose(out.cpu().numpy(),
                               oracle.tbe_fwd(a16.astype(np.float32), tro, dim, idx, offsets, B, mode="mean"),
                               rtol=RTOL, atol=1e-6)
    out = ops.tbe_forward(ar, _t(idx, cuda_device), _t(offsets, cuda_device), B, per_sample_weights=_t(psw, cuda_device))
    np.testing.assert_allclose(out.cpu().numpy(),
                               oracle.tbe_fwd(a16.astype(np.float32), tro, dim, idx, offsets, B, psw=psw),
                               rtol=RTOL, atol=1e-5)


@pytest.mark.parametrize("optimizer", ["exact_sgd", "exact_row_wise_adagrad"])
def test_fp16_tables_fused_update(cuda_device, oracle, optimizer):
    from param_b200 import ops
    rng = np.random.default_rng(41)
    T, B, dim, lr = 3, 160, 64, 0.1
    rows = [40, 500, 9]
    tro, arena, offsets, idx = _request(rng, T, B, dim, rows, max_len=10)
    a16 = arena.astype(np.float16)
    g = rng.standard_normal((B, T * dim)).astype(np.float32)
    dense = oracle.tbe_bwd(int(tro[-1]), tro, dim, idx, offsets, B, g, dtype=np.float64)
    want, m_want = oracle.fused_optimizer_step(a16.astype(np.float64), dense, optimizer, lr=lr, eps=1e-8)
    common = (_t(tro, cuda_device), T, dim, _t(idx, cuda_device), _t(offsets, cuda_device), B, _t(g, cuda_device))
    # round to nearest: within half an fp16 ulp of the float64 result (+ fp32 accumulation noise)
    w = _t(a16, cuda_device)
    state = torch.zeros(int(tro[-1]), device=cuda_device)
    ops.tbe_backward_fused(w, *common, optimizer=optimizer, lr=lr, eps=1e-8, state=state)
    got = w.cpu().numpy().astype(np.float64)
    assert np.all(np.abs(got - want) <= 0.5 * F16_ULP * np.maximum(np.abs(want), 2.0 ** -14) * 1.01 + 1e-6)
    if m_want is not None:
        assert _rel(state.cpu().numpy(), m_want) <= RTOL
    # stochastic rounding: within one fp16 ulp, unbiased on average, and not equal to RN everywhere
    w_sr = _t(a16, cuda_device)
    state.zero_()
    ops.tbe_backward_fused(w_sr, *common, optimizer=optimizer, lr=lr, eps=1e-8, state=state,
                           stochastic_rounding=True, sr_seed=1234)
    got_sr = w_sr.cpu().numpy().astype(np.float64)
    touched = np.abs(dense).sum(axis=1) > 0
    err = (got_sr - want)[touched]
    assert np.all(np.abs(err) <= F16_ULP * np.maximum(np.abs(want[touched]), 2.0 ** -14) * 1.01 + 1e-6)
    assert abs(err.mean()) <= 0.05 * F16_ULP        # mean error of ~10^4 elements, each |err| < ulp
    assert np.array_equal(got_sr[~touched], a16.astype(np.float64)[~touched])
    assert not np.array_equal(got_sr, got)


def test_tbe_module_rowwise_adagrad_and_fp16(cuda_device, oracle):
    """B200TBE as comms_utils.py:1995-2017 builds the fbgemm op: optimizer=EXACT_ROWWISE_ADAGRAD; and
    the compute/python operator plugin with weights_precision=fp16."""
    from param_b200.compute.operator import B200BatchedEmbeddingBagOp
    from param_b200.compute.tbe import B200TBE
    rng = np.random.default_rng(51)
    T, B, dim = 3, 64, 128
    specs = [(300, dim), (120, dim), (77, dim)]
    op = B200TBE(specs, optimizer="exact_row_wise_adagrad", learning_rate=0.05, eps=1e-8, device=cuda_device)
    assert op.bwd_algo == "exact" and op.momentum1 is not None
    tro = op.arena.row_offsets.cpu().numpy()
    w_want, m_want = op.weights.detach().cpu().numpy().astype(np.float64), None
    for _ in range(2):
        _, _, offsets, idx = _request(rng, T, B, dim, [r for r, _ in specs], max_len=9)
        out = op.forward(_t(idx, cuda_device), _t(offsets, cuda_device), None)
        assert np.array_equal(out.detach().cpu().numpy(),
                              oracle.tbe_fwd(w_want.astype(np.float32), tro, dim, idx, offsets, B))
        g = torch.randn_like(out)
        out.backward(g)
        dense = oracle.tbe_bwd(int(tro[-1]), tro, dim, idx, offsets, B, g.cpu().numpy(), dtype=np.float64)
        w_want, m_want = oracle.fused_optimizer_step(w_want, dense, "exact_row_wise_adagrad", lr=0.05, eps=1e-8,
                                                     state=m_want)
        # the next forward reads the fp32 weights the kernel wrote: compare and re-base on them
        got = op.weights.detach().cpu().numpy().astype(np.float64)
        assert _rel(got, w_want) <= RTOL
        assert _rel(op.momentum1.cpu().numpy(), m_want) <= RTOL
        w_want = got
        m_want = op.momentum1.cpu().numpy().astype(np.float64)

    plug = B200BatchedEmbeddingBagOp()
    plug.device = str(cuda_device)
    plug.build(2, 500, 64, pooling=0, weighted=False, weights_precision="fp16",
               optimizer="exact_row_wise_adagrad", lr=0.1, eps=1e-8)
    assert plug.op.weights.dtype == torch.float16 and plug.op.stochastic_rounding
    _, _, offsets, idx = _request(rng, 2, 32, 64, [500, 500], fixed_len=6)
    w0 = plug.op.weights.detach().cpu().numpy()
    tro = plug.op.arena.row_offsets.cpu().numpy()
    out = plug.forward(_t(idx, cuda_device), _t(offsets, cuda_device), None)
    assert np.array_equal(out.detach().cpu().numpy(), oracle.tbe_fwd(w0.astype(np.float32), tro, 64, idx, offsets, 32))
    plug.backward()                 # create_grad: ones_like, as the reference does
    dense = oracle.tbe_bwd(int(tro[-1]), tro, 64, idx, offsets, 32, np.ones((32, 128), np.float32), dtype=np.float64)
    want, _ = oracle.fused_optimizer_step(w0.astype(np.float64), dense, "exact_row_wise_adagrad", lr=0.1, eps=1e-8)
    got = plug.op.weights.detach().cpu().numpy().astype(np.float64)
    assert np.all(np.abs(got - want) <= F16_ULP * np.maximum(np.abs(want), 2.0 ** -14) * 1.01 + 1e-6)
    plug.cleanup()


def test_fused_backward_argument_errors(cuda_device):
    from param_b200 import ops
    from param_b200._cabi import PB200Error
    w = torch.zeros((10, 6), device=cuda_device)            # dim % 4 != 0: no vector path
    ro = torch.tensor([0, 10], dtype=torch.int64, device=cuda_device)
    idx = torch.zeros(4, dtype=torch.int64, device=cuda_device)
    off = torch.tensor([0, 2, 4], dtype=torch.int64, device=cuda_device)
    with pytest.raises(PB200Error):
        ops.tbe_backward_fused(w, ro, 1, 6, idx, off, 2, torch.ones((2, 6), device=cuda_device))
    w = torch.zeros((10, 8), device=cuda_device)
    with pytest.raises(PB200Error):                          # Adagrad without a state
        ops.tbe_backward_fused(w, ro, 1, 8, idx, off, 2, torch.ones((2, 8), device=cuda_device),
                               optimizer="exact_row_wise_adagrad")
    with pytest.raises(PB200Error):
        ops.tbe_backward_fused(w, ro, 1, 8, idx, off, 2, torch.ones((2, 8), device=cuda_device), optimizer="adam")


def test_et_replay_style_dispatch_of_the_fused_backward(cuda_device, oracle):
    """et_replay rebuilds an op from node.name + node.op_schema through TorchScript IR
    (et_replay/et_replay_utils.py:171-212): b200::tbe_backward_fused_ must be callable that way."""
    import param_b200.et  # noqa: F401
    ir = """
graph(%0: Tensor, %1: Tensor?, %2: Tensor, %3: int, %4: Tensor, %5: Tensor, %6: int, %7: Tensor, %8: int,
      %9: int, %10: Tensor?, %11: int, %12: float, %13: float, %14: bool, %15: int):
    %output: Tensor = b200::tbe_backward_fused_(%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15)
    return (%output)
"""
    fn = torch._C.CompilationUnit().create_function("b200::tbe_backward_fused_", torch._C.parse_ir(ir))
    rng = np.random.default_rng(61)
    T, B, dim = 2, 40, 64
    tro, arena, offsets, idx = _request(rng, T, B, dim, [90, 30], max_len=7)
    g = rng.standard_normal((B, T * dim)).astype(np.float32)
    w = _t(arena, cuda_device)
    state = torch.zeros(int(tro[-1]), device=cuda_device)
    ret = fn(w, state, _t(tro, cuda_device), dim, _t(idx, cuda_device), _t(offsets, cuda_device), B,
             _t(g, cuda_device), 0, 0, None, 2, 0.05, 1e-8, False, 0)
    assert ret.data_ptr() == w.data_ptr()
    dense = oracle.tbe_bwd(int(tro[-1]), tro, dim, idx, offsets, B, g, dtype=np.float64)
    want, m_want = oracle.fused_optimizer_step(arena, dense, "exact_row_wise_adagrad", lr=0.05, eps=1e-8)
    assert _rel(w.cpu().numpy(), want) <= RTOL and _rel(state.cpu().numpy(), m_want) <= RTOL


def test_tbe_training_steps_match_torch_optim_golden(cuda_device, golden_dir):
    """tests/golden/tbe_optim_torch.npz — 3 steps of the reference's TBE benchmark loop (forward,
    create_grad = ones_like, backward with the fused optimizer; split_table_batched_embeddings_ops.py:311-324)
    computed with torch only (per-table nn.EmbeddingBag + torch.optim.SGD / Adagrad; with a ones gradient
    rowwise Adagrad == elementwise Adagrad).  B200TBE must follow it step for step: 1e-5 relative."""
    from param_b200.compute.tbe import B200TBE
    d = np.load(golden_dir / "tbe_optim_torch.npz")
    rows, dim, B = [int(r) for r in d["rows"]], int(d["dim"]), int(d["batch"])
    lr, eps = float(d["lr"]), float(d["eps"])
    for name, optimizer in (("sgd", "exact_sgd"), ("adagrad", "exact_row_wise_adagrad")):
        for bwd_algo in (("exact", "sorted") if name == "sgd" else ("exact",)):
            op = B200TBE([(r, dim) for r in rows], optimizer=optimizer, learning_rate=lr, eps=eps,
                         device=cuda_device, bwd_algo=bwd_algo)
            op.weights.copy_(_t(d["w0"], cuda_device))
            for s in range(int(d["steps"])):
                out = op.forward(_t(d[f"s{s}_indices"], cuda_device), _t(d[f"s{s}_offsets"], cuda_device), None)
                np.testing.assert_allclose(out.detach().cpu().numpy(), d[f"{name}_s{s}_out"], rtol=RTOL, atol=1e-6)
                out.backward(torch.ones_like(out))
                assert _rel(op.weights.cpu().numpy(), d[f"{name}_s{s}_w"]) <= RTOL, (name, bwd_algo, s)
            if name == "adagrad":
                assert _rel(op.momentum1.cpu().numpy(), d["adagrad_state"]) <= RTOL


def test_rowwise_adagrad_matches_torch_only_golden_on_nonconstant_gradients(cuda_device, golden_dir):
    """tests/golden/tbe_rowwise_adagrad_torch.npz — exact rowwise Adagrad with random incoming gradients and one
    weighted step, computed with torch only (autograd dense gradients + the published update in float64).
    B200TBE (pb200_tbe_bwd_fused) must follow it step for step, weights and state: 1e-5 relative."""
    from param_b200.compute.tbe import B200TBE
    d = np.load(golden_dir / "tbe_rowwise_adagrad_torch.npz")
    rows, dim = [int(r) for r in d["rows"]], int(d["dim"])
    op = B200TBE([(r, dim) for r in rows], optimizer="exact_row_wise_adagrad", learning_rate=float(d["lr"]),
                 eps=float(d["eps"]), device=cuda_device)
    op.weights.copy_(_t(d["w0"], cuda_device))
    for s in range(int(d["steps"])):
        psw = _t(d[f"s{s}_psw"], cuda_device) if d[f"s{s}_psw"].size else None
        out = op.forward(_t(d[f"s{s}_indices"], cuda_device), _t(d[f"s{s}_offsets"], cuda_device), psw)
        np.testing.assert_allclose(out.detach().cpu().numpy(), d[f"s{s}_out"], rtol=RTOL, atol=1e-6)
        out.backward(_t(d[f"s{s}_grad"], cuda_device))
        assert _rel(op.weights.cpu().numpy(), d[f"s{s}_w"]) <= RTOL, s
        assert _rel(op.momentum1.cpu().numpy(), d[f"s{s}_state"]) <= RTOL, s
