"""Host-side logic on the CPU: generators, size maths, split planning, the plugin surface, and a
2-rank gloo run of the sparse-input redistribution plan checked against the oracle."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT / "tests" / "golden"))   # make_golden._ref_paths (reference import helper)


def test_init_indices_compat_is_bit_identical_to_reference_golden(golden_dir):
    from param_b200.compute.pt.pytorch_emb import init_indices
    d = np.load(golden_dir / "init_indices_ref.npz")
    f, b, n = (int(x) for x in d["uniform_args"])
    torch.manual_seed(int(d["uniform_seed"]))
    assert torch.equal(init_indices(0.0, f, b, n), torch.from_numpy(d["uniform"]))
    f, b, n = (int(x) for x in d["zipf_args"])
    np.random.seed(int(d["zipf_seed"]))
    assert torch.equal(init_indices(float(d["zipf_alpha"]), f, b, n), torch.from_numpy(d["zipf"]))
    # the driver hands --alpha over as a str (driver.py:44-46): accepted here
    np.random.seed(int(d["zipf_seed"]))
    assert torch.equal(init_indices(str(float(d["zipf_alpha"])), f, b, n), torch.from_numpy(d["zipf"]))


def test_init_indices_compat_fails_like_the_reference_when_bag_cannot_be_filled():
    from param_b200.compute.pt.pytorch_emb import init_indices
    np.random.seed(0)
    with pytest.raises(ValueError):
        init_indices(3.0, 50, 64, 20)      # extreme skew: < 20 distinct values among 40 draws


def test_zipf_cdf_matches_reference_pmf():
    from param_b200.compute.pt.pytorch_emb import zipf_cdf
    cdf = zipf_cdf(1.15, 1000)
    pmf = np.power(np.arange(1, 1001, dtype=np.float64), -1.15)
    pmf /= pmf.sum()
    np.testing.assert_allclose(np.diff(np.concatenate([[0.0], cdf])), pmf, rtol=1e-12, atol=1e-15)
    assert cdf[-1] == 1.0


def test_offsets_and_datasets():
    from param_b200.compute.pt import dataset
    from param_b200.compute.pt.pytorch_emb import make_offsets
    assert make_offsets(4, 20).tolist() == [0, 20, 40, 60]
    assert len(dataset.emb_A) == 16 and len(dataset.emb_B) == 6
    assert dataset.emb_A[0] == (14000000, 128, 30, 512) and dataset.emb_A[-1] == (26000000, 128, 30, 65536)
    assert dataset.emb_B[0] == (4800000, 56, 34, 2048) and dataset.emb_B[-1] == (4800000, 56, 34, 65536)
    if REF.exists():
        sys.path.insert(0, str(REF / "train" / "compute" / "pt"))
        sys.dont_write_bytecode = True
        import importlib
        ref_ds = importlib.import_module("dataset")
        assert list(ref_ds.emb_A) == dataset.emb_A and list(ref_ds.emb_B) == dataset.emb_B


def test_module_surface_matches_nn_embeddingbag_contract():
    from param_b200._cabi import PB200Error
    from param_b200.compute.pt.pytorch_emb import B200EmbeddingBag
    emb = B200EmbeddingBag(100, 16, mode="sum")
    assert tuple(emb.weight.shape) == (100, 16) and emb.weight.element_size() == 4
    assert abs(float(emb.weight.std()) - 1.0) < 0.2                      # N(0,1) default init
    with pytest.raises(PB200Error):                                        # no CPU path
        emb(torch.zeros(4, dtype=torch.int64), torch.zeros(2, dtype=torch.int64))
    with pytest.raises(PB200Error):
        B200EmbeddingBag(10, 4, mode="max")


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present on this box")
def test_emb_driver_prints_the_reference_table(capsys, monkeypatch):
    """run() of the EmbeddingBag driver (train/compute/pt/pytorch_emb.py:208-234): with the measurement stubbed to the
    same numbers on both sides, the reference's run() and the mirror's print the same text."""
    import types
    from make_golden import _ref_paths
    _ref_paths()
    import pytorch_emb as ref_emb
    from param_b200.compute.pt import pytorch_emb as mine
    dataset = [(4800000, 56, 34, 2048), (1000000, 64, 20, 512), (1024, 8, 1, 1)]
    fake = {d: (0.0123 * (i + 1), d[3] * d[2] * d[1] * 4) for i, d in enumerate(dataset)}
    args = types.SimpleNamespace(steps=7)
    monkeypatch.setattr(ref_emb, "run_single", lambda a, f, e, n, b: fake[(f, e, n, b)])
    monkeypatch.setattr(mine, "run_single", lambda a, f, e, n, b: fake[(f, e, n, b)])
    ref_emb.run(args, dataset)
    want = capsys.readouterr().out
    mine.run(args, dataset)
    got = capsys.readouterr().out
    assert got == want and want.count("\n") == 3 + len(dataset)


def test_size_maths():
    from param_b200.comms.pt.comms import get_sizes, parsesize
    assert parsesize("1K") == 1024 and parsesize("1G") == 1 << 30 and parsesize("512") == 512
    assert parsesize("4MB") == 4 << 20
    assert get_sizes(1024, 8192, 2) == [1024, 2048, 4096, 8192]
    assert get_sizes(1024, 5000, 4) == [1024, 4096]


def test_split_helpers_match_reference():
    from param_b200.comms.pt.dlrm import lengths_exchange_splits, owner_slice, parse_embedding_sizes, split_lengths
    assert split_lengths(5, 3) == [2, 2, 1] and split_lengths(8, 8) == [1] * 8 and split_lengths(512, 8) == [64] * 8
    assert owner_slice(1, [2, 2, 1]) == slice(2, 4)
    assert parse_embedding_sizes("1000-2000-30") == [1000, 2000, 30]
    assert parse_embedding_sizes("1000000x4") == [1000000] * 4
    assert lengths_exchange_splits([2, 2, 1], 2, 4) == ([8, 8, 4], [4, 4, 4])
    if REF.exists():
        from make_golden import _ref_paths
        _ref_paths()
        import dlrm as ref_dlrm
        for n, w in [(5, 3), (64, 8), (7, 4), (3, 3)]:
            for r in range(w):
                my, splits = ref_dlrm.paramDLRM_Net.get_split_lengths_by_len(None, n, r, w)
                assert splits == split_lengths(n, w) and my == split_lengths(n, w)[r]
                ref_sl = ref_dlrm.paramDLRM_Net.get_slice_sparse(None, r, splits, w)
                assert (ref_sl.start, ref_sl.stop) == (owner_slice(r, splits).start, owner_slice(r, splits).stop)


def test_dlrm_report_regions_and_mlp_shapes_match_the_reference():
    """The runner's 21 region rows, MLP stage shapes and report layout (dlrm.py:400-411, 575-604, 961-1009,
    1015-1037, 1069-1177)."""
    from param_b200.comms.pt import dlrm as mine
    names = [r[0] for r in mine.REGIONS]
    assert len(names) == 21 and names[3] == "offset_xchg" and names[16] == "iter_time" and names[-1] == "iter_bwd_a2a"
    assert all(a in mine.MARKS and b in mine.MARKS for _, a, b in mine.REGIONS)
    assert mine.mlp_layer_shapes([13, 512, 256, 128]) == [[512, 13], [256, 512], [128, 256]]
    # 26 tables + the dense feature, bottom MLP ends at 128: 27*26/2 + 128 = 479 inputs to the top MLP
    assert mine.top_mlp_dims(26, [13, 512, 256, 128], [1024, 1]) == [479, 1024, 1]
    assert mine.top_mlp_dims(26, [13, 512, 256, 128], [1024, 1], interaction_itself=True)[0] == 27 * 28 // 2 + 128
    assert mine.top_mlp_dims(3, [4, 3, 2], [4, 2, 1], "cat")[0] == 8
    with pytest.raises(Exception):
        mine.top_mlp_dims(3, [4, 3, 2], [4, 2, 1], "sum")
    marks = {k: 0.001 * i for i, k in enumerate(mine.MARKS)}
    t = mine.region_times_us(marks)
    assert abs(t["iter_time"] - 16000.0) < 1e-6 and abs(t["offset_xchg"] - 1000.0) < 1e-6
    assert abs(t["iter_data_prep"] - 7000.0) < 1e-6 and abs(t["bwd_a2a"] - 1000.0) < 1e-6
    W, R, n = 2, len(mine.REGIONS), 4
    lat = torch.arange(W * R * n, dtype=torch.float64).view(W, R, n)
    mem = torch.zeros(W, R, n, dtype=torch.float64)
    mem[:, 3, :] = 4096
    rows_all, rows_mean = mine.percentile_rows(lat, mem)
    assert rows_all[3][0] == "offset_xchg" and rows_all[3][1] == 4096 and rows_all[0][1] == 0
    assert rows_all[3][2] == float(lat[:, 3, :].min())
    assert rows_all[3][3] == float(np.percentile(lat[:, 3, :].to(torch.float32).numpy(), 50))      # single precision, as the reference
    # running sum(p50) skips the iter_* rows, as the reference's does
    assert rows_all[16][6] == rows_all[15][6] == rows_all[20][6]
    assert rows_mean[5][3] == float(np.percentile(lat[:, 5, :].to(torch.float32).mean(dim=1).numpy(), 50))
    text = mine.format_report(n, rows_all)
    lines = [ln for ln in text.split("\n") if ln.strip()]
    assert lines[0].split()[:2] == ["iters", "region"] and "total_time" in lines[-1] and len(lines) == 23
    if REF.exists():
        from make_golden import _ref_paths
        _ref_paths()
        import dlrm as ref_dlrm
        bench = ref_dlrm.commsDLRMBench()
        timers = bench.initTimers()
        assert list(bench.measured_regions) == names and set(timers) == set(mine.MARKS)
        for name, a, b in mine.REGIONS:
            assert (bench.measured_regions[name]["start"], bench.measured_regions[name]["end"]) == (a, b)
        ref_mlp = ref_dlrm.paramDLRM_Net.create_mlp(None, 1, np.array([13, 512, 256, 128]))
        assert [list(map(int, x)) for x in ref_mlp] == mine.mlp_layer_shapes([13, 512, 256, 128])


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present on this box")
def test_dlrm_runner_flags_carry_the_reference_names(monkeypatch):
    """Every model / loop flag of param_b200.comms.pt.dlrm exists under the same name in the reference's runner
    (commsDLRMBench.readArgs, dlrm.py:657-742); what is added here is listed."""
    import argparse
    import inspect
    import re
    from make_golden import _ref_paths
    _ref_paths()
    import dlrm as ref_dlrm
    from param_b200.comms.pt import dlrm as mine
    ap = argparse.ArgumentParser()
    monkeypatch.setattr(sys, "argv", ["dlrm.py"])            # readArgs ends in parser.parse_args() (dlrm.py:733)
    ref_dlrm.commsDLRMBench().readArgs(ap)
    ref_flags = {s for a in ap._actions for s in a.option_strings}
    my_flags = set(re.findall(r'add_argument\("(--[a-z0-9-]+)"', inspect.getsource(mine._parse)))
    assert my_flags - ref_flags == {"--alpha", "--compare-nccl", "--json", "--lr", "--two-collective-dist",
                                    "--unfused-forward"}
    assert {"--arch-mlp-bot", "--arch-mlp-top", "--arch-interaction-op", "--arch-sparse-feature-size",
            "--arch-embedding-size", "--mini-batch-size", "--num-batches", "--warmup-batches",
            "--num-indices-per-lookup", "--num-indices-per-lookup-fixed", "--perf-debug"} <= my_flags & ref_flags
    a = mine._parse(["--arch-mlp-bot", "13-512-256-128", "--arch-mlp-top", "1024-1", "--perf-debug"])
    assert a.arch_mlp_bot == "13-512-256-128" and a.perf_debug and a.arch_interaction_op == "dot"


def test_dlrm_report_text_equals_what_the_reference_prints(golden_dir):
    """Fixture: the reference's own commsDLRMBench.reportBenchTime (dlrm.py:1011-1193) run on seeded samples for 1 and
    3 ranks (tests/golden/make_golden.py dlrm_report).  Both tables — percentiles over all samples, and over the
    per-rank means — must come out character for character."""
    from param_b200.comms.pt import dlrm as mine
    d = np.load(golden_dir / "dlrm_report_ref.npz")
    assert [str(n) for n in d["regions"]] == [r[0] for r in mine.REGIONS]
    for w in (1, 3):
        lat = torch.from_numpy(d[f"w{w}_lat"])
        mem = torch.from_numpy(d[f"w{w}_mem"]).to(torch.float64)
        warm = int(d[f"w{w}_warm"])
        rows_all, rows_mean = mine.percentile_rows(lat.to(torch.float64), mem[:, :, warm:])   # the runner gathers float64
        parts = [[ln.rstrip() for ln in part.split("\n") if ln.strip()] for part in str(d[f"w{w}_text"]).split("-" * 125)]
        ref_all, ref_mean = [p for p in parts if p]
        got_all = [ln.rstrip() for ln in mine.format_report(lat.shape[2], rows_all).split("\n") if ln.strip()]
        got_mean = [ln.rstrip() for ln in mine.format_report(lat.shape[2], rows_mean, header=False).split("\n") if ln.strip()]
        assert got_all == ref_all and got_mean == ref_mean, w


def test_sparse_batch_from_offsets_matches_reference_calculate_lengths(golden_dir):
    from param_b200.comms.pt.dlrm import SparseBatch
    d = np.load(golden_dir / "dlrm_sparse_ref.npz")
    feat = int(d["cl_feat"])
    offs = [torch.from_numpy(d[f"cl_off{f}"]) for f in range(feat)]
    idxs = [torch.from_numpy(d[f"cl_idx{f}"]) for f in range(feat)]
    sb = SparseBatch.from_offsets(offs, idxs, device="cpu")
    assert sb.count == feat and sb.batch_size == int(d["cl_batch"])
    assert np.array_equal(sb.lengths.numpy(), d["cl_lengths"])
    assert np.array_equal(sb.indices.numpy(), d["cl_indices"])


def test_backend_surface_is_complete():
    """every name the reference runners call on a backend exists on B200Backend"""
    from param_b200.comms.pt.backend import B200Backend
    needed = """all_to_all_single all_to_all all_to_allv all_reduce reduce broadcast all_gather all_gather_base
        reduce_scatter_base barrier sync_barrier complete_accel_ops device_sync wait noop gemm emb_lookup
        get_reduce_op get_mem_size getBusBW alloc_ones alloc_random alloc_empty alloc_embedding_tables
        clear_memory get_local_rank get_global_rank get_world_size get_local_size get_group_rank
        get_group_size get_device get_hw_device get_default_group get_groups get_num_pgs get_next_group
        set_device get_new_stream get_new_event get_current_stream switch_stream sync_stream
        initialize_backend initialize_groups initialize_tcpstore sayHello benchmark_comms set_up tear_down
        store_get store_set tensor_list_to_numpy barrier_all_ranks""".split()
    for name in needed:
        assert callable(getattr(B200Backend, name, None)), name
    be = B200Backend.__new__(B200Backend)
    ca = type("CA", (), {"world_size": 8})()
    assert abs(B200Backend.getBusBW(be, "all_to_all_single", 100.0, ca) - 87.5) < 1e-9
    assert abs(B200Backend.getBusBW(be, "all_reduce", 100.0, ca) - 175.0) < 1e-9


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present on this box")
def test_plugin_registers_with_the_reference_registry():
    from make_golden import _ref_paths
    _ref_paths()
    import inspect
    from param_b200.integration import param_plugin
    cls = param_plugin.register("b200")
    from param_bench.train.comms.pt import pytorch_backend_utils as pbu
    from param_bench.train.comms.pt.pytorch_dist_backend import PyTorchDistBackend
    assert pbu.customized_backend["b200"] is cls
    assert issubclass(cls, PyTorchDistBackend) and not inspect.isabstract(cls)
    # hot-path entries resolve to the B200 overrides, the rest to the reference's c10d code
    from param_b200.comms.pt.backend import B200CommsMixin
    for name in ("all_to_all_single", "all_to_allv", "all_to_all", "emb_lookup", "alloc_empty",
                 "alloc_embedding_tables", "complete_accel_ops"):
        assert getattr(cls, name) is getattr(B200CommsMixin, name), name
    assert cls.all_reduce is PyTorchDistBackend.all_reduce


# ---- world_size-2 gloo: the redistribution plan end to end on the CPU -------------------------
def _dist_worker(rank, world, port, tmp):
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    from param_b200.comms.pt.dlrm import (SparseBatch, indices_exchange_counts, lengths_exchange_splits,
                                          split_lengths)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T, b = 5, 6
    ts = split_lengths(T, world)
    rng = np.random.default_rng(100 + rank)
    offs, idxs = [], []
    for t in range(T):
        lens = rng.integers(0, 4, size=b)
        offs.append(torch.from_numpy(np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)))
        idxs.append(torch.from_numpy(rng.integers(0, 10**6, size=int(lens.sum())).astype(np.int64)))
    sb = SparseBatch.from_offsets(offs, idxs, device="cpu")
    in_s, out_s = lengths_exchange_splits(ts, rank, b)
    lengths_out = torch.empty(sum(out_s), dtype=torch.int64)
    dist.all_to_all_single(lengths_out, sb.lengths, out_s, in_s)
    send, recv = indices_exchange_counts(sb.lengths, lengths_out, ts, b)
    indices_out = torch.empty(int(recv.sum()), dtype=torch.int64)
    dist.all_to_all_single(indices_out, sb.indices, recv.tolist(), send.tolist())
    np.savez(os.path.join(tmp, f"r{rank}.npz"), lengths=sb.lengths.numpy(), indices=sb.indices.numpy(),
             lengths_out=lengths_out.numpy(), indices_out=indices_out.numpy(),
             **{f"off{t}": offs[t].numpy() for t in range(T)}, **{f"idx{t}": idxs[t].numpy() for t in range(T)})
    dist.barrier()
    dist.destroy_process_group()


def test_sparse_redistribution_plan_two_ranks_gloo(oracle, tmp_path):
    import torch.multiprocessing as mp
    from param_b200.comms.pt.dlrm import owner_slice, split_lengths
    W, T, b = 2, 5, 6
    mp.spawn(_dist_worker, args=(W, 29671, str(tmp_path)), nprocs=W, join=True)
    data = [np.load(tmp_path / f"r{r}.npz") for r in range(W)]
    ts = split_lengths(T, W)
    for r in range(W):
        Tl = ts[r]
        sl = owner_slice(r, ts)
        _, offsets, indices = oracle.split_per_table(data[r]["lengths_out"], data[r]["indices_out"], W, Tl, b)
        for f in range(Tl):
            g = sl.start + f
            lo, hi = offsets[f * W * b], offsets[(f + 1) * W * b]
            # table g over the global batch == concatenation over ranks of that rank's bags
            want_idx = np.concatenate([data[s][f"idx{g}"] for s in range(W)])
            assert np.array_equal(indices[lo:hi], want_idx)
            starts, base = [], 0
            for s in range(W):
                starts.append(data[s][f"off{g}"] + base)
                base += data[s][f"idx{g}"].size
            assert np.array_equal(offsets[f * W * b:(f + 1) * W * b] - lo, np.concatenate(starts))


# ---- et_replay hooks against the real reference package (CPU, skipped when it is absent) -------
def _et_replay_importable():
    if not REF.exists():
        return False
    import types
    sys.dont_write_bytecode = True
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    for name, attrs in (("pydot", ["Dot", "Node", "Edge", "Cluster"]), ("intervaltree", ["Interval", "IntervalTree"])):
        if name not in sys.modules:              # absent third-party deps of et_replay
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, type(a, (), {}))
            sys.modules[name] = m
    return True


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present on this box")
def test_et_replay_builds_b200_ops_from_name_and_schema():
    """et_replay re-creates compute nodes from node.name + node.op_schema
    (et_replay/et_replay_utils.py:129-212); the b200:: ops registered by `import param_b200.et`
    (what the replay config's "import modules" does) must build through that exact function."""
    assert _et_replay_importable()
    import json
    import types
    from et_replay import et_replay_utils
    cfg = json.loads((ROOT / "param_b200" / "et" / "replay-config-b200.json").read_text())
    assert cfg["import modules"] == ["param_b200.et"]
    for mod in cfg["import modules"]:
        __import__(mod)
    schemas = {
        "b200::embedding_bag": ("b200::embedding_bag(Tensor weight, Tensor indices, Tensor offsets, int mode, "
                                "Tensor? per_sample_weights, bool include_last_offset) -> Tensor", 6, 1),
        "b200::tbe_forward": ("b200::tbe_forward(Tensor weights, Tensor row_offsets, int dim, Tensor indices, "
                              "Tensor offsets, int batch, int mode, Tensor? per_sample_weights, int layout) -> Tensor", 9, 1),
        "b200::regroup_sparse": ("b200::regroup_sparse(Tensor lengths, Tensor indices, int world, int tables_local, "
                                 "int local_batch) -> (Tensor, Tensor, Tensor)", 5, 3),
        "b200::tbe_backward_": ("b200::tbe_backward_(Tensor(a!) dst, Tensor row_offsets, int dim, Tensor indices, "
                                "Tensor offsets, int batch, Tensor grad_out, int layout, float scale, int mode, "
                                "int algo) -> Tensor(a!)", 11, 1),
        "b200::tbe_backward_fused_": ("b200::tbe_backward_fused_(Tensor(a!) weights, Tensor(b!)? state, "
                                      "Tensor row_offsets, int dim, Tensor indices, Tensor offsets, int batch, "
                                      "Tensor grad_out, int layout, int mode, Tensor? per_sample_weights, "
                                      "int optimizer, float lr, float eps, bool stochastic_rounding, int sr_seed) "
                                      "-> Tensor(a!)", 16, 1),
    }
    for name, (schema, n_in, n_out) in schemas.items():
        # the schema string must be the one the dispatcher really holds (what an ET capture records)
        op = getattr(torch.ops.b200, name.split("::")[1])
        assert str(op.default._schema).replace(" ", "") == schema.replace(" ", ""), name
        node = types.SimpleNamespace(name=name, op_schema=schema, id=1,
                                     input_types=["x"] * n_in, output_types=["y"] * n_out)
        func, out_count = et_replay_utils.build_torchscript_func(node)
        assert func is not None and out_count == n_out, name


def _reference_dlrm_trace_nodes():
    """fbgemm:: nodes of rank 0 of the reference's own DLRM test trace (et_replay/tests/inputs/dlrm_pytorch_et.tar.gz)"""
    import io
    import json
    import tarfile
    tar = REF / "et_replay" / "tests" / "inputs" / "dlrm_pytorch_et.tar.gz"
    with tarfile.open(tar) as tf:
        member = [m for m in tf.getmembers() if m.name.endswith("dlrm_eg_0.json")][0]
        nodes = json.load(io.TextIOWrapper(tf.extractfile(member)))["nodes"]
    return [n for n in nodes if n["name"].startswith("fbgemm::")]


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present on this box")
def test_fbgemm_named_ops_carry_the_schemas_of_the_reference_dlrm_trace():
    """The reference's DLRM test trace records the fbgemm_gpu form of the path; the shims must be the ops et_replay
    would build from those nodes: same names, character-for-character the same schemas, buildable through
    et_replay_utils.build_torchscript_func (et_replay/et_replay_utils.py:129-212)."""
    assert _et_replay_importable()
    import json
    import types
    from et_replay import et_replay_utils
    cfg = json.loads((ROOT / "param_b200" / "et" / "replay-config-b200-fbgemm.json").read_text())
    for mod in cfg["import modules"]:
        __import__(mod)
    from param_b200.et import fbgemm_ops
    nodes = _reference_dlrm_trace_nodes()
    names = {n["name"] for n in nodes}
    assert names == {"fbgemm::" + k for k in fbgemm_ops.SCHEMAS}
    for n in nodes:
        short = n["name"].split("::")[1]
        assert n["op_schema"] == "fbgemm::" + fbgemm_ops.SCHEMAS[short], short
        held = str(getattr(torch.ops.fbgemm, short).default._schema)
        assert held.replace(" ", "") == n["op_schema"].replace(" ", ""), short
        n_out = len(n["outputs"]) if "outputs" in n else len(n.get("output_types", []))
        if n_out == 0:
            assert n["name"] in cfg["skip nodes"]            # () -> nothing to rebuild: on the skip list
            continue
        node = types.SimpleNamespace(name=n["name"], op_schema=n["op_schema"], id=n["id"],
                                     input_types=n.get("input_types", n.get("inputs", [])),
                                     output_types=["y"] * n_out)
        func, out_count = et_replay_utils.build_torchscript_func(node)
        assert func is not None and out_count == n_out, short


def test_fbgemm_bookkeeping_ops_and_layout_mapping(oracle):
    """asynchronous_complete_cumsum / permute_2D_sparse_data semantics (they are ATen integer glue and run on CPU
    tensors), and the mapping of fbgemm's flat weight buffer + per-feature offsets onto a table arena, checked by
    running the ORACLE lookup on the mapped arguments against a direct per-feature computation."""
    from param_b200._cabi import PB200Error
    from param_b200.et import fbgemm_ops as fb
    rng = np.random.default_rng(5)
    for dtype in (torch.int32, torch.int64):
        t = torch.from_numpy(rng.integers(0, 9, size=37)).to(dtype)
        out = torch.ops.fbgemm.asynchronous_complete_cumsum(t)
        assert out.dtype == dtype and out.tolist() == [0] + np.cumsum(t.numpy()).tolist()
    assert torch.ops.fbgemm.asynchronous_complete_cumsum(torch.zeros(0, dtype=torch.int64)).tolist() == [0]
    # permute: rows repeated, dropped and reordered; with and without weights / the length hint
    T, B = 5, 4
    lengths = torch.from_numpy(rng.integers(0, 4, size=(T, B))).to(torch.int32)
    n = int(lengths.sum())
    values = torch.arange(100, 100 + n, dtype=torch.int64)
    weights = torch.from_numpy(rng.standard_normal(n).astype(np.float32))
    seg = lengths.sum(1).tolist()
    starts = np.concatenate([[0], np.cumsum(seg)])
    for perm in ([4, 0, 2, 2, 1], [3], [0, 1, 2, 3, 4], [1, 1, 1]):
        want_v = np.concatenate([values.numpy()[starts[p]:starts[p + 1]] for p in perm] + [np.zeros(0, np.int64)])
        want_w = np.concatenate([weights.numpy()[starts[p]:starts[p + 1]] for p in perm] + [np.zeros(0, np.float32)])
        pl, pv, pw = torch.ops.fbgemm.permute_2D_sparse_data(torch.tensor(perm, dtype=torch.int32), lengths, values,
                                                             weights, None)
        assert torch.equal(pl, lengths[perm]) and np.array_equal(pv.numpy(), want_v) and np.array_equal(pw.numpy(), want_w)
        pl2, pv2, pw2 = torch.ops.fbgemm.permute_2D_sparse_data(torch.tensor(perm, dtype=torch.int32), lengths, values,
                                                                None, int(want_v.size))
        assert pw2 is None and torch.equal(pv2, pv) and torch.equal(pl2, pl)
    # layout: 3 features of dim 8 stored out of order in one flat buffer, with a gap
    D, rows = 8, [7, 5, 9]
    elem_off = [5 * D, 0, 16 * D]                      # feature 0 at row 5, feature 1 at row 0, feature 2 at row 16
    n_w = 27 * D
    flat = rng.standard_normal(n_w).astype(np.float32)
    Bq = 6
    lens = rng.integers(0, 5, size=len(rows) * Bq)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.concatenate([rng.integers(0, rows[f], size=int(lens[f * Bq:(f + 1) * Bq].sum())) for f in range(3)]).astype(np.int64)
    Tn, Bn, Dn, row_off, rows_out = fb.tbe_layout(elem_off, [0, 8, 16, 24], [0, 7, 12, 21], 24, 8, offsets.size, n_w)
    assert (Tn, Bn, Dn, rows_out) == (3, Bq, D, rows) and row_off == [5, 0, 16, 27]
    got = oracle.tbe_fwd(flat.reshape(-1, D), np.array(row_off, dtype=np.int64), D, idx, offsets, Bq)
    want = np.zeros((Bq, 3 * D), dtype=np.float32)
    for f in range(3):
        table = flat[elem_off[f]:elem_off[f] + rows[f] * D].reshape(rows[f], D)
        for b in range(Bq):
            lo, hi = offsets[f * Bq + b], offsets[f * Bq + b + 1]
            acc = np.zeros(D, dtype=np.float32)
            for i in idx[lo:hi]:
                acc = acc + table[i]
            want[b, f * D:(f + 1) * D] = acc
    assert np.array_equal(got, want)
    for bad in (lambda: fb.tbe_layout(elem_off, [0, 8, 16, 20], [0, 7, 12, 21], 20, 8, offsets.size, n_w),     # mixed dims
                lambda: fb.tbe_layout([3, 0, 128], [0, 8, 16, 24], [0, 7, 12, 21], 24, 8, offsets.size, n_w),   # off a row
                lambda: fb.tbe_layout(elem_off, [0, 8, 16, 24], [0, 7, 12, 40], 24, 8, offsets.size, n_w),     # past the end
                lambda: fb.tbe_layout(elem_off, [0, 8, 16, 24], [0, 7, 12, 21], 24, 8, offsets.size + 1, n_w)):
        with pytest.raises(PB200Error):
            bad()
    # a replayed trace without saved integral data: metadata tensors are constant-filled, scalars and sizes survive
    # (the reference's DLRM test trace: total_D 48, max_D 16, 385 offsets, 640 649 648 weights)
    Tn, Bn, Dn, row_off, rows_out = fb.tbe_layout_even(48, 16, 385, 640649648)
    assert (Tn, Bn, Dn) == (3, 128, 16) and rows_out == [13346867] * 3 and row_off == [0, 13346867, 26693734, 40040601]
    assert row_off[-1] * 16 <= 640649648
    with pytest.raises(PB200Error):
        fb.tbe_layout_even(48, 10, 385, 640649648)
    # the lookups themselves are CUDA only
    with pytest.raises(PB200Error):
        torch.ops.fbgemm.dense_embedding_codegen_lookup_function(
            torch.from_numpy(flat), torch.tensor(elem_off), torch.tensor([0, 8, 16, 24], dtype=torch.int32), 24, 8,
            torch.tensor([0, 7, 12, 21]), 5, torch.from_numpy(idx), torch.from_numpy(offsets), 0, None, None, 0)


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present on this box")
def test_et_replay_comm_backend_registers():
    assert _et_replay_importable()
    import inspect
    from et_replay.comm.backend import base_backend
    from param_b200.et.backend import B200ETStandalone, register_et_backend
    cls = register_et_backend("b200")
    assert base_backend.customized_backend["b200"] is cls and not inspect.isabstract(cls)
    missing = [n for n in base_backend.BaseBackend.__abstractmethods__ if not callable(getattr(B200ETStandalone, n, None))]
    assert not missing, missing
    from param_b200.comms.pt.backend import B200CommsMixin
    assert cls.all_to_allv is B200CommsMixin.all_to_allv     # replayed alltoall_base lands on the push kernel


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present on this box")
def test_generated_basic_trace_parses_with_the_reference_parser(tmp_path):
    """tools/make_basic_trace.py writes one DLRM step in the reference's basic trace format; the reference's own
    parser (train/comms/pt/commsTraceParser.py:64-152) must read it, incl. the "compute": "emb_lookup" entries
    (:137-147) that its replay hands to init_emb_lookup / the backend's emb_lookup."""
    import json
    import subprocess
    from param_b200.integration import refpath
    refpath.setup()
    from param_bench.train.comms.pt import commsTraceParser as parser
    W, T, b, L, E, rows = 4, 8, 64, 5, 32, 1000
    out = tmp_path / "step.json"
    subprocess.run([sys.executable, str(ROOT / "tools" / "make_basic_trace.py"), "--world", str(W), "--out", str(out),
                    "--tables", str(T), "--local-batch", str(b), "--bag", str(L), "--dim", str(E), "--rows", str(rows)],
                   check=True, capture_output=True)
    trace = parser.parseTrace(json.loads(out.read_text()), "basic", 0, W)
    kinds = [(c.comms, c.compute) for c in trace]
    assert kinds == [("all_to_all", None), ("wait", None), (None, "emb_lookup"), ("all_to_all", None), ("wait", None),
                     ("all_to_all", None), ("wait", None), (None, "emb_lookup")]
    idx, fwd, pooled, bwd = trace[0], trace[2], trace[3], trace[7]
    assert (idx.inMsgSize, idx.outMsgSize, idx.dtype) == (T * W * b * L, T * W * b * L, "long")     # ELEMENTS per rank
    assert (pooled.inMsgSize, pooled.dtype) == (T * b * W * E, "float") and trace[5].inMsgSize == pooled.inMsgSize
    for c, direction in ((fwd, "forward"), (bwd, "backward")):
        assert (c.direction, c.emb_dim, c.num_embs, c.batch_size, c.num_emb_tables_per_device, c.bag_size, c.count) == \
            (direction, E, rows, b * W, T, L, 1)
    assert [c.req for c in trace if c.comms] == [0, 0, 1, 1, 2, 2]


def test_tbe_request_generator_layout():
    """TBE request layout (split_table_batched_embeddings_ops.py:93-135,191-213): table-major
    indices, cumulative offsets over the concatenation, the reference's alpha regimes."""
    from param_b200.comms.pt.emb_lookup import generate_requests
    B, T, L, E = 6, 3, 4, 50
    (idx, off, w), = generate_requests(1, B, T, L, E, alpha=0.0)
    assert off.tolist() == [i * L for i in range(T * B + 1)] and idx.numel() == T * B * L and w is None
    assert idx[:B * L].tolist() == [i % L for i in range(B * L)]                 # alpha == 0: i % L
    (idx, off, w), = generate_requests(1, B, T, L, E, alpha=0.3, weighted=True)
    assert idx[:B * L].tolist() == [i % E for i in range(B * L)] and w.numel() == idx.numel()
    for alpha in (1.0, 1.15):
        (idx, _, _), = generate_requests(1, B, T, L, E, alpha=alpha, seed=1)
        assert int(idx.min()) >= 0 and int(idx.max()) < E


def test_operator_plugin_protocol():
    """OperatorInterface protocol of train/compute/python (lib/operator.py:8-45)"""
    from param_b200._cabi import PB200Error
    from param_b200.compute.operator import B200BatchedEmbeddingBagOp
    op = B200BatchedEmbeddingBagOp()
    for name in ("build", "cleanup", "forward", "create_grad", "backward"):
        assert callable(getattr(op, name))
    op.device = "cpu"
    with pytest.raises(PB200Error):
        op.build(2, 100, 16)
    if REF.exists():
        from make_golden import _ref_paths
        _ref_paths()
        from param_bench.train.compute.python.lib import operator as ref_operator
        assert issubclass(B200BatchedEmbeddingBagOp, ref_operator.OperatorInterface)   # __subclasshook__
        from param_b200.compute import operator as plugin
        name = "B200BatchedEmbeddingBag_test"
        registered = plugin.register(name)
        assert ref_operator.op_map[name] is registered
        with pytest.raises(ValueError):
            plugin.register(name)


def _payload_worker(rank, world, port, tmp):
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    from param_b200.comms.pt.comms import _expected, _payload
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for dtype in (torch.float32, torch.int64, torch.uint8):
        numel = 64 * world
        x = _payload(rank, numel, dtype, "cpu")
        out = torch.empty_like(x)
        dist.all_to_all_single(out, x)
        ok &= bool(torch.equal(out, _expected(rank, world, numel, dtype, "cpu")))
        # a deliberately wrong permutation (blocks reversed) must be caught by the position code
        wrong = torch.cat(list(reversed(out.split(numel // world))))
        ok &= not bool(torch.equal(wrong, _expected(rank, world, numel, dtype, "cpu")))
    Path(tmp, f"ok{rank}").write_text("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_position_coded_data_check_matches_c10d_two_ranks_gloo(tmp_path):
    """the sweep's `--c 1` oracle (_payload/_expected) against a real c10d all_to_all_single, and its
    power to detect a wrong permutation (the reference's constant-fill dcheck cannot)"""
    import torch.multiprocessing as mp
    mp.spawn(_payload_worker, args=(2, 29681, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").read_text() == "1" and (tmp_path / "ok1").read_text() == "1"


def test_aten_override_registers_cuda_kernels_only():
    """param_b200.et.aten_override (own process: the registration is process-wide) replaces the CUDA
    kernels of the three aten EmbeddingBag ops and leaves the CPU path alone."""
    import subprocess
    code = (
        "import sys, torch\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "import param_b200.et.aten_override\n"
        "for op in ('_embedding_bag', '_embedding_bag_forward_only', '_embedding_bag_backward'):\n"
        "    d = torch._C._dispatch_dump('aten::' + op)\n"
        "    cuda = [l for l in d.splitlines() if l.startswith('CUDA:')]\n"
        "    assert len(cuda) == 1 and 'aten_override.py' in cuda[0], (op, cuda)\n"
        "    cpu = [l for l in d.splitlines() if l.startswith('CPU:')]\n"
        "    assert len(cpu) == 1 and 'aten_override.py' not in cpu[0], (op, cpu)\n"
        "e = torch.nn.EmbeddingBag(10, 4, mode='sum')\n"
        "o = e(torch.tensor([1, 2, 3, 4]), torch.tensor([0, 2]))\n"
        "o.sum().backward()\n"
        "assert e.weight.grad[1].tolist() == [1.0] * 4\n"
        "print('OK')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
    import json
    cfg = json.loads((ROOT / "param_b200" / "et" / "replay-config-b200-aten.json").read_text())
    assert cfg["import modules"] == ["param_b200.et", "param_b200.et.aten_override"]


def test_fused_backward_and_fp16_paths_refuse_cpu_tensors_and_bad_arguments():
    """no CPU fallback anywhere: the new entries raise PB200Error before touching the library"""
    from param_b200 import ops
    from param_b200._cabi import PB200Error
    w = torch.zeros(10, 8)
    ro = torch.tensor([0, 10])
    idx, off, g = torch.zeros(4, dtype=torch.int64), torch.tensor([0, 2, 4]), torch.ones(2, 8)
    with pytest.raises(PB200Error, match="CUDA tensors only"):
        ops.tbe_backward_fused(w, ro, 1, 8, idx, off, 2, g)
    with pytest.raises(PB200Error, match="CUDA tensors only"):
        ops.tbe_forward(ops.TableArena(w.half(), ro, [10], 8), idx, off, 2)
    with pytest.raises(PB200Error):
        ops.TableArena.allocate([10], 8, "cpu", dtype=torch.bfloat16)        # fp32 / fp16 tables only
    assert ops._OPTIMIZER["exact_row_wise_adagrad"] == ops.OPT_ROWWISE_ADAGRAD == 2
    assert ops._OPTIMIZER["exact_sgd"] == ops.OPT_SGD == 1                     # values of include/param_b200.h
    from param_b200.compute.tbe import B200TBE
    with pytest.raises(PB200Error):
        B200TBE([(10, 8), (10, 16)], device="cpu")                            # mixed dims
    with pytest.raises(PB200Error):
        B200TBE([(10, 8)], device="cpu", weights_precision="int8")
    with pytest.raises(PB200Error):
        B200TBE([(10, 8)], device="cpu", optimizer="adam")
    with pytest.raises(PB200Error):
        B200TBE([(10, 8)], device="cpu", optimizer="exact_row_wise_adagrad", bwd_algo="sorted")


def test_comms_compute_runner_flags_match_the_reference_names():
    """commsComputeBench flag names and defaults (commsComputeBench.py:37-136)"""
    from param_b200.comms.pt.comms_compute import _args
    a = _args([])
    assert (a.mode, a.kernel, a.num_compute, a.emb_dim, a.num_embs, a.batch_size, a.num_emb_tables_per_device,
            a.num_emb_tables_batched, a.bag_size) == ("comms-compute", "emb_lookup", 100, 128, 100000, 512, 8, -1, 20)
    a = _args(["--mode", "compute", "--ntables", "4", "--num-compute-per-iteration", "7", "--begin-size", "2M"])
    assert a.mode == "compute" and a.num_emb_tables_per_device == 4 and a.num_compute == 7 and a.b == "2M"
    with pytest.raises(SystemExit):
        _args(["--kernel", "gemm"])            # dense kernels are outside the hot path


def test_compute_python_plugin_registers_iterator_generator_and_clear_cache():
    """f2: operator + input iterator + input-data generator in the reference's registries, same JSON schema as
    the reference's split_table_batched_embeddings_ops.json; _clear_cache gets the sm_100 entry."""
    import json
    from param_b200.integration import refpath
    if refpath.setup(require=False) is None:
        pytest.skip("reference tree not present")
    from param_b200.compute import python_plugin as pp
    pp.register_all()
    pp.register_all()                                   # idempotent
    from param_bench.train.compute.python.lib import data as d, iterator as it, operator as o
    from param_bench.train.compute.python.lib.pytorch import op_executor
    assert pp.OP_NAME in o.op_map and pp.ITERATOR_NAME in it.config_iterator_map
    assert pp.GENERATOR_NAME in d.data_generator_map and op_executor._clear_cache._pb200
    cfg = json.loads((ROOT / "param_b200" / "compute" / "configs" / "b200_batched_embedding_bag.json").read_text())
    spec = cfg[pp.OP_NAME]
    assert spec["input_iterator"] == pp.ITERATOR_NAME and spec["input_data_generator"] == pp.GENERATOR_NAME
    c0 = spec["config"][0]
    build = {"args": c0["build"][0]["args"], "kwargs": c0["build"][0]["kwargs"]}
    runs = list(it.config_iterator_map[pp.ITERATOR_NAME]({"build": build, "input": c0["input"]}, "input", "cpu"))
    assert [a["value"] for a in runs[0][1]["args"]] == [16, 1000000, 128, 65536, 20, False, "fp32"]
    ic = runs[0][1]
    ic["args"][0]["value"], ic["args"][1]["value"], ic["args"][3]["value"], ic["args"][4]["value"] = 3, 50, 6, 4
    (idx, off, w), kw = d.data_generator_map[pp.GENERATOR_NAME]().get_data(ic, "cpu", alpha=1.0)
    assert idx.numel() == 3 * 6 * 4 and off.tolist() == list(range(0, 73, 4)) and w is None and kw == {}
    assert int(idx.max()) < 50 and int(idx.min()) >= 0
    # the reference's alpha convention: 0 -> arange % L, <= 0.5 -> arange % E, > 1 -> Zipf % E
    i0, o0, _ = pp.generate_requests(4, 3, 10, 0, alpha=0.0)
    assert i0.tolist() == [0, 1, 2] * 4 and o0.tolist() == [0, 3, 6, 9, 12]
    i1, o1, w1 = pp.generate_requests(4, 3, 10, 12, alpha=1.2, weighted=True)
    assert o1.tolist() == [15, 18, 21, 24] and int(i1.max()) < 10 and w1.numel() == 12
