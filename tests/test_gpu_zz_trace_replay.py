"""cfg5 in miniature (BASELINE.json configs[4]): capture an ExecutionTrace of one DLRM step built from stock PyTorch
modules on the GPU, then let the reference's own et_replay (et_replay/tools/et_replay.py:1125-1264, unmodified, from
baseline/_ref) replay it with the aten override in place — the trace's `aten::_embedding_bag*` nodes run on
libparam_b200.  Named to run after the kernel tests; skipped where the reference tree did not travel."""
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _env(port):
    return dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1",
                MASTER_PORT=str(port), PYTHONPATH=str(ROOT))


def test_captured_dlrm_trace_replays_on_the_b200_kernels(cuda_device, tmp_path):
    from param_b200.integration import refpath
    if refpath.find_reference() is None:
        pytest.skip("reference tree not present (baseline/_ref, tools/make_baseline_ref.sh)")
    trace_dir = tmp_path / "trace"
    r = subprocess.run([sys.executable, "tools/cfg5_capture.py", "--out", str(trace_dir), "--tables-per-rank", "3",
                        "--rows", "20000", "--dim", "64", "--local-batch", "256", "--bag", "8"],
                       cwd=ROOT, env=_env(29731), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    trace = trace_dir / "rank-0.json"
    assert trace.exists() and any((trace_dir / "rank-0_resources").iterdir())
    results = {}
    for name, cfg in (("b200", "replay-config-b200-aten.json"), ("stock", "replay-config-stock.json")):
        r = subprocess.run([sys.executable, "-m", "param_b200.integration.param_plugin", "et_replay", "--input",
                            str(trace), "-m", "comp", "--warmup-iter", "1", "--iter", "2", "--replay-config",
                            str(ROOT / "param_b200" / "et" / cfg)],
                           cwd=ROOT, env=_env(29732), capture_output=True, text=True, timeout=300)
        text = r.stdout + r.stderr
        assert r.returncode == 0 and "Replay finished successfully" in text, text[-4000:]
        cov = re.search(r"Operator coverage: = ([0-9.]+)", text)
        launched = re.search(r"libparam_b200 kernels launched by this process: (\d+)", text)
        results[name] = (float(cov.group(1)) if cov else 0.0, int(launched.group(1)) if launched else 0)
    # same operator coverage either way; the override run must actually have gone through the library
    assert results["b200"][0] == results["stock"][0] and results["b200"][0] > 0.9
    assert results["b200"][1] > 0 and results["stock"][1] == 0


@pytest.mark.skipif(os.environ.get("PB200_RUN_UNVERIFIED") != "1",
                    reason="written after the round's GPU budget was spent: not yet run on hardware "
                           "(set PB200_RUN_UNVERIFIED=1 to run it)")
def test_fbgemm_named_lookups_and_bounds_check_vs_oracle(cuda_device, oracle):
    """param_b200.et.fbgemm_ops on the GPU: both lookup functions against the oracle on a flat weight buffer with
    out-of-order tables, and bounds_check_indices in its three modes."""
    import numpy as np
    import torch
    from param_b200._cabi import PB200Error
    from param_b200.et import fbgemm_ops  # noqa: F401
    rng = np.random.default_rng(9)
    D, rows, B = 16, [40, 25, 60], 33
    elem_off = [25 * D, 0, 70 * D]
    n_w = 130 * D
    flat = rng.standard_normal(n_w).astype(np.float32)
    lens = rng.integers(0, 6, size=3 * B)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.concatenate([rng.integers(0, rows[f], size=int(lens[f * B:(f + 1) * B].sum())) for f in range(3)]).astype(np.int64)
    dev = cuda_device
    t = lambda a, dt=None: torch.from_numpy(np.asarray(a)).to(device=dev, dtype=dt)  # noqa: E731
    w, wo, do = t(flat), t(elem_off, torch.int64), t([0, D, 2 * D, 3 * D], torch.int32)
    hs, ix, off = t([0, 40, 65, 125], torch.int64), t(idx), t(offsets)
    want = oracle.tbe_fwd(flat.reshape(-1, D), np.array([25, 0, 70, 130], dtype=np.int64), D, idx, offsets, B)
    empty_f = torch.empty(0, device=dev)
    empty_i = torch.empty(0, dtype=torch.int32, device=dev)
    got = torch.ops.fbgemm.dense_embedding_codegen_lookup_function(w, wo, do, 3 * D, D, hs, 7, ix, off, 0, None, None, 0)
    assert np.array_equal(got.cpu().numpy(), want)
    got = torch.ops.fbgemm.split_embedding_codegen_lookup_adagrad_function(
        empty_f, w, empty_f, torch.empty((0, 0), device=dev), empty_i, wo, do, 3 * D, D, hs, 7, ix, off, 0, None, None,
        empty_i, False, 1.0, True, torch.zeros_like(w), empty_f, empty_i, wo, 1e-8, 0.01, 0)
    assert np.array_equal(got.cpu().numpy(), want)
    rpt = t(rows, torch.int64)
    warning = torch.zeros(1, dtype=torch.int64, device=dev)
    torch.ops.fbgemm.bounds_check_indices(rpt, ix, off, 0, warning, None)          # all in range: nothing happens
    assert int(warning.item()) == 0
    bad = ix.clone()
    first_of_table_1 = int(offsets[B])
    bad[first_of_table_1] = rows[1]                                                  # one past the end of table 1
    bad[0] = -3
    with pytest.raises(PB200Error):
        torch.ops.fbgemm.bounds_check_indices(rpt, bad.clone(), off, 0, warning, None)
    fixed = bad.clone()
    torch.ops.fbgemm.bounds_check_indices(rpt, fixed, off, 1, warning, None)
    assert int(warning.item()) == 2 and int(fixed[0]) == 0 and int(fixed[first_of_table_1]) == 0
    untouched = torch.ones_like(ix, dtype=torch.bool)
    untouched[0] = untouched[first_of_table_1] = False
    assert torch.equal(fixed[untouched], ix[untouched])
