"""The stand-alone runners (mirrors of the reference CLIs) end to end on ONE GPU, each in its own
process with WORLD_SIZE=1: the compute driver (config 1 plumbing on the GPU), the all-to-all sweep
through B200Backend with the position-coded data check, and the DLRM pattern runner."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _run(args, port):
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1",
               MASTER_PORT=str(port), PYTHONPATH=str(ROOT))
    r = subprocess.run([sys.executable, *args], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return r.stdout


def test_compute_driver_config1(cuda_device):
    """config 1: single-table EmbeddingBag 1M x 64, batch 512, bag 20 through the op driver"""
    out = _run(["-m", "param_b200.compute.pt.pytorch_emb", "--features", "1000000", "--embdim", "64",
                "--nnz", "20", "--batch", "512", "--steps", "20", "--warmups", "3", "-d", "gpu"], 29701)
    row = [ln for ln in out.splitlines() if ln.strip().startswith("1000000")]
    assert len(row) == 1 and float(row[0].split(",")[-1]) > 0.0          # BW(GB/s) column
    out = _run(["-m", "param_b200.compute.pt.driver", "--steps", "5", "--warmups", "2", "--device", "gpu",
                "emb", "--dataset", "cfg1", "--alpha", "1.15", "--fast-indices"], 29702)
    assert "with emb dataset  cfg1" in out and "1000000" in out


@pytest.mark.parametrize("collective", ["all_to_all_single", "all_to_allv", "all_to_all"])
def test_comms_sweep_world1(cuda_device, collective):
    out = _run(["-m", "param_b200.comms.pt.comms", "--collective", collective, "--begin-size", "1K",
                "--end-size", "4M", "--step-factor", "8", "--num-iters", "3", "--num_warmup_iters", "1",
                "--check-data", "1", "--backend", "b200", "--json"], 29703)
    recs = [json.loads(ln) for ln in out.splitlines() if ln.startswith("{")]
    assert len(recs) == 5 and all(r["check"] == "PASS" and r["collective"] == collective for r in recs)


def test_comms_sweep_cuda_graph_world1(cuda_device):
    out = _run(["-m", "param_b200.comms.pt.comms", "--collective", "all_to_all_single", "--begin-size", "64K",
                "--end-size", "64K", "--num-iters", "4", "--num_warmup_iters", "1", "--check-data", "1",
                "--graph-launches", "2", "--backend", "b200", "--json"], 29704)
    recs = [json.loads(ln) for ln in out.splitlines() if ln.startswith("{")]
    assert len(recs) == 1 and recs[0]["check"] == "PASS"


def test_dlrm_pattern_runner_world1(cuda_device):
    out = _run(["-m", "param_b200.comms.pt.dlrm", "--mini-batch-size", "256", "--num-batches", "3",
                "--warmup-batches", "1", "--arch-embedding-size", "5000x6", "--arch-sparse-feature-size", "64",
                "--num-indices-per-lookup", "8", "--num-indices-per-lookup-fixed", "false", "--lr", "0.01",
                "--json"], 29705)
    rec = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
    assert rec["iter_time_ms_p50_max_rank"] > 0 and "offset_idx_xchg_ms_p50_max_rank" in rec
    # the reference's 21 report rows (dlrm.py:1015-1037) incl. the two MLP all-reduce stages
    reg = rec["regions_us_p50"]
    assert len(reg) == 21 and reg["iter_time"] > 0 and reg["bwd_a2a"] > 0 and reg["bwd_top_ar"] > 0 and reg["bwd_bot_ar"] > 0
    out = _run(["-m", "param_b200.comms.pt.dlrm", "--mini-batch-size", "128", "--num-batches", "2",
                "--warmup-batches", "1", "--arch-embedding-size", "5000x4", "--arch-sparse-feature-size", "32",
                "--num-indices-per-lookup", "4", "--arch-mlp-bot", "13-64-32", "--arch-mlp-top", "64-1",
                "--two-collective-dist", "--unfused-forward", "--perf-debug"], 29715)
    rows = [ln.split() for ln in out.splitlines() if ln.startswith("\t2\t")]
    assert len(rows) == 2 * 22 and rows[3][1] == "offset_xchg" and rows[21][1] == "total_time"
    assert float(rows[5][4]) > 0 and float(rows[7][4]) > 0        # idx_xchg and the lookup gap are real rows here


@pytest.mark.parametrize("direction", ["forward", "backward"])
def test_comms_compute_overlap_runner_world1(cuda_device, direction):
    """commsComputeBench --kernel emb_lookup (commsComputeBench.py:303-312) without fbgemm_gpu"""
    out = _run(["-m", "param_b200.comms.pt.comms_compute", "--mode", "comms-compute", "--kernel", "emb_lookup",
                "--collective", "all_to_all_single", "--b", "1M", "--e", "4M", "--f", "4", "--n", "3", "--w", "1",
                "--num-compute", "2", "--emb-dim", "64", "--num-embs", "20000", "--batch-size", "256",
                "--ntables", "4", "--num-emb-tables-batched", "2", "--bag-size", "8", "--direction", direction,
                "--json"], 29706)
    recs = [json.loads(ln) for ln in out.splitlines() if ln.startswith("{")]
    assert len(recs) == 2
    for r in recs:
        assert r["iter_us"] > 0 and r["comm_dev_us"] > 0 and r["compute_dev_us"] > 0 and r["lookups_per_s"] > 0
    out = _run(["-m", "param_b200.comms.pt.comms_compute", "--mode", "compute", "--n", "2", "--w", "1",
                "--num-compute", "3", "--emb-dim", "128", "--num-embs", "5000", "--batch-size", "64",
                "--ntables", "2", "--bag-size", "4", "--direction", direction, "--json"], 29707)
    rec = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
    assert rec["size_bytes"] == 0 and rec["compute_dev_us"] > 0


def test_aten_embedding_bag_override_in_its_own_process(cuda_device):
    """param_b200.et.aten_override: stock nn.EmbeddingBag (what a captured DLRM trace replays as
    aten::embedding_bag / aten::_embedding_bag_backward) lands on the B200 kernels and
    reproduces the torch-CPU goldens.  Own process: the registration is process-wide."""
    out = _run([str(ROOT / "tests" / "helpers" / "aten_override_check.py")], 29708)
    assert "ATEN-OVERRIDE-OK" in out
