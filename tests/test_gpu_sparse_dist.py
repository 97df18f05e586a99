"""Device-side sparse-input regroup vs the reference's splitPerTable (golden) and the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_regroup_matches_reference_golden(cuda_device, golden_dir):
    from param_b200 import ops
    d = np.load(golden_dir / "dlrm_sparse_ref.npz")
    for name in "abc":
        W, Tl, b = (int(x) for x in d[f"sp_{name}_dims"])
        lens = torch.from_numpy(d[f"sp_{name}_lengths"]).to(cuda_device)
        ind = torch.from_numpy(d[f"sp_{name}_indices"]).to(cuda_device)
        _, offsets_out, indices_out = ops.regroup_sparse(lens, ind, W, Tl, b)
        offsets_out, indices_out = offsets_out.cpu().numpy(), indices_out.cpu().numpy()
        for f in range(Tl):
            lo, hi = offsets_out[f * W * b], offsets_out[(f + 1) * W * b]
            assert np.array_equal(offsets_out[f * W * b:(f + 1) * W * b] - lo, d[f"sp_{name}_off{f}"])
            assert np.array_equal(indices_out[lo:hi], d[f"sp_{name}_idx{f}"])


@pytest.mark.parametrize("shape", [(8, 64, 128), (2, 3, 1), (4, 16, 1000), (1, 5, 7)])
def test_regroup_vs_oracle(cuda_device, oracle, shape):
    from param_b200 import ops
    W, Tl, b = shape
    rng = np.random.default_rng(W * 100 + Tl)
    lens = rng.integers(0, 30, size=W * Tl * b).astype(np.int64)
    ind = rng.integers(0, 1 << 40, size=int(lens.sum())).astype(np.int64)
    lo, oo, io = oracle.split_per_table(lens, ind, W, Tl, b)
    l2, o2, i2 = ops.regroup_sparse(torch.from_numpy(lens).to(cuda_device),
                                    torch.from_numpy(ind).to(cuda_device), W, Tl, b)
    assert np.array_equal(l2.cpu().numpy(), lo)
    assert np.array_equal(o2.cpu().numpy(), oo)
    assert np.array_equal(i2.cpu().numpy(), io)
