"""Run in its own process (the override is process-wide): stock nn.EmbeddingBag on CUDA with
param_b200.et.aten_override imported must produce the torch-CPU golden outputs and gradients, and
must do it on the libparam_b200 kernels (launch counter moves)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import param_b200.et.aten_override  # noqa: E402,F401  (what the replay config's "import modules" does)
from param_b200 import _cabi  # noqa: E402

dev = torch.device("cuda:0")
d = np.load(ROOT / "tests" / "golden" / "embbag_torch_cpu.npz")
n_checked = 0
for k in range(int(d["n_cases"])):
    p = f"c{k}_"
    weight, indices, offsets = d[p + "weight"], d[p + "indices"], d[p + "offsets"]
    psw, mode = d[p + "psw"], str(d[p + "mode"])
    for sparse in (False, True):
        emb = torch.nn.EmbeddingBag(weight.shape[0], weight.shape[1], mode=mode, sparse=sparse,
                                    _weight=torch.from_numpy(weight.copy())).to(dev)
        n0 = _cabi.launch_count()
        out = emb(torch.from_numpy(indices).to(dev), torch.from_numpy(offsets).to(dev),
                  per_sample_weights=torch.from_numpy(psw).to(dev) if psw.size else None)
        assert _cabi.launch_count() > n0, "forward did not reach libparam_b200"
        if mode == "sum" and not psw.size:
            assert np.array_equal(out.detach().cpu().numpy(), d[p + "out"]), f"case {k} forward not bit-exact"
        else:
            np.testing.assert_allclose(out.detach().cpu().numpy(), d[p + "out"], rtol=1e-5, atol=1e-6)
        n1 = _cabi.launch_count()
        out.backward(torch.from_numpy(d[p + "grad_out"]).to(dev))
        assert _cabi.launch_count() > n1, "backward did not reach libparam_b200"
        g = emb.weight.grad
        g = g.to_dense() if g.is_sparse else g
        np.testing.assert_allclose(g.cpu().numpy(), d[p + "grad_weight"], rtol=1e-5, atol=1e-5,
                                   err_msg=f"case {k} sparse={sparse}")
        n_checked += 1

# the way et_replay re-creates the node: name + schema -> TorchScript IR (et_replay_utils.py:171-212)
ir = """
graph(%0: Tensor, %1: Tensor, %2: Tensor, %3: bool, %4: int, %5: bool, %6: Tensor?, %7: bool):
    %8: Tensor, %9: Tensor, %10: Tensor, %11: Tensor = aten::embedding_bag(%0, %1, %2, %3, %4, %5, %6, %7)
    %output : (Tensor, Tensor, Tensor, Tensor) = prim::TupleConstruct(%8, %9, %10, %11)
    return (%output)
"""
fn = torch._C.CompilationUnit().create_function("aten::embedding_bag", torch._C.parse_ir(ir))
w = torch.randn(50, 64, device=dev)
idx = torch.randint(0, 50, (40,), device=dev)
off = torch.arange(0, 40, 5, device=dev)
n0 = _cabi.launch_count()
got = fn(w, idx, off, False, 0, False, None, False)[0]
assert _cabi.launch_count() > n0
want = torch.stack([w[idx[i:i + 5]].cpu().double().sum(0) for i in range(0, 40, 5)])
assert torch.allclose(got.cpu().double(), want, rtol=1e-5, atol=1e-5)
try:
    torch.nn.EmbeddingBag(10, 8, mode="max").to(dev)(idx % 10, off)
    raise SystemExit("mode=max should have been refused")
except RuntimeError as exc:           # PB200Error is a RuntimeError
    assert "not covered" in str(exc)
print(f"ATEN-OVERRIDE-OK {n_checked}")
