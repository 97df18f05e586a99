"""B200-native mirror of the reference op driver train/compute/pt/pytorch_emb.py.

Same surface (init_indices, measure_gpu, run_single, run, main; CLI flags --features --embdim --nnz
--batch --steps --warmups --randomseed --alpha -d/--device), but the module that is timed is
`B200EmbeddingBag`, whose forward/backward are the sm_100a kernels behind include/param_b200.h.
`B200EmbeddingBag` follows the call contract the reference relies on (pytorch_emb.py:179, :48-69,
and the in-tree precedent for swapping the module, XlaEmbeddingBag :14-34):
constructible as Cls(features, embdim, mode="sum"), exposes .weight, supports .to(device), and
__call__(indices int64[B*L], offsets int64[B]) -> fp32 [B, D] asynchronously on the current stream.

There is no CPU path: `--device cpu` is refused here (run the reference for that), and a CPU tensor
reaching the module raises.
"""
from __future__ import annotations

import sys
import time
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from ... import ops
from ..._cabi import PB200Error


# ---------------------------------------------------------------------------------------------
# module
# ---------------------------------------------------------------------------------------------
_SPARSE_GRAD_MIN_ROWS_PER_LOOKUP = 8      # sparse=True returns a COO gradient when rows > 8 x lookups


class _EmbeddingBagFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, indices, offsets, psw, mode, include_last_offset, fwd_algo, bwd_algo, sparse):
        out = ops.embedding_bag_forward(weight, indices, offsets, mode=mode,
                                        per_sample_weights=psw,
                                        include_last_offset=include_last_offset, algo=fwd_algo)
        ctx.save_for_backward(indices, offsets, psw if psw is not None else torch.empty(0))
        ctx.meta = (tuple(weight.shape), mode, include_last_offset, psw is not None, bwd_algo, sparse)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        indices, offsets, psw = ctx.saved_tensors
        (rows, dim), mode, include_last, weighted, bwd_algo, sparse = ctx.meta
        dev = grad_out.device
        indices = indices.contiguous().view(-1)
        offsets = offsets.contiguous().view(-1)
        if sparse and rows > _SPARSE_GRAD_MIN_ROWS_PER_LOOKUP * indices.numel():
            # nn.EmbeddingBag(sparse=True): an uncoalesced COO gradient with one entry per lookup — what the
            # reference's tables produce (pytorch_dist_backend.py:923-934) — instead of a dense zero-filled
            # [rows, dim] buffer (5.1 GB for a 10 M x 128 table).  When the table is NOT much larger than the
            # batch's lookups the dense form is the cheaper one (autograd accumulates it in place, while COO
            # gradients are concatenated step after step: 107 ms vs 57 ms per iteration under the reference's
            # dlrm.py with 200 k-row tables, profiles/r02g_dlrm_*.log), so it is used there.
            g = ops.embedding_bag_backward_sparse(grad_out.contiguous(), indices, offsets, rows, mode=mode,
                                                  per_sample_weights=psw if weighted else None,
                                                  include_last_offset=include_last)
            return g, None, None, None, None, None, None, None, None
        if not include_last:  # the batched kernel wants B+1 offsets
            offsets = ops.close_offsets(offsets, indices.numel())
        n_bags = offsets.numel() - 1
        grad_w = torch.zeros((rows, dim), dtype=torch.float32, device=dev)
        row_off = ops.single_table_row_offsets(rows, dev)
        ops.tbe_backward(grad_w, row_off, 1, dim, indices, offsets, n_bags, grad_out.contiguous(),
                         layout="TBD", scale=1.0, mode=mode,
                         per_sample_weights=psw if weighted else None, algo=bwd_algo)
        return grad_w, None, None, None, None, None, None, None, None


class B200EmbeddingBag(nn.Module):
    """Drop-in for torch.nn.EmbeddingBag on the PARAM hot path (sum / mean pooling, fp32)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, mode: str = "sum",
                 sparse: bool = False, include_last_offset: bool = False,
                 _weight: Optional[torch.Tensor] = None, device=None,
                 fwd_algo: str = "auto", bwd_algo: str = "auto") -> None:
        super().__init__()
        if mode not in ("sum", "mean"):
            raise PB200Error(f"mode {mode!r} is not on the PARAM hot path (sum/mean only)")
        self.num_embeddings, self.embedding_dim = int(num_embeddings), int(embedding_dim)
        self.mode, self.sparse, self.include_last_offset = mode, sparse, include_last_offset
        self.fwd_algo, self.bwd_algo = fwd_algo, bwd_algo
        if _weight is None:
            w = torch.empty((self.num_embeddings, self.embedding_dim), dtype=torch.float32, device=device)
            nn.init.normal_(w)  # nn.EmbeddingBag's default initialisation
        else:
            w = _weight
        self.weight = nn.Parameter(w)

    def forward(self, indices: torch.Tensor, offsets: Optional[torch.Tensor] = None,
                per_sample_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
        if indices.dim() == 2:
            if offsets is not None:
                raise PB200Error("offsets must be None for 2-D indices")
            b, l = indices.shape
            offsets = torch.arange(0, b * l, l, dtype=indices.dtype, device=indices.device)
            indices = indices.reshape(-1)
            ilo = False
        else:
            if offsets is None:
                raise PB200Error("offsets required for 1-D indices")
            ilo = self.include_last_offset
        return _EmbeddingBagFn.apply(self.weight, indices, offsets, per_sample_weights, self.mode,
                                     ilo, self.fwd_algo, self.bwd_algo, self.sparse)

    def extra_repr(self) -> str:
        return f"{self.num_embeddings}, {self.embedding_dim}, mode={self.mode!r}"


# ---------------------------------------------------------------------------------------------
# synthetic indices (reference: init_indices, pytorch_emb.py:138-160)
# ---------------------------------------------------------------------------------------------
def zipf_cdf(alpha: float, features: int) -> np.ndarray:
    """Normalised inclusive CDF of the truncated Zipf pmf k^-alpha, k = 1..features (float64)."""
    pmf = np.arange(1, features + 1, dtype=np.float64) ** (-float(alpha))
    cdf = np.cumsum(pmf)
    cdf /= cdf[-1]
    return cdf


def _first_distinct(row, nnz):
    """First `nnz` distinct values of `row`, listed in CPython set order — what the reference's
    per-bag loop leaves in `list(r)` (pytorch_emb.py:147-157)."""
    seen = set()
    for value in row:
        seen.add(value)
        if len(seen) == nnz:
            break
    return list(seen)


def init_indices(alpha, features, batch, nnz, compat: bool = True, seed: Optional[int] = None,
                 device=None) -> torch.Tensor:
    """int64 [batch*nnz] lookup indices, uniform (alpha == 0) or truncated Zipf(alpha).

    compat=True  reproduces the reference generator call for call (torch.randint under the torch
                 seed; np.random.choice over the global numpy RNG + per-bag first-nnz-distinct in
                 set order) so that seeded runs are bit-identical to the reference's.  Like the
                 reference it fails when a bag has fewer than nnz distinct draws among 2*nnz.
    compat=False the scalable generator: bag-wise inverse-CDF draws on the device with the same
                 "distinct inside a bag" rule, counter-based (value = f(seed, bag, draw)), never
                 fails, no O(batch*nnz) python loop.  Distribution-equivalent, not bit-equal.
    """
    alpha = float(alpha)  # the reference driver passes --alpha through as a str (driver.py:44-46)
    if compat:
        if alpha == 0.0:
            return torch.randint(0, features, (batch * nnz,))
        weights = np.power(np.arange(1, features + 1, dtype=np.float64), -alpha)
        draws = np.random.choice(features, size=(batch, 2 * nnz), replace=True, p=weights / weights.sum())
        picked = np.empty((batch, nnz), dtype=draws.dtype)
        for b in range(batch):
            keep = _first_distinct(draws[b], nnz)
            if len(keep) != nnz:
                raise ValueError(
                    f"bag {b}: only {len(keep)} distinct indices among {2 * nnz} Zipf draws "
                    "(the reference generator fails here too; use compat=False)")
            picked[b] = keep
        return torch.from_numpy(picked.reshape(-1)).to(torch.int64)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    seed = 0 if seed is None else int(seed)
    out = torch.empty(batch * nnz, dtype=torch.int64, device=device)
    if alpha == 0.0:
        cdf = torch.linspace(1.0 / features, 1.0, features, dtype=torch.float64, device=device)
        return ops.fill_zipf_indices_(out, nnz, cdf, seed, dedupe=False)
    cdf = torch.from_numpy(zipf_cdf(alpha, features)).to(device)
    return ops.fill_zipf_indices_(out, nnz, cdf, seed, dedupe=True)


def make_offsets(batch: int, nnz: int) -> torch.Tensor:
    """offsets[i] = i*nnz, no trailing entry (pytorch_emb.py:171-174)."""
    return torch.arange(batch, dtype=torch.int64) * nnz


# ---------------------------------------------------------------------------------------------
# timing loops (reference: measure_gpu, pytorch_emb.py:48-69)
# ---------------------------------------------------------------------------------------------
def measure_gpu(warmups, steps, h_emb, h_indices, h_offsets):
    """Same loop shape as the reference: module + inputs moved to cuda:0 once, warmups + steps
    asynchronous launches, one synchronize at the end; returns (wall seconds of `steps`, results)."""
    dev = torch.device("cuda:0")
    with torch.cuda.device(dev):
        g_emb = h_emb.to(dev)
        g_indices = h_indices.to(dev)
        g_offsets = h_offsets.to(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        results = None
        for i in range(warmups + steps):
            results = g_emb(g_indices, g_offsets)
            if i < warmups:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
    return t1 - t0, results


def run_single(args, features, embdim, nnz, batch):
    if args.device != "gpu":
        raise PB200Error("param_b200 has no CPU path: use --device gpu (the reference driver covers cpu)")
    if not torch.cuda.is_available():
        print("CUDA is not available, could not run on GPU")
        sys.exit(1)
    torch.manual_seed(args.randomseed)
    np.random.seed(args.randomseed)  # the reference leaves numpy unseeded
    compat = not getattr(args, "fast_indices", False)
    h_indices = init_indices(args.alpha, features, batch, nnz, compat=compat, seed=args.randomseed)
    h_offsets = make_offsets(batch, nnz)
    with torch.no_grad():
        h_emb = B200EmbeddingBag(features, embdim, mode="sum", device="cuda:0")
        total_bytes = batch * nnz * embdim * h_emb.weight.element_size()
        emb_times, _ = measure_gpu(args.warmups, args.steps, h_emb, h_indices, h_offsets)
    return emb_times, total_bytes


_HEADER = "    Features    embdim    nnz     batch      Time(s)/step   Data(MB)   BW(GB/s)"


def run(args, dataset):
    rule = "-" * 81
    print(rule)
    print(_HEADER)
    print(rule)
    for features, embdim, nnz, batch in dataset:
        elap, total_bytes = run_single(args, features, embdim, nnz, batch)
        per_step = elap / args.steps
        mb = total_bytes / 1.0e6
        print("{:10},  {:6},  {:6},  {:8},    {:10.6f}, {:10.1f},  {:8.3f}".format(
            features, embdim, nnz, batch, per_step, mb, mb / per_step / 1.0e3))


def main(argv=None) -> None:
    import argparse

    ap = argparse.ArgumentParser(description="Measure the performance of the B200 EmbeddingBag")
    ap.add_argument("--features", type=int, default=1024)
    ap.add_argument("--embdim", type=int, default=64)
    ap.add_argument("--nnz", type=int, default=10)
    ap.add_argument("--batch", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmups", type=int, default=1)
    ap.add_argument("--randomseed", type=int, default=0)
    ap.add_argument("-t", "--dtype", type=str, default="float32")
    ap.add_argument("-d", "--device", choices=["cpu", "gpu", "tpu"], type=str, default="gpu")
    ap.add_argument("--usexlabag", action="store_true")
    ap.add_argument("--alpha", type=float, default=0.0, help="Zipf param. Use uniform if == 0.0")
    ap.add_argument("--fast-indices", action="store_true",
                    help="device-side generator instead of the reference-compatible one")
    args = ap.parse_args(argv)
    run(args, [(args.features, args.embdim, args.nnz, args.batch)])


if __name__ == "__main__":
    main()
