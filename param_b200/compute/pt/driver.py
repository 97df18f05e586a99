"""`driver.py --device gpu emb --dataset A|B` on the B200 kernels.

Mirror of the reference CLI train/compute/pt/driver.py:13-109 for the ONE kernel on the hot path
(`emb`); gemm / linear are dense tensor-core work and out of scope, so asking for
them here is an error that points back at the reference.  `--alpha` is parsed as float (the
reference forgets type=, driver.py:44-46, and then fails inside np.power).
"""
from __future__ import annotations

import argparse

from . import dataset
from . import pytorch_emb as kemb


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description="Measuring the EmbeddingBag kernel on B200")
    ap.add_argument("--warmups", type=int, default=10, help="warmup times")
    ap.add_argument("--steps", type=int, default=100, help="repeat times")
    ap.add_argument("--device", type=str, choices=["cpu", "gpu", "tpu"], required=True)
    sub = ap.add_subparsers(title="kernels", dest="kernel")
    sub.required = True
    p_emb = sub.add_parser("emb", help="measure EmbeddingBag performance")
    p_emb.add_argument("-d", "--dataset", choices=["A", "B", "cfg1"], default="A")
    p_emb.add_argument("--randomseed", type=int, default=0)
    p_emb.add_argument("--usexlabag", action="store_true")
    p_emb.add_argument("--alpha", type=float, default=0.0, help="Zipf param. Use uniform if == 0.0")
    p_emb.add_argument("--fast-indices", action="store_true")
    for other in ("gemm", "linear"):
        sub.add_parser(other, help="not on the B200 hot path — use the reference driver")
    args = ap.parse_args(argv)
    print("Measuring the performance of ", args.kernel, " on device = ", args.device)
    print("Steps = ", args.steps, " warmups = ", args.warmups)
    if args.kernel != "emb":
        raise SystemExit(f"kernel {args.kernel!r} is outside the B200 hot path (emb only)")
    print("with emb dataset ", args.dataset)
    data = {"A": dataset.emb_A, "B": dataset.emb_B, "cfg1": dataset.emb_cfg1}[args.dataset]
    kemb.run(args, data)


if __name__ == "__main__":
    main()
