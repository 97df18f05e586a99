#!/usr/bin/env python3
"""Sweep the push kernel's grid size for all_to_all_single and the pooled exchange (torchrun, N GPUs)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from param_b200.comms.pt.peer_window import PeerWindow  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
win = PeerWindow.create(dist.group.WORLD, 3 << 30, dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


res = []
f = (world - 1) / world
for size in (64 << 20, 512 << 20):
    n = size // 4 // world * world
    x = torch.randn(n, device=dev)
    out = win.view(0, n, torch.float32)
    ref = torch.empty(n, device=dev)
    t_n = timeit(lambda: dist.all_to_all_single(ref, x))
    for ctas in (16, 32, 64, 96, 148):
        win.configure(max_ctas=ctas)
        t = timeit(lambda: win.all_to_all_single(out, x))
        res.append({"op": "all_to_all_single", "bytes": n * 4, "ctas": ctas, "us": t * 1e3,
                    "busbw_gbs": n * 4 / t / 1e6 * f, "nccl_busbw_gbs": n * 4 / t_n / 1e6 * f})
# pooled exchange at cfg4 shape
T_l, b, E = 64, 8192, 128
N, T_g = b * world, T_l * world
pooled = torch.randn(N, T_l * E, device=dev)
S = b * T_g * E * 4
for ctas in (16, 32, 64, 96, 148):
    win.configure(max_ctas=ctas)
    t = timeit(lambda: win.pooled_forward(pooled, [b] * world, [T_l] * world, E, out_window_off=0))
    res.append({"op": "pooled_forward", "bytes": S, "ctas": ctas, "us": t * 1e3, "busbw_gbs": S / t / 1e6 * f})
if rank == 0:
    for r in res:
        print(json.dumps(r))
assert win.error() == 0
dist.barrier()
dist.destroy_process_group()
