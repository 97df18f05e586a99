#!/usr/bin/env python3
"""NVLink evidence for the two collective kernels without a profiler (ncu replays a kernel ~40 times, and a replayed
kernel that waits for peer flags cannot make progress): the per-link data counters of `nvidia-smi nvlink -gt d` are
read on every rank before and after K calls of (a) the pooled exchange (a2a_push_kernel) and (b) the fused lookup +
exchange (tbe_fwd_a2a_kernel); bytes per call on the wire are compared with what the algorithm must move
(S * (W - 1) / W per direction) and the CUDA-event time gives GB/s per direction per GPU against 900 nominal.

    torchrun --nproc-per-node N tools/nvlink_probe.py [--tables 64 --local-batch 8192 --rows 1000000]
"""
import argparse
import os
import re
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def nvlink_kib(index):
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True,
                             timeout=30).stdout
    except Exception as exc:  # noqa: BLE001
        return None, None, str(exc)
    tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    links = len(re.findall(r"Data Tx:", out))
    return tx, rx, links


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tables", type=int, default=64)
    ap.add_argument("--local-batch", type=int, default=8192)
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--bag", type=int, default=20)
    ap.add_argument("--calls", type=int, default=20)
    a = ap.parse_args()
    from param_b200.comms.pt.dlrm import DLRMParallelEmbedding, SparseBatch
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    T_g = a.tables * world
    model = DLRMParallelEmbedding(dist.group.WORLD, [a.rows] * T_g, a.dim, a.local_batch, a.bag, dev, lr=1e-6, seed=3)
    batch = SparseBatch.synthetic([a.rows] * T_g, a.local_batch, a.bag, True, seed=17 + rank, device=dev, alpha=1.15)
    offsets, indices = model.sparse_data_dist(batch)
    win = model.window
    N = a.local_batch * world
    pooled = torch.randn(N, a.tables * a.dim, device=dev)
    S = a.local_batch * T_g * a.dim * 4                       # bytes of this rank's output tensor
    expect = S * (world - 1) / world                          # leaves (and enters) this GPU per call

    def measure(tag, fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        tx0, rx0, links = nvlink_kib(lr)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.calls):
            fn()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        tx1, rx1, _ = nvlink_kib(lr)
        ms = e0.elapsed_time(e1) / a.calls
        if tx0 is None:
            print(f"[nvlink_probe] rank {rank} {tag}: counters unavailable ({links})", flush=True)
            return
        tx = (tx1 - tx0) * 1024 / a.calls
        rx = (rx1 - rx0) * 1024 / a.calls
        print(f"[nvlink_probe] rank {rank} {tag}: {ms:.3f} ms/call; NVLink tx {tx / 1e6:.1f} MB/call rx {rx / 1e6:.1f} MB/call "
              f"over {links} links (algorithm: {expect / 1e6:.1f} MB each way, tx ratio {tx / expect:.3f}); "
              f"{expect / ms / 1e6:.1f} GB/s per direction = {expect / ms / 1e6 / 900:.3f} of 900 nominal", flush=True)

    measure("a2a_push_kernel (pooled exchange fwd)",
            lambda: win.pooled_forward(pooled, model.batch_split, model.tables_split, a.dim, out_window_off=model.off_pooled))
    measure("tbe_fwd_a2a_kernel (fused lookup + exchange)",
            lambda: win.lookup_forward_fused(model.arena, indices, offsets, model.batch_split, model.tables_split,
                                             out_window_off=model.off_pooled))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
