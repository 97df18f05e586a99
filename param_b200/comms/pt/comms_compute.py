"""Compute / communication overlap benchmark with the `emb_lookup` kernel on the B200 backend.

Mirror of the reference runner train/comms/pt/commsComputeBench.py for the one compute kernel on the
hot path (`--kernel emb_lookup`, commsComputeBench.py:303-312): same flags (--mode compute |
comms-compute, --num-compute, --emb-dim, --num-embs, --batch-size, --num-emb-tables-per-device,
--num-emb-tables-batched, --bag-size, plus the collective flags of comms.py) and the same iteration
(runColl, commsComputeBench.py:155-257): numCollPerIter collectives on the current stream, then
num_compute calls of backendFuncs.emb_lookup(collectiveArgs) on the compute stream, one
sync_barrier per iteration; wall clock per iteration (time.monotonic) and a device timer per stream
(paramDeviceTimer -> CUDA events here).  The reference needs fbgemm_gpu for this kernel
(comms_utils.py:1967-1981); here init_emb_lookup builds B200TBE ops (emb_lookup.py), so the
benchmark runs without it.  The other compute kernels of the reference (gemm, add, copy, ...) are
dense / elementwise work outside the hot path and are not mirrored.

  torchrun --nproc-per-node 8 -m param_b200.comms.pt.comms_compute --mode comms-compute \
      --kernel emb_lookup --collective all_to_all_single --b 16M --e 256M --f 4 \
      --num-compute 4 --emb-dim 128 --num-embs 1000000 --batch-size 8192 --ntables 64 --bag-size 20
"""
from __future__ import annotations

import argparse
import json
import os
import time
import types

import torch
import torch.distributed as dist

from ..._cabi import PB200Error
from .backend import B200Backend
from .comms import _DTYPES, _payload, get_sizes, parsesize
from .emb_lookup import init_emb_lookup


def _args(argv=None):
    ap = argparse.ArgumentParser(description="PARAM commsComputeBench-style overlap benchmark (B200, emb_lookup)")
    ap.add_argument("--mode", default="comms-compute", choices=["compute", "comms-compute"])
    ap.add_argument("--kernel", default="emb_lookup", choices=["emb_lookup"])
    ap.add_argument("--num-compute", "--num-compute-per-iteration", dest="num_compute", type=int, default=100)
    ap.add_argument("--num-coll", "--num-coll-per-iteration", dest="num_coll", type=int, default=1)
    ap.add_argument("--emb-dim", type=int, default=128)
    ap.add_argument("--num-embs", type=int, default=100000)
    ap.add_argument("--batch-size", type=int, default=512)
    ap.add_argument("--num-emb-tables-per-device", "--ntables", "--num-emb-tables", dest="num_emb_tables_per_device",
                    type=int, default=8)
    ap.add_argument("--num-emb-tables-batched", type=int, default=-1)
    ap.add_argument("--bag-size", type=int, default=20)
    ap.add_argument("--direction", default="forward", choices=["forward", "backward"])
    ap.add_argument("--emb-optimizer", default="exact_row_wise_adagrad")
    ap.add_argument("--collective", default="all_to_all_single", choices=["all_to_all_single", "all_to_allv"])
    ap.add_argument("--b", "--begin-size", dest="b", default="1M")
    ap.add_argument("--e", "--end-size", dest="e", default="64M")
    ap.add_argument("--f", "--step-factor", dest="f", type=int, default=4)
    ap.add_argument("--n", "--num-iters", dest="n", type=int, default=10)
    ap.add_argument("--w", "--num_warmup_iters", dest="w", type=int, default=3)
    ap.add_argument("--data-type", default="float32", choices=sorted(_DTYPES))
    ap.add_argument("--backend", default="b200", choices=["b200"])
    ap.add_argument("--master-ip", default=os.environ.get("MASTER_ADDR", "127.0.0.1"))
    ap.add_argument("--master-port", default=os.environ.get("MASTER_PORT", "29500"))
    ap.add_argument("--json", action="store_true")
    return ap.parse_args(argv)


class _DeviceTimer:
    """accumulating CUDA-event timer, the role of comms_utils.paramDeviceTimer"""

    def __init__(self):
        self.pairs, self.total_ms = [], 0.0

    def start(self, stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        self.pairs.append((e0, e1))

    def stop(self, stream):
        self.pairs[-1][1].record(stream)

    def collect(self):
        for e0, e1 in self.pairs:
            self.total_ms += e0.elapsed_time(e1)
        self.pairs.clear()

    def reset(self):
        self.pairs.clear()
        self.total_ms = 0.0


def run(argv=None):
    a = _args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    boot = types.SimpleNamespace(global_rank=rank, local_rank=local_rank, world_size=world, local_size=world,
                                 master_ip=a.master_ip, master_port=a.master_port)
    params = types.SimpleNamespace(device="cuda", backend="nccl", use_ext_dist=False, init_only=False,
                                   direction=a.direction, emb_dim=a.emb_dim, num_embs=a.num_embs,
                                   batch_size=a.batch_size, num_emb_tables_per_device=a.num_emb_tables_per_device,
                                   num_emb_tables_batched=a.num_emb_tables_batched, bag_size=a.bag_size,
                                   emb_optimizer=a.emb_optimizer)
    be = B200Backend(boot, params)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    be.groups = {0: dist.GroupMember.WORLD}
    dtype = _DTYPES[a.data_type]
    es = torch.empty(0, dtype=dtype).element_size()
    comms_on = a.mode == "comms-compute"
    sizes = get_sizes(parsesize(a.b), parsesize(a.e), a.f) if comms_on else [0]
    compute_stream = torch.cuda.Stream(device=dev)
    ca = types.SimpleNamespace(group=dist.GroupMember.WORLD, asyncOp=False, waitObj=[], waitObjIds={},
                               ipTensor_split=[], opTensor_split=[], device=dev, world_size=world,
                               all2all_qcomm=None, collective=a.collective, reuseTensors=True)
    # comms_utils.init_emb_lookup: emb, embRequests, LookupOut, grad_output.  Done on the compute stream:
    # autograd replays a backward on the stream its forward ran on, and the backward direction
    # differentiates the LookupOut created here.
    with torch.cuda.stream(compute_stream):
        init_emb_lookup(ca, params, be)
    compute_stream.synchronize()
    lookups_per_compute = sum(int(idx.numel()) for idx, _, _ in ca.embRequests)
    comm_fn = {"all_to_all_single": be.all_to_all_single, "all_to_allv": be.all_to_allv}[a.collective]
    results = []
    if rank == 0 and not a.json:
        print(f"# mode {a.mode} kernel {a.kernel} collective {a.collective} world {world} num_coll {a.num_coll} "
              f"num_compute {a.num_compute} emb_dim {a.emb_dim} num_embs {a.num_embs} batch_size {a.batch_size} "
              f"tables {a.num_emb_tables_per_device} bag {a.bag_size} direction {a.direction}")
        print("COMMS-COMPUTE-RES-HDR  size(B)  iter(us)  comm_dev(us)  compute_dev(us)  overlap  algBW(GB/s)  "
              "busBW(GB/s)  lookups/s")
    for size in sizes:
        numel = max(size // es // world, 1) * world
        if comms_on:
            be.clear_memory(ca)
            ca.ipTensor = be.alloc_empty(numel, dev, dtype)
            ca.opTensor = be.alloc_empty(numel, dev, dtype)
            ca.ipTensor.copy_(_payload(rank, numel, dtype, dev))
            if a.collective == "all_to_allv":
                ca.ipTensor_split = [numel // world] * world
                ca.opTensor_split = [numel // world] * world
        comm_t, comp_t, span_t = _DeviceTimer(), _DeviceTimer(), _DeviceTimer()
        elapsed = 0.0
        for it in range(a.w + a.n):
            if it == a.w:
                torch.cuda.synchronize(dev)
                elapsed = 0.0
                comm_t.reset()
                comp_t.reset()
                span_t.reset()
            cur = torch.cuda.current_stream(dev)
            compute_stream.wait_stream(cur)          # both legs start from the same point
            start = time.monotonic()
            span_t.start(cur)                        # device time from the common start to the join of both legs
            if comms_on:
                comm_t.start(cur)
                for _ in range(a.num_coll):
                    comm_fn(ca)
                be.complete_accel_ops(ca, devSync=False)
                comm_t.stop(cur)
            with torch.cuda.stream(compute_stream):
                comp_t.start(compute_stream)
                for _ in range(a.num_compute):
                    be.emb_lookup(ca)
                comp_t.stop(compute_stream)
            cur.wait_stream(compute_stream)
            span_t.stop(cur)
            torch.cuda.synchronize(dev)
            dist.barrier(device_ids=[local_rank])    # sync_barrier(desc="runColl_sync")
            elapsed += time.monotonic() - start
            comm_t.collect()
            comp_t.collect()
            span_t.collect()
        it_us = elapsed / a.n * 1e6
        comm_us, comp_us = comm_t.total_ms / a.n * 1e3, comp_t.total_ms / a.n * 1e3
        span_us = span_t.total_ms / a.n * 1e3
        t = torch.tensor([it_us, comm_us, comp_us, span_us], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        it_us, comm_us, comp_us, span_us = (float(x) for x in t)
        nbytes = numel * es if comms_on else 0
        alg = nbytes * a.num_coll / (it_us * 1e-6) / 1e9 if comms_on else 0.0       # getAlgBW over the iteration
        bus = be.getBusBW(a.collective, alg, ca) if comms_on else 0.0
        # device-side: 1.0 = the shorter leg is completely hidden behind the longer one, 0.0 = the legs
        # ran back to back (iter_us is wall clock and also holds the host launches and the barrier)
        overlap = (comm_us + comp_us - span_us) / min(comm_us, comp_us) if comms_on and min(comm_us, comp_us) > 0 else 0.0
        rec = {"mode": a.mode, "kernel": a.kernel, "collective": a.collective if comms_on else None, "world": world,
               "size_bytes": nbytes, "iter_us": it_us, "comm_dev_us": comm_us, "compute_dev_us": comp_us,
               "dev_span_us": span_us,
               "overlap": overlap, "algbw_gbs": alg, "busbw_gbs": bus,
               "lookups_per_s": world * lookups_per_compute * a.num_compute / (it_us * 1e-6),
               "direction": a.direction, "emb_optimizer": a.emb_optimizer}
        results.append(rec)
        if rank == 0:
            if a.json:
                print(json.dumps(rec))
            else:
                print(f"COMMS-COMPUTE-RES  {nbytes:>12}  {it_us:>10.1f}  {comm_us:>10.1f}  {comp_us:>10.1f}  "
                      f"{overlap:>6.2f}  {alg:>9.2f}  {bus:>9.2f}  {rec['lookups_per_s']:.3e}")
    if be._window is not None and be._window.error():
        raise PB200Error("a peer wait timed out during the benchmark")
    return results


if __name__ == "__main__":
    run()
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
