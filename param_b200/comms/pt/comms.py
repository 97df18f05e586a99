"""nccl-tests-style size sweep for the all-to-all family on the B200 backend.

Mirror of the reference runner train/comms/pt/comms.py for the collectives on the hot path
(all_to_all_single / all_to_allv / all_to_all): same flags (--b --e --f --n --w --z --c
--collective --data-type --backend --device), same size progression (getSizes,
comms_utils.py:189-215), same per-size loop (run_coll_non_graph, comms.py:452-545: optional barrier,
comm_fn(collectiveArgs), complete_accel_ops), same bandwidth definitions (getAlgBW
comms_utils.py:168-186: bytes of the OUTPUT tensor / time; getBusBW pytorch_backend_utils.py:221-234:
algBW * (W-1)/W) and a COMMS-RES result line per size.  Differences, on purpose:
  * `--backend b200` runs the peer-push kernel; `--backend nccl` runs c10d/NCCL through the same loop
    (the comparator the reference itself would be on this box);
  * `--c 1` checks a POSITION-CODED payload (value = f(source rank, element index)); the reference's
    dcheck fills a constant and cannot detect a wrong permutation (comms_utils.py:997-1055);
  * `--graph-launches` captures the loop in a CUDA graph (comms.py:375-450) — the push kernel keeps
    its epoch in device memory so replays stay correct.

  torchrun --nproc-per-node 8 -m param_b200.comms.pt.comms --collective all_to_all_single \
      --b 1K --e 1G --f 2 --n 20 --w 5 --z 1 --c 1 --backend b200
"""
from __future__ import annotations

import argparse
import json
import os
import types
from typing import List

import torch
import torch.distributed as dist

from ..._cabi import PB200Error
from .backend import B200Backend

_UNITS = {"": 1, "K": 1 << 10, "M": 1 << 20, "G": 1 << 30}
_DTYPES = {"float32": torch.float32, "int32": torch.int32, "int64": torch.int64, "long": torch.int64,
           "float16": torch.float16, "bfloat16": torch.bfloat16, "uint8": torch.uint8, "int8": torch.int8}


def parsesize(text: str) -> int:
    """'1K' -> 1024 … (comms_utils.parsesize)."""
    text = str(text).strip().upper().rstrip("B")
    unit = text[-1] if text and text[-1] in "KMG" else ""
    return int(float(text[:-1] if unit else text) * _UNITS[unit])


def get_sizes(begin: int, end: int, step_factor: int, step_bytes: int = 0) -> List[int]:
    """geometric (x f) or arithmetic (+ step_bytes) progression, inclusive (getSizes)."""
    out, cur = [], begin
    while cur <= end:
        out.append(cur)
        cur = cur * step_factor if step_bytes == 0 else cur + step_bytes
        if step_bytes == 0 and step_factor <= 1:
            break
    return out


def _args(argv=None):
    ap = argparse.ArgumentParser(description="PARAM-Comm style all-to-all sweep (B200)")
    ap.add_argument("--collective", default="all_to_all_single",
                    choices=["all_to_all_single", "all_to_allv", "all_to_all"])
    ap.add_argument("--b", "--begin-size", dest="b", default="1K")
    ap.add_argument("--e", "--end-size", dest="e", default="256M")
    ap.add_argument("--f", "--step-factor", dest="f", type=int, default=2)
    ap.add_argument("--n", "--num-iters", dest="n", type=int, default=20)
    ap.add_argument("--w", "--num_warmup_iters", dest="w", type=int, default=5)
    ap.add_argument("--z", "--blocking", dest="z", type=int, default=1)
    ap.add_argument("--c", "--check-data", dest="c", type=int, default=0)
    ap.add_argument("--data-type", default="float32", choices=sorted(_DTYPES))
    ap.add_argument("--backend", default="b200", choices=["b200", "nccl"])
    ap.add_argument("--device", default="cuda", choices=["cuda"])
    ap.add_argument("--graph-launches", type=int, default=0)
    ap.add_argument("--master-ip", default=os.environ.get("MASTER_ADDR", "127.0.0.1"))
    ap.add_argument("--master-port", default=os.environ.get("MASTER_PORT", "29500"))
    ap.add_argument("--json", action="store_true", help="one JSON object per size instead of a table")
    return ap.parse_args(argv)


class _NcclA2A:
    """comparator: the reference's all_to_all_single body (pytorch_dist_backend.py:330-357)"""

    def __init__(self, device):
        self.device = device

    def alloc_empty(self, n, dev, dtype):
        return torch.empty(n, device=dev, dtype=dtype)

    def run(self, ca):
        work = dist.all_to_all_single(ca.opTensor, ca.ipTensor, ca.opTensor_split or None,
                                      ca.ipTensor_split or None, group=ca.group, async_op=ca.asyncOp)
        if ca.asyncOp:
            ca.waitObj.append(work)


def _payload(rank: int, numel: int, dtype, device):
    """position code: element k of rank r carries (r * 4099 + k) mod 16381 (exact in fp16..int64)"""
    k = torch.arange(numel, device=device, dtype=torch.int64)
    v = (k + rank * 4099) % (120 if dtype in (torch.uint8, torch.int8) else 16381 if dtype != torch.float16 else 2039)
    return v.to(dtype)


def _expected(rank: int, world: int, numel: int, dtype, device):
    """equal splits: my output block s = source s's elements [rank*chunk, (rank+1)*chunk)"""
    chunk = numel // world
    parts = []
    for s in range(world):
        k = torch.arange(rank * chunk, (rank + 1) * chunk, device=device, dtype=torch.int64)
        mod = 120 if dtype in (torch.uint8, torch.int8) else 16381 if dtype != torch.float16 else 2039
        parts.append(((k + s * 4099) % mod).to(dtype))
    return torch.cat(parts)


def run(argv=None):
    a = _args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local_rank)
    boot = types.SimpleNamespace(global_rank=rank, local_rank=local_rank, world_size=world,
                                 local_size=world, master_ip=a.master_ip, master_port=a.master_port)
    params = types.SimpleNamespace(device="cuda", backend="nccl", use_ext_dist=False, init_only=False)
    be = B200Backend(boot, params)
    torch.cuda.set_device(dev)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    be.groups = {0: dist.GroupMember.WORLD}
    dtype = _DTYPES[a.data_type]
    es = torch.empty(0, dtype=dtype).element_size()
    sizes = get_sizes(parsesize(a.b), parsesize(a.e), a.f)
    nccl = _NcclA2A(dev)
    results = []
    if rank == 0 and not a.json:
        print(f"# collective {a.collective} backend {a.backend} world {world} dtype {a.data_type} blocking {a.z}")
        print("COMMS-RES-HDR  size(B)  nElem  lat_p50(us)  lat_p95(us)  algBW(GB/s)  busBW(GB/s)  check")
    for size in sizes:
        numel = max(size // es // world, 1) * world          # equal splits (comms_utils.py:1212-1217)
        ca = types.SimpleNamespace(group=dist.GroupMember.WORLD, asyncOp=False, waitObj=[], waitObjIds={},
                                   ipTensor_split=[], opTensor_split=[], device=dev, world_size=world,
                                   all2all_qcomm=None, collective=a.collective)
        if a.backend == "b200":
            be.clear_memory(ca)
            ca.ipTensor = be.alloc_empty(numel, dev, dtype)
            ca.opTensor = be.alloc_empty(numel, dev, dtype)
            fn = {"all_to_all_single": be.all_to_all_single, "all_to_allv": be.all_to_allv,
                  "all_to_all": be.all_to_all}[a.collective]
        else:
            ca.ipTensor = torch.empty(numel, device=dev, dtype=dtype)
            ca.opTensor = torch.empty(numel, device=dev, dtype=dtype)
            fn = nccl.run
        ca.ipTensor.copy_(_payload(rank, numel, dtype, dev))
        ca.opTensor.zero_()
        if a.collective == "all_to_all":
            chunk = numel // world
            ca.ipTensor = list(ca.ipTensor.split(chunk))
            ca.opTensor = list(ca.opTensor.split(chunk))
            if a.backend == "nccl":
                fn = lambda c: dist.all_to_all(c.opTensor, c.ipTensor, group=c.group)  # noqa: E731
        elif a.collective == "all_to_allv":
            ca.ipTensor_split = [numel // world] * world
            ca.opTensor_split = [numel // world] * world
        lat = []
        graph = None
        if a.graph_launches > 0:
            st = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(st):
                for _ in range(3):
                    fn(ca)
                st.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=st):
                    for _ in range(a.n):
                        fn(ca)
        for it in range(a.w + a.n):
            if a.z:
                dist.barrier(device_ids=[local_rank])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            if graph is not None:
                graph.replay()
            else:
                fn(ca)
            e1.record()
            torch.cuda.synchronize()
            if it >= a.w:
                lat.append(e0.elapsed_time(e1) * 1e3 / (a.n if graph is not None else 1))
        ok = "-"
        if a.c:
            got = torch.cat([t.reshape(-1) for t in ca.opTensor]) if isinstance(ca.opTensor, list) else ca.opTensor
            ok = "PASS" if torch.equal(got, _expected(rank, world, numel, dtype, dev)) else "FAIL"
        # max over ranks of the per-rank percentiles (device time)
        t = torch.tensor(sorted(lat), device=dev)
        p50, p95 = t[len(t) // 2].view(1), t[min(len(t) - 1, int(len(t) * 0.95))].view(1)
        dist.all_reduce(p50, op=dist.ReduceOp.MAX)
        dist.all_reduce(p95, op=dist.ReduceOp.MAX)
        flag = torch.tensor([1 if ok in ("PASS", "-") else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        nbytes = numel * es
        alg = nbytes / (float(p50) * 1e-6) / 1e9
        bus = be.getBusBW(a.collective, alg, ca)
        rec = {"collective": a.collective, "backend": a.backend, "world": world, "size_bytes": nbytes,
               "num_elements": numel, "lat_p50_us": float(p50), "lat_p95_us": float(p95),
               "algbw_gbs": alg, "busbw_gbs": bus, "check": ("PASS" if int(flag) else "FAIL") if a.c else "-"}
        results.append(rec)
        if rank == 0:
            if a.json:
                print(json.dumps(rec))
            else:
                print(f"COMMS-RES  {nbytes:>12}  {numel:>11}  {float(p50):>10.2f}  {float(p95):>10.2f}  "
                      f"{alg:>10.2f}  {bus:>10.2f}  {rec['check']}")
        if a.c and not int(flag):
            raise PB200Error(f"data check failed at {nbytes} B")
    if a.backend == "b200" and be._window is not None and be._window.error():
        raise PB200Error("a peer wait timed out during the sweep")
    return results


if __name__ == "__main__":
    run()
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
