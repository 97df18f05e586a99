#!/usr/bin/env python3
"""Writes a DLRM-step-shaped "basic" comms trace (the format train/comms/pt/commsTraceParser.py:_parseBasicTrace
reads) for a given world size: index exchange, TBE lookup ("compute": "emb_lookup", :137-147), pooled exchange, the
transpose exchange, TBE backward.  Sizes are in ELEMENTS per rank, as the reference's parser expects.
    python tools/make_basic_trace.py --world 8 --out gpurun_out/dlrm_step_basic.json"""
import argparse
import json

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, required=True)
ap.add_argument("--out", required=True)
ap.add_argument("--tables", type=int, default=8, help="tables per rank")
ap.add_argument("--local-batch", type=int, default=2048)
ap.add_argument("--bag", type=int, default=20)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--rows", type=int, default=200000)
a = ap.parse_args()
W, T, b, L, E = a.world, a.tables, a.local_batch, a.bag, a.dim
N = b * W
idx_elems = T * W * b * L            # this rank sends the indices of all T*W tables for its b samples
pooled = T * N * E                   # [N, T*E] pooled rows leave each rank (and as many arrive)


def comm(markers, name, elems, dtype, req, t):
    return {"markers": markers, "comms": name, "in_msg_size": elems, "out_msg_size": elems, "dtype": dtype,
            "req": req, "startTime_ns": t, "world_size": W}


def emb(markers, direction):
    return {"markers": markers, "compute": "emb_lookup", "direction": direction, "emb_dim": E, "num_embs": a.rows,
            "batch_size": N, "num_emb_tables": T, "bag_size": L, "count": 1}


trace = [
    comm(["dlrm_fwd"], "all_to_all", idx_elems, "Long", 0, 0),
    {"markers": ["dlrm_fwd"], "comms": "wait", "req": 0, "startTime_ns": 1000, "world_size": W},
    emb(["dlrm_fwd"], "forward"),
    comm(["dlrm_fwd"], "all_to_all", pooled, "Float", 1, 2000),
    {"markers": ["dlrm_fwd"], "comms": "wait", "req": 1, "startTime_ns": 3000, "world_size": W},
    comm(["dlrm_bwd"], "all_to_all", pooled, "Float", 2, 4000),
    {"markers": ["dlrm_bwd"], "comms": "wait", "req": 2, "startTime_ns": 5000, "world_size": W},
    emb(["dlrm_bwd"], "backward"),
]
json.dump(trace, open(a.out, "w"), indent=1)
print(f"wrote {a.out}: {len(trace)} entries, world {W}")
