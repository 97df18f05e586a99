"""CPU arm of the benchmark: the reference's own CPU implementation of the path, timed on the
box's host cores.  TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference
legs); nothing under param_b200/ imports this.

The reference computes this path by calling torch.nn.EmbeddingBag on the host
(train/compute/pt/pytorch_emb.py:179 constructs it, measure_cpu :37-45 is the timed loop:
`for i in range(warmups + steps): results = h_emb(h_indices, h_offsets)` with the clock restarted
after the warm-ups) and autograd for the backward (train/comms/pt/pytorch_dist_backend.py:849-857;
tables are sparse=True there, :924).  /root/reference does not exist on the GPU box and the
reference is pure Python over torch, so the loop is restated here around the same torch op.
"""
from __future__ import annotations

import os
import time

import torch
import torch.nn as nn


def time_embeddingbag_cpu(weights, indices, offsets, steps: int, warmups: int, backward: bool,
                          threads: int | None = None):
    """weights: list of fp32 CPU tensors [rows, dim] (one per table); indices/offsets: lists of
    int64 CPU tensors (nn.EmbeddingBag contract).  Returns (seconds per step over all tables,
    threads used).  One step = forward (+ backward with a ones gradient) over every table."""
    if threads is None:
        threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    embs = []
    for w in weights:
        e = nn.EmbeddingBag(w.shape[0], w.shape[1], mode="sum", sparse=True, _weight=w)
        e.weight.requires_grad_(backward)
        embs.append(e)
    grads = [torch.ones(off.numel(), w.shape[1]) for off, w in zip(offsets, weights)] if backward else None
    start = time.perf_counter()
    for i in range(warmups + steps):
        for t, e in enumerate(embs):
            res = e(indices[t], offsets[t])
            if backward:
                e.weight.grad = None
                res.backward(grads[t])
        if i < warmups:
            start = time.perf_counter()
    end = time.perf_counter()
    return (end - start) / max(steps, 1), torch.get_num_threads()
