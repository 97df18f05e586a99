#!/usr/bin/env python3
"""Protocol overhead of the push kernel on ONE GPU: W virtual ranks (local group), small messages.
Time = host-timed average over many back-to-back collectives across all W ranks (no NVLink involved:
this isolates handshake / fence / barrier costs of the kernel itself)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from param_b200.comms.pt.peer_window import PeerWindow  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
grp = PeerWindow.local_group(W, 64 << 20, dev, max_ctas=max(1, 144 // W), spin_timeout_s=10.0)
for size in (1 << 10, 64 << 10, 1 << 20, 16 << 20):
    n = size // 4 // W * W
    xs = [torch.randn(n, device=dev) for _ in range(W)]
    outs = [w.view(0, n, torch.float32) for w in grp.windows]
    iters = 50

    def once():
        for r, (w, st) in enumerate(zip(grp.windows, grp.streams)):
            with torch.cuda.stream(st):
                w.all_to_all_single(outs[r], xs[r], stream=st)

    for _ in range(5):
        once()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(W)]
    for r, st in enumerate(grp.streams):
        ev[r][0].record(st)
    for _ in range(iters):
        once()
    for r, st in enumerate(grp.streams):
        ev[r][1].record(st)
    torch.cuda.synchronize()
    us = max(a.elapsed_time(b) for a, b in ev) / iters * 1e3
    assert all(w.error() == 0 for w in grp.windows)
    print(f"W={W} bytes={n * 4:>9}  {us:8.2f} us per all_to_all_single (device time, max over virtual ranks)")
