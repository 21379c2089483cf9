"""Measure pinned-host <-> device copy bandwidth on the GPU box (context for bench.py's e2e number)."""
import json
import time

import torch

torch.cuda.init()
dev = torch.device("cuda", 0)
out = {}
for mb in (8, 32, 172, 512):
    n = mb * 1000 * 1000
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    h2d = n * 10 / (e0.elapsed_time(e1) / 1e3) / 1e9
    e0.record()
    for _ in range(10):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    d2h = n * 10 / (e0.elapsed_time(e1) / 1e3) / 1e9
    # both directions at once on two streams
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    both = n * 10 / (time.perf_counter() - t0) / 1e9
    out["%d MB" % mb] = {"h2d_GBps": round(h2d, 1), "d2h_GBps": round(d2h, 1), "each_direction_when_both_GBps": round(both, 1)}
print(json.dumps(out))
