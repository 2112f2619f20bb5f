#!/usr/bin/env python3
"""Measure pinned host<->device copy bandwidth on this box (ceiling for bench.py's e2e figure)."""
import json
import time

import torch

n = 1592524800  # one 256 x 1080p RGB batch
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for nstreams in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    part = n // nstreams
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * part:(i + 1) * part].copy_(h[i * part:(i + 1) * part], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    out[f"h2d_{nstreams}_streams_GBps"] = round(n / dt / 1e9, 2)
torch.cuda.synchronize()
t0 = time.perf_counter()
h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
out["d2h_GBps"] = round(n / (time.perf_counter() - t0) / 1e9, 2)
print(json.dumps(out))
