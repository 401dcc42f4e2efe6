#!/usr/bin/env python
"""What the host link of the GPU box sustains: pinned host<->device copies of the sizes the headline end-to-end
step moves (160 MB in, 216 MB out per 1M Sentinel-2A samples, compact FP64 result), alone and both directions at
once, whole and in 64 Ki-sample pieces.  Names the limiter of bench.py's `e2e`.
usage: python tools/pcie_probe.py   (run on the GPU box)"""
import json
import time

import torch

dev = torch.device("cuda", 0)
MB = 1 << 20
h2d_bytes, d2h_bytes = 160_000_000, 216_000_000
hin = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
hout = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
din = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
dout = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def h2d(pieces=1):
    with torch.cuda.stream(s1):
        step = h2d_bytes // pieces
        for i in range(pieces):
            din[i * step:(i + 1) * step].copy_(hin[i * step:(i + 1) * step], non_blocking=True)


def d2h(pieces=1):
    with torch.cuda.stream(s2):
        step = d2h_bytes // pieces
        for i in range(pieces):
            hout[i * step:(i + 1) * step].copy_(dout[i * step:(i + 1) * step], non_blocking=True)


res = {}
for pieces in (1, 16):
    t_in = timed(lambda: h2d(pieces))
    t_out = timed(lambda: d2h(pieces))
    t_both = timed(lambda: (h2d(pieces), d2h(pieces)))
    res[f"pieces_{pieces}"] = {
        "h2d_alone_GBps": h2d_bytes / t_in / 1e9, "d2h_alone_GBps": d2h_bytes / t_out / 1e9,
        "both_ms": t_both * 1e3, "both_h2d_plus_d2h_GBps": (h2d_bytes + d2h_bytes) / t_both / 1e9,
        "simulations_per_s_if_only_copies": 1e6 / t_both}
res["note"] = "160 MB host->device and 216 MB device->host per 1M simulations (bench.py e2e, compact FP64 result)"
print(json.dumps(res))
