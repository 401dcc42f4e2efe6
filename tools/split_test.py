#!/usr/bin/env python
"""Experiment: does splitting the 1M-sample step into k pieces on k streams (so that each kernel's partial
last wave overlaps another piece's kernels) shorten the step?  usage (GPU box): python tools/split_test.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "spart-python_b200"))
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
import spart_b200 as sb  # noqa: E402

dev = torch.device("cuda", 0)
eng = sb.default_engine(dev)
cfg = bench.CONFIGS[2]
n = 1_000_000
P = bench.synthetic_params_torch(n, 2, 1, dev)
out = torch.empty((n, 13, 3), dtype=torch.float64, device=dev)
ref = eng.forward_bands(P, "Sentinel2A-MSI", broadcast_rows=cfg["bcast"]).clone()


def timed(fn, steps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


print("single stream: %.4f ms" % timed(lambda: eng.forward_bands(P, "Sentinel2A-MSI", out=out, broadcast_rows=cfg["bcast"])))
for k in (2, 3, 4, 8):
    streams = [torch.cuda.Stream(dev) for _ in range(k)]
    bounds = [(i * n // k // 128 * 128, (i + 1) * n // k // 128 * 128 if i < k - 1 else n) for i in range(k)]
    wss = [eng.workspace(hi - lo) for lo, hi in bounds]

    def step():
        main = torch.cuda.current_stream(dev)
        ev = torch.cuda.Event()
        ev.record(main)
        for st, (lo, hi), ws in zip(streams, bounds, wss):
            st.wait_event(ev)
            with torch.cuda.stream(st):
                eng.forward_bands(P[:, lo:hi], "Sentinel2A-MSI", out=out[lo:hi], broadcast_rows=cfg["bcast"], workspace=ws)
        for st in streams:
            main.wait_stream(st)
    ms = timed(step)
    torch.cuda.synchronize()
    print("%d streams: %.4f ms  (bit-identical: %s)" % (k, ms, torch.equal(out, ref)))
