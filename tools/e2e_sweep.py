#!/usr/bin/env python
"""Host-path tuning: end-to-end rate of run_batch_params on 1M config-2 samples for several span / chunk sizes
(SPART_HOST_SPAN, SPART_HOST_CHUNK; the first pair is the default), pinned and pageable buffers.
usage (GPU box): python tools/e2e_sweep.py"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "spart-python_b200"))
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
import spart_b200 as sb  # noqa: E402

dev = torch.device("cuda", 0)
n = 1_000_000
cfg = bench.CONFIGS[2]
P = bench.synthetic_params_torch(n, 2, 1, dev)
res = {}
for variant in ("compact", "compact_nobcast", "f32", "full", "pageable"):
    f32 = variant == "f32"
    compact = variant in ("compact", "f32", "compact_nobcast")
    bc = () if variant == "compact_nobcast" else cfg["bcast"]
    dt = torch.float32 if f32 else torch.float64
    hin = torch.empty((27, n), dtype=dt)
    hout = torch.empty(n * 13 * 2 + n if compact else (n, 13, 3), dtype=dt)
    if variant != "pageable":
        hin, hout = hin.pin_memory(), hout.pin_memory()
    hin.copy_(P.to(dt))
    src, dst = (hin.numpy(), hout.numpy()) if variant == "pageable" else (hin, hout)
    for span, chunk in ((131072, 65536), (131072, 131072), (262144, 65536), (262144, 262144)):
        os.environ["SPART_HOST_SPAN"] = str(span)
        os.environ["SPART_HOST_CHUNK"] = str(chunk)
        call = lambda: sb.run_batch_params(src, "Sentinel2A-MSI", out=dst, precision="fp32" if f32 else "fp64",
                                           broadcast_rows=bc, compact=compact)
        for _ in range(2):
            call()
        t0 = time.perf_counter()
        for _ in range(5):
            call()
        dt_s = (time.perf_counter() - t0) / 5
        res[f"{variant}_span{span}_chunk{chunk}"] = {"ms": dt_s * 1e3, "Msim_per_s": n / dt_s / 1e6}
        print(variant, span, chunk, res[f"{variant}_span{span}_chunk{chunk}"], flush=True)
print(json.dumps(res))
