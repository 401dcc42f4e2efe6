#!/usr/bin/env python
"""Timeline of one host-path call (SPART_HOST_TRACE=1): per chunk, when its H2D, kernels and D2H finished.
usage (GPU box): [SPART_HOST_CHUNK=...] python tools/e2e_trace.py"""
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
hin = torch.empty((27, n), dtype=torch.float64).pin_memory()
hout = torch.empty(n * 13 * 2 + n, dtype=torch.float64).pin_memory()
hin.copy_(P)
call = lambda: sb.run_batch_params(hin, "Sentinel2A-MSI", out=hout, broadcast_rows=cfg["bcast"], compact=True)
for _ in range(3):
    call()
t0 = time.perf_counter()
for _ in range(5):
    call()
print("ms per call", (time.perf_counter() - t0) / 5 * 1e3, flush=True)
os.environ["SPART_HOST_TRACE"] = "1"
call()
