#!/usr/bin/env python
"""Kernel timing harness for tuning: per-kernel CUDA-event times of spart_forward_bands on the
bench workload.  SPART_B200_LIB selects an alternative build of the library.
usage: python tools/kbench.py [n] [sensor] [config]"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "spart-python_b200"))
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
import spart_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sensor = sys.argv[2] if len(sys.argv) > 2 else "Sentinel2A-MSI"
dev = torch.device("cuda", 0)
eng = spart_b200.default_engine(dev)
P = bench.synthetic_params_torch(n, 2, 123, dev)
if len(sys.argv) > 3 and sys.argv[3] == "3":      # random geometry
    g = torch.Generator(device=dev).manual_seed(5)
    P[19] = torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 65
    P[20] = torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 40
    P[21] = torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 180
_, st = eng.sensor(sensor)
out = torch.empty((n, st.n_bands, 3), dtype=torch.float64, device=dev)
precision = sys.argv[4] if len(sys.argv) > 4 else "fp64"
uniform = not (len(sys.argv) > 3 and sys.argv[3] == "3") and os.environ.get("SPART_NO_UNIFORM") is None
bc = bench.CONFIGS[2]["bcast"] if uniform else ()
for _ in range(3):
    eng.forward_bands(P, sensor, out=out, broadcast_rows=bc, precision=precision)
torch.cuda.synchronize()
eng.profile_enable(sensor, True)
for _ in range(8):
    eng.forward_bands(P, sensor, out=out, broadcast_rows=bc, precision=precision)
torch.cuda.synchronize()
r = eng.profile_read(sensor)
c = r["calls"]
print(json.dumps({"lib": os.environ.get("SPART_B200_LIB", "default"), "n": n, "sensor": sensor,
                  "uniform": uniform, "precision": precision, "lidf_ms": r["lidf_ms"] / c, "geometry_ms": r["geometry_ms"] / c,
                  "band_ms": r["band_ms"] / c,
                  "checksum": float(out.sum().item())}))
