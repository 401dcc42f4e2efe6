#!/usr/bin/env python
"""Small run of every kernel for compute-sanitizer (memcheck): FP64 / FP32 / SRF band modes, a ragged
batch size, the spectrum, SAILH-stage and leaf-angle entry points."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "spart-python_b200"))
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import spart_b200 as sb  # noqa: E402

dev = torch.device("cuda", 0)
n = 4097
P = bench.synthetic_params_torch(n, 2, 1, dev)
eng = sb.default_engine(dev)
for prec in ("fp64", "fp32"):
    for uni in (False, True):
        eng.forward_bands(P, "TerraAqua-MODIS", precision=prec, uniform_geometry=uni)
eng.forward_bands(P[:, :257].contiguous(), "Sentinel2A-MSI", band_mode="srf")
eng.forward_bands_multi(P, ["Sentinel2A-MSI", "Sentinel2B-MSI"])
eng.forward_spectrum(P[:, :65].contiguous())
spec = torch.rand(2162, dtype=torch.float64, device=dev) * 0.4 + 0.05
eng.sailh(P[:, :33].contiguous(), spec, spec, spec * 0.9)
eng.leafangles(np.random.default_rng(0).uniform(-0.5, 0.5, (131, 2)))
host = sb.run_batch_params(P.cpu().numpy(), "LANDSAT8-OLI")
Pb = P.clone()
Pb[16], Pb[17] = -0.35, -0.15
eng.forward_bands(Pb, "LANDSAT8-OLI", broadcast_rows=(16, 17))            # leaf angles once per batch
torch.cuda.synchronize()
print("sanitize run ok", host.shape)
