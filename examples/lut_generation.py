"""Look-up-table generation: one million parameter sets for Sentinel-2A and -2B in one call.

    PYTHONPATH=spart-python_b200 python examples/lut_generation.py [n] [fp64|fp32]
"""
import sys
import time

import numpy as np
import torch

import SPART

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
precision = sys.argv[2] if len(sys.argv) > 2 else "fp64"
rng = np.random.default_rng(0)
u = rng.uniform
leaf = np.stack([u(5, 80, n), u(0.002, 0.02, n), u(0.005, 0.05, n), u(0, 0.5, n), u(1, 25, n), u(0, 10, n),
                 u(1, 3, n)], axis=1)                                   # Cab Cdm Cw Cs Cca Cant N
soil = np.stack([u(0.2, 0.8, n), u(0, 25, n), u(90, 115, n), u(5, 55, n)], axis=1)   # B lat lon SMp
canopy = np.stack([u(0.1, 8, n), u(-0.5, 0.5, n), u(-0.5, 0.5, n), u(0.01, 0.2, n)], axis=1)
atm = np.stack([u(0.05, 0.6, n), u(0.25, 0.45, n), u(0.5, 4, n), u(900, 1030, n)], axis=1)

params = torch.from_numpy(SPART.pack_batch(leaf, soil, canopy, [40.0, 0.0, 0.0], atm, 180)).cuda()
torch.cuda.synchronize()
t0 = time.perf_counter()
s2a, s2b = SPART.run_batch_params(params, ["Sentinel2A-MSI", "Sentinel2B-MSI"], precision=precision,
                                  uniform_geometry=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{n} simulations x 2 sensors in {dt * 1e3:.1f} ms ({n / dt / 1e6:.1f} M simulations/s), "
      f"R_TOC band 8 mean {float(s2a[:, 7, 0].mean()):.4f}")
np.savez_compressed("lut_s2.npz", params=params.cpu().numpy().T.astype(np.float32),
                    s2a=s2a.cpu().numpy().astype(np.float32), s2b=s2b.cpu().numpy().astype(np.float32))
print("wrote lut_s2.npz")
