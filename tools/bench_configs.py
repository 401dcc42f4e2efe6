#!/usr/bin/env python
"""Resident-input throughput of the other BASELINE.json configurations (the headline bench.py line
is configs[1]).  Prints one JSON object; tools/profiles_post.py stores it under profiles/.
usage: python tools/bench_configs.py [steps]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "spart-python_b200"))
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
import spart_b200  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
eng = spart_b200.default_engine(dev)


def params(n, config):
    P = bench.synthetic_params_torch(n, 20261018 + config, dev)
    g = torch.Generator(device=dev).manual_seed(99 + config)
    u = lambda lo, hi: torch.rand(n, generator=g, device=dev, dtype=torch.float64) * (hi - lo) + lo
    if config == 3:      # PROSPECT-PRO leaves, random sun / view angles
        P[1] = 0.0
        P[7] = u(0, 0.003)
        P[8] = u(0, 0.01)
        P[19], P[20], P[21] = u(0, 65), u(0, 40), u(0, 180)
    return P


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_step": ms, "simulations_per_s": n / (ms * 1e-3)}


res = {}
n = 1_000_000
for prec in ("fp64", "fp32"):
    P = params(n, 2)
    out = torch.empty((n, 13, 3), dtype=torch.float64, device=dev)
    res[f"cfg2_S2A_fixed_geometry_{prec}"] = timed(
        lambda: eng.forward_bands(P, "Sentinel2A-MSI", out=out, uniform_geometry=True, precision=prec), n)
    P3 = params(n, 3)
    out3 = torch.empty((n, 9, 3), dtype=torch.float64, device=dev)
    res[f"cfg3_L8_PRO_random_angles_{prec}"] = timed(
        lambda: eng.forward_bands(P3, "LANDSAT8-OLI", out=out3, precision=prec), n)
    outb = torch.empty((n, 13, 3), dtype=torch.float64, device=dev)

    def both():
        eng.forward_bands_multi(P, ["Sentinel2A-MSI", "Sentinel2B-MSI"], outs=[out, outb], uniform_geometry=True,
                                precision=prec)
    res[f"cfg5_S2A+S2B_{prec}"] = timed(both, n)
    n4 = 100_000
    info = spart_b200.synthetic_fullspectrum_sensorinfo()
    P4 = params(n4, 4)
    out4 = torch.empty((n4, 2001, 3), dtype=torch.float64, device=dev)
    r = timed(lambda: eng.forward_bands(P4, info, out=out4, uniform_geometry=True, precision=prec), n4)
    r["output_gb_per_s"] = n4 * 2001 * 3 * 8 / (r["ms_per_step"] * 1e-3) / 1e9
    res[f"cfg4_fullspectrum_2001_bands_{prec}"] = r
    del out4
ns = 100_000
Ps = params(ns, 2)
outs = torch.empty((ns, 13, 3), dtype=torch.float64, device=dev)
res["srf_band_mode_S2A_fp64"] = timed(
    lambda: eng.forward_bands(Ps, "Sentinel2A-MSI", out=outs, uniform_geometry=True, band_mode="srf"), ns)
nspec = 20_000
Psp = params(nspec, 2)
outsp = torch.empty((nspec, 9, 2162), dtype=torch.float64, device=dev)
r = timed(lambda: eng.forward_spectrum(Psp, out=outsp), nspec)
r["output_gb_per_s"] = nspec * 9 * 2162 * 8 / (r["ms_per_step"] * 1e-3) / 1e9
res["canopyopt_spectra_9x2162_fp64"] = r
# retrieval: nearest LUT entry for 100k observations against a 1M-entry Sentinel-2A LUT
g = torch.Generator(device=dev).manual_seed(3)
L = torch.rand((1_000_000, 13), generator=g, device=dev, dtype=torch.float32)
O = torch.rand((100_000, 13), generator=g, device=dev, dtype=torch.float32)
r = timed(lambda: spart_b200.lut.nearest(L, O), 100_000)
pairs = 1e6 * 1e5
r["pairs_per_s"] = pairs / (r["ms_per_step"] * 1e-3)
r["fp32_tflops"] = pairs * 13 * 3 / (r["ms_per_step"] * 1e-3) / 1e12      # sub + fma per band
res["lut_nearest_100k_obs_x_1M_entries_13_bands"] = r
print(json.dumps(res, indent=1))
