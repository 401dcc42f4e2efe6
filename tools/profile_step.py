#!/usr/bin/env python
"""One step of a BASELINE configuration for ncu: a warm-up step outside the profiled range, then one step
between cudaProfilerStart / Stop (run ncu with --profile-from-start off).
usage: python tools/profile_step.py --config 2 [--precision fp32] [--n 262144] [--mode bands|srf|spectrum|lut]"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "spart-python_b200"))
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
import spart_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--precision", default="fp64")
ap.add_argument("--n", type=int, default=262144)
ap.add_argument("--mode", default="bands")
a = ap.parse_args()
dev = torch.device("cuda", 0)
eng = spart_b200.default_engine(dev)
cfg = bench.CONFIGS[a.config]
sensors = [spart_b200.synthetic_fullspectrum_sensorinfo() if s == "SYNTH2001" else s for s in cfg["sensors"]]
n = a.n if a.config != 4 else min(a.n, 32768)
P = bench.synthetic_params_torch(n, a.config, 20261018 + a.config, dev)
ws = eng.workspace(n)

if a.mode == "bands":
    def step():
        for i, s in enumerate(sensors):
            eng.forward_bands(P, s, precision=a.precision, broadcast_rows=cfg["bcast"], reuse_record=i > 0, workspace=ws)
elif a.mode == "srf":
    def step():
        eng.forward_bands(P, sensors[0], band_mode="srf", broadcast_rows=cfg["bcast"])
elif a.mode == "spectrum":
    P = P[:, :4096].contiguous()

    def step():
        eng.forward_spectrum(P)
elif a.mode == "lut":
    g = torch.Generator(device=dev).manual_seed(3)
    L = torch.rand((1_000_000, 13), generator=g, device=dev, dtype=torch.float32)
    O = torch.rand((20_000, 13), generator=g, device=dev, dtype=torch.float32)

    def step():
        spart_b200.lut.nearest(L, O)
step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", a.mode, "config", a.config, a.precision, "n", n)
