#!/usr/bin/env python
"""Convert the reference's static data tables (pickles) into neutral .npz files.

Run in the build container only (it reads /root/reference, which does not exist on the
GPU box).  Nothing is recomputed: every array keeps the dtype and shape it has in the
reference pickle, because NumPy promotion of the float32 Sentinel-2 SMAC coefficients is
part of the reference's observable behaviour.

Sources (reference file:line of the loaders):
  src/SPART/SPART.py:399-406  optical_params.pkl   -> data/optical.npz
  src/SPART/SPART.py:409-416  ET_irradiance.pkl    -> data/optical.npz (Ea, wl_Ea)
  src/SPART/SPART.py:419-424  sensor_information/* -> data/sensors/<sensor>.npz
"""
import pickle
import sys
import warnings
from pathlib import Path

import numpy as np

REF = Path("/root/reference/src/SPART")
OUT = Path(__file__).resolve().parents[1] / "spart-python_b200" / "spart_b200" / "data"

OPT_KEYS = ["nr", "Kab", "Kca", "Ks", "Kw", "Kdm", "Kant", "cbc", "prot", "nw", "GSV"]


def main():
    warnings.simplefilter("ignore")
    with open(REF / "model_parameters/optical_params.pkl", "rb") as f:
        op = pickle.load(f)
    with open(REF / "model_parameters/ET_irradiance.pkl", "rb") as f:
        et = pickle.load(f)
    out = {k: np.ascontiguousarray(op[k]) for k in OPT_KEYS}
    out["wl"] = np.ascontiguousarray(op["wl"])
    out["Ea"] = np.ascontiguousarray(et["Ea"])
    out["wl_Ea"] = np.ascontiguousarray(et["wl_Ea"])
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / "optical.npz", **out)
    print("optical.npz", {k: (v.shape, str(v.dtype)) for k, v in out.items()})

    (OUT / "sensors").mkdir(exist_ok=True)
    for pkl in sorted((REF / "sensor_information").glob("*.pkl")):
        with open(pkl, "rb") as f:
            s = pickle.load(f)
        d = {}
        for k, v in s["SMAC_coef"].items():
            d["SMAC_coef." + k] = np.ascontiguousarray(v)
        d["wl_smac"] = np.ascontiguousarray(s["wl_smac"])
        d["wl_srf_smac"] = np.ascontiguousarray(s["wl_srf_smac"])
        d["p_srf_smac"] = np.ascontiguousarray(s["p_srf_smac"])
        d["band_id_smac"] = np.array([str(b) for b in s["band_id_smac"]], dtype=np.str_)
        np.savez_compressed(OUT / "sensors" / (pkl.stem + ".npz"), **d)
        print(pkl.stem, d["wl_smac"].shape, str(d["wl_smac"].dtype), str(d["SMAC_coef.wo"].dtype))


if __name__ == "__main__":
    sys.exit(main())
