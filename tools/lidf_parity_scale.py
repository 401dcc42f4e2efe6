#!/usr/bin/env python
"""Leaf-angle parity at benchmark scale: the reference's truncated fixed-point iteration
(sailh.py:368-384, restated in oracle/spart_oracle.py::leafangles) against the GPU's
exact-steps + Taylor-model reproduction, for n (LIDFa, LIDFb) pairs of the bench distribution.
A step-count mismatch in one of the 12 n iterations shows up as a difference of ~1e-9..1e-8.
usage: python tools/lidf_parity_scale.py [n]   (run on the GPU box)"""
import json
import os
import sys
from multiprocessing import Pool
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "spart-python_b200"))
sys.path.insert(0, str(ROOT / "oracle"))


def _oracle_chunk(ab):
    import spart_oracle as so
    return so.leafangles(ab[:, 0], ab[:, 1])


def main(n):
    import torch

    import spart_b200
    rng = np.random.default_rng(20261018)
    ab = rng.uniform(-0.5, 0.5, size=(n, 2))            # LIDFa, LIDFb ~ U(-0.5, 0.5)
    eng = spart_b200.default_engine(torch.device("cuda", 0))
    got = eng.leafangles(ab)
    cores = len(os.sched_getaffinity(0))
    chunks = np.array_split(ab, max(cores * 4, 1))
    with Pool(cores) as pool:
        want = np.concatenate(pool.map(_oracle_chunk, chunks))
    d = np.abs(got - want)
    print(json.dumps({
        "n_samples": n, "n_iterations": 12 * n, "distribution": "LIDFa, LIDFb ~ U(-0.5, 0.5)",
        "max_abs_diff_lidf": float(d.max()),
        "count_gt": {t: int((d > float(t)).sum()) for t in ("1e-13", "1e-12", "1e-11", "1e-10", "1e-9")},
        "samples_with_any_diff_gt_1e-12": int((d > 1e-12).any(axis=1).sum()),
    }))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000)
