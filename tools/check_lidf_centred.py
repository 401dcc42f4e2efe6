#!/usr/bin/env python
"""NumPy prototype / check of the leaf-angle scheme of lidf_kernel (round 2): exact steps until the Newton estimate
of the distance to the fixed point is below R (1 - lam)^(1/3), then a degree-6 Taylor model of the map centred on
the Newton estimate of the fixed point (quantised to 2^-8 like the kernel's 8-bit offset), iterated with the
reference's stopping rule.  Prints the deviation of the resulting lidf from the step-by-step iteration of the oracle
(oracle/spart_oracle.py::leafangles, sailh.py:351-398) over the whole |a| + |b| <= 1 domain including its boundary,
and the step counts of both stages.
usage: python tools/check_lidf_centred.py [seed]"""
import math
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))
import spart_oracle as so  # noqa: E402


def lidf_centred(a, b, R, deg=6):
    n = a.size
    rd = np.pi / 180
    C = R ** 3 / 16
    thetas = [10 * i for i in range(1, 9)] + [82, 84, 86, 88]
    F = np.zeros((n, 14))
    itA = np.zeros(n)
    itB = np.zeros(n)
    for i, th in enumerate(thetas, 1):
        t2 = np.full(n, 2 * rd * th)
        x = t2.copy()
        y = np.zeros(n)
        stage = np.zeros(n, int)        # 0 = exact steps, 1 = handed over, 2 = converged on exact steps
        q = np.zeros(n)
        while (stage == 0).any():
            act = stage == 0
            s, c = np.sin(x), np.cos(x)
            ynew = s * (a + b * c)
            dx = 0.5 * (ynew - x + t2)
            yp = a * c + b * (2 * c * c - 1)
            adx = np.abs(dx)
            done = ~(adx > 1e-8)
            w = 1 - yp
            sw = (adx * adx * adx <= C * (w * w) * (w * w)) & ~done      # lidf2_hand
            ypf = yp.astype(np.float32)
            with np.errstate(all="ignore"):                               # lidf2_leave: centre - next iterate
                d = dx.astype(np.float32) * (1 + ypf) / (1 - ypf)
                qq = np.clip(np.rint(d * np.float32(256.0)), -127, 127)
            q = np.where(act & sw, qq, q)
            y = np.where(act, ynew, y)
            x = np.where(act, x + dx, x)
            itA += act
            stage = np.where(act & done, 2, np.where(act & sw, 1, stage))
        inB = stage == 1
        xc = x + q / 256.0
        u = x - xc
        s, c = np.sin(xc), np.cos(xc)
        s2, c2 = 2 * s * c, 2 * c * c - 1
        cyc1, cyc2 = [s, c, -s, -c], [s2, c2, -s2, -c2]
        co = [(a * cyc1[k % 4] + b * (2.0 ** (k - 1)) * cyc2[k % 4]) / math.factorial(k) for k in range(deg + 1)]
        k0 = t2 - xc
        act = inB.copy()
        yt = co[0].copy()
        while act.any():
            p = co[deg]
            for k in range(deg - 1, -1, -1):
                p = p * u + co[k]
            du = 0.5 * (p - u + k0)
            yt = np.where(act, p, yt)
            u = np.where(act, u + du, u)
            itB += act
            act &= np.abs(du) > 1e-8
        F[:, i] = (2 * np.where(inB, yt, y) + t2) / np.pi
    F[:, 13] = 1
    return np.diff(F, axis=1), itA, itB


def main(seed):
    rng = np.random.default_rng(seed)
    n = 100000
    a = rng.uniform(-1, 1, n)
    b = rng.uniform(-1, 1, n)
    k = np.abs(a) + np.abs(b) <= 1
    a, b = a[k], b[k]
    ne = 3000                               # the boundary |a| + |b| = 1 and a few special pairs
    ae = rng.uniform(-1, 1, ne)
    be = (1 - np.abs(ae)) * rng.choice([-1, 1], ne)
    a = np.concatenate([a, ae, np.array([1.0, 0.999999, -1.0, 0, 0, 0.0, -0.35])])
    b = np.concatenate([b, be, np.array([0, 0, 0, 1.0, -1.0, 0, -0.15])])
    ref = so.leafangles(a, b)
    bench = (np.abs(a) <= 0.5) & (np.abs(b) <= 0.5)
    for R in (0.10, 0.15, 0.20):
        lidf, ia, ib = lidf_centred(a, b, R)
        e = np.abs(lidf - ref)
        print("R %.2f: max |lidf - step-by-step| %.1e, samples above 3e-15: %d of %d; steps per task on the bench "
              "distribution: %.2f exact (the first is free), %.2f polynomial; whole domain %.2f / %.2f, maxima %d / %d"
              % (R, e.max(), (e.max(1) > 3e-15).sum(), a.size, ia[bench].mean() / 12, ib[bench].mean() / 12,
                 ia.mean() / 12, ib.mean() / 12, ia.max(), ib.max()))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
