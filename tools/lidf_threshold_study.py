#!/usr/bin/env python
"""What a one-step difference in the leaf-angle iteration does to the outputs.

The reference stops dcum's fixed-point iteration at |dx| <= 1e-8 (sailh.py:374-383), so its result depends on the
step count; when |dx| of some step sits within the last bits of 1e-8 the count depends on the libm's sin.  This
tool constructs such threshold-straddling cases deliberately -- for a grid of (LIDFa, LIDFb, angle) it bisects
LIDFa until a step's |dx| equals 1e-8 to the last representable bit -- and evaluates the model on both sides of
the flip: the leaf inclination distribution with the step count N and with N + 1 for that one angle, pushed
through the oracle's SAILH + SMAC for random leaf / soil / atmosphere parameters.  Reported: the largest change
of any F value, and of any R_TOC / R_TOA / L_TOA output (relative).
usage: python tools/lidf_threshold_study.py [n_cases]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import spart_oracle as so  # noqa: E402

THETAS = [10, 20, 30, 40, 50, 60, 70, 80, 82, 84, 86, 88]


def dcum_trace(a, b, theta):
    """The reference's iteration for one angle: list of |dx| per step and of the y used at each step."""
    rd = np.pi / 180
    theta2 = 2 * rd * theta
    x = theta2
    out = []
    for _ in range(400):
        y = a * np.sin(x) + 0.5 * b * np.sin(2 * x)
        dx = 0.5 * (y - x + theta2)
        x = x + dx
        out.append((abs(dx), y))
        if abs(dx) <= 1e-8:
            break
    return out, theta2


def f_after(a, b, theta, steps):
    """F(theta) if the iteration is stopped after exactly `steps` steps."""
    rd = np.pi / 180
    theta2 = 2 * rd * theta
    x = theta2
    y = 0.0
    for _ in range(steps):
        y = a * np.sin(x) + 0.5 * b * np.sin(2 * x)
        x = x + 0.5 * (y - x + theta2)
    return (2 * y + theta2) / np.pi


def straddle(a0, b, theta):
    """Bisect LIDFa near a0 so that the stopping step's |dx| is as close to 1e-8 as float64 allows.
    Returns (a, N) with |dx_N(a)| <= 1e-8 < |dx_N(a + ulp)| (or the other way round), or None."""
    tr, _ = dcum_trace(a0, b, theta)
    N = len(tr)
    g = lambda a: dcum_dx(a, b, theta, N) - 1e-8
    lo, hi = a0, None
    for da in (1e-4, -1e-4, 1e-3, -1e-3, 1e-2, -1e-2):
        if abs(a0 + da) + abs(b) <= 1.0 and g(a0 + da) * g(a0) < 0:
            hi = a0 + da
            break
    if hi is None:
        return None
    glo = g(lo)
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if mid == lo or mid == hi:
            break
        gm = g(mid)
        if gm * glo > 0:
            lo, glo = mid, gm
        else:
            hi = mid
    return (lo, hi, N)


def dcum_dx(a, b, theta, N):
    rd = np.pi / 180
    theta2 = 2 * rd * theta
    x = theta2
    dx = 1.0
    for _ in range(N):
        y = a * np.sin(x) + 0.5 * b * np.sin(2 * x)
        dx = 0.5 * (y - x + theta2)
        x = x + dx
    return abs(dx)


def main(n_cases=60, seed=0):
    rng = np.random.default_rng(seed)
    worst_f, worst_out, used = 0.0, 0.0, 0
    P = so.synthetic_params(16, 3, seed=11)
    while used < n_cases:
        a0, b = rng.uniform(-0.5, 0.5, 2)
        ti = int(rng.integers(0, 12))
        theta = THETAS[ti]
        s = straddle(a0, b, theta)
        if s is None:
            continue
        lo, hi, N = s
        used += 1
        for a in (lo, hi):
            # the two admissible results for this angle: stop after N steps or after N + 1 (N - 1 on the other side)
            tr, _ = dcum_trace(a, b, theta)
            n_here = len(tr)
            lidf = so.leafangles(np.array([a]), np.array([b]))[0]
            F = np.concatenate([[0.0], np.cumsum(lidf)])
            # the flip this is about: |dx_N| is on the threshold, so the loop stops after N or after N + 1 steps
            for alt in ((N + 1,) if n_here == N else (N,) if n_here == N + 1 else ()):
                F2 = F.copy()
                F2[ti + 1] = f_after(a, b, theta, alt)
                worst_f = max(worst_f, abs(F2[ti + 1] - F[ti + 1]))
                lidf2 = np.diff(F2)
                Q = P.copy()
                Q[:, so.LIDFA], Q[:, so.LIDFB] = a, b
                outs = []
                for L in (lidf, lidf2):
                    outs.append(bands_with_lidf(Q, "LANDSAT8-OLI", np.repeat(L[None, :], Q.shape[0], 0)))
                worst_out = max(worst_out, float(np.max(np.abs(outs[1] - outs[0]) / np.abs(outs[0]))))
    return {"cases": used, "max_abs_change_of_F": worst_f, "max_rel_change_of_outputs": worst_out,
            "note": "one iteration step more or less on one of the 12 angles, LIDFa bisected onto the 1e-8 threshold"}


def bands_with_lidf(params, sensor, lidf):
    """spart_bands of the oracle with an imposed leaf inclination distribution [n, 13]."""
    opt = so.load_optical()
    sen = so.load_sensor(sensor)
    lo, hi, frac = so.band_sample_points(sen["wl_smac"].T[0])
    refl, tran, _ = so.prospect(params[:, so.CAB:so.CBC + 1], opt, lo)
    rwet, _ = so.bsm(params[:, so.SOIL_B:so.FILM + 1], opt, lo)
    geo = so.sail_geometry(params[:, so.LAI:so.HOT_Q + 1], params[:, so.SZA:so.RAA + 1], lidf=lidf)
    rso, rdo, rsd, rdd = so.sailh(rwet, refl, tran, None, None, geo=geo)
    atmo = so.smac(params[:, so.SZA:so.RAA + 1], params[:, so.AOT550:so.PA + 1], sen["SMAC_coef"])
    La = so.et_band_radiance(params[:, so.DOY], params[:, so.SZA], opt, sen)
    return np.stack(so.toc_to_toa(rso, rdo, rdd, rsd, atmo, La), axis=2)


if __name__ == "__main__":
    print(json.dumps(main(int(sys.argv[1]) if len(sys.argv) > 1 else 60), indent=1))
