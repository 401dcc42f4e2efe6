#!/usr/bin/env python
"""Generate csrc/tau_coeffs.h: piecewise polynomial tables for the PROSPECT plate
transmissivity  tau(K) = (1-K) e^-K + K^2 E1(K)   (reference prospect_5d.py:182-196).

The reference evaluates E1 with one QUADPACK call per wavelength.  The CUDA path uses
closed forms with no cancellation instead:

  K <  1 :  E1(K) = -gamma - ln K + K * P(u),   u = 2K - 1          (interval 0)
  K >= 1 :  tau   = e^-K * t * H_i(u),  t = 1/K, binade i = 1..7 of K,
            H(t) = (K e^K E1(K) - 1 + t) / t^2  = 2 - 6 t + 24 t^2 - ...
            u = (t - mid_i) / half_i maps the binade's t-range to [-1, 1].

All tables are degree-DEG polynomials in the monomial basis of u (Horner), obtained from
Chebyshev interpolation in 60-digit arithmetic (mpmath), so they are accurate to the
rounding of the coefficients (~1e-16 relative on tau; checked below against mpmath).
"""
import sys
from pathlib import Path

import mpmath as mp
import numpy as np
from numpy.polynomial import chebyshev as C

mp.mp.dps = 60
DEG = 16
DEGF = 7      # FP32 mode: Chebyshev interpolants of degree 7 are within 6.2e-8 (float epsilon) on every interval
OUT = Path(__file__).resolve().parents[1] / "spart-python_b200" / "csrc" / "tau_coeffs.h"

BINADES = [(1, 2), (2, 4), (4, 8), (8, 16), (16, 32), (32, 64), (64, None)]


def P_exact(x):
    """P(x) = sum_{n>=1} (-1)^(n+1) x^(n-1) / (n n!)  so that E1 = -gamma - ln x + x P(x)."""
    x = mp.mpf(x)
    s = mp.mpf(0)
    for n in range(1, 60):
        s += (-1) ** (n + 1) * x ** (n - 1) / (n * mp.factorial(n))
    return s


def H_exact(t):
    t = mp.mpf(t)
    if t == 0:
        return mp.mpf(2)
    x = 1 / t
    if x > 200:   # asymptotic series is exact to 60 digits long before it diverges
        s = mp.mpf(0)
        for n in range(2, 120):
            term = (-1) ** n * mp.factorial(n) * t ** (n - 2)
            s += term
            if abs(term) < mp.mpf(10) ** -55:
                break
        return s
    return (x * mp.exp(x) * mp.e1(x) - 1 + t) / t ** 2


def cheb_coeffs(f, a, b, deg):
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * i + 1) / (2 * n)) for i in range(n)]
    fx = [f((a + b) / 2 + (b - a) / 2 * t) for t in nodes]
    c = []
    for j in range(n):
        s = sum(fx[i] * mp.cos(mp.pi * j * (2 * i + 1) / (2 * n)) for i in range(n))
        c.append(2 * s / n)
    c[0] /= 2
    return c


def cheb_to_monomial(c):
    """Exact (mp) conversion of a Chebyshev series to monomial coefficients."""
    n = len(c)
    T = [[mp.mpf(0)] * n for _ in range(n)]
    T[0][0] = mp.mpf(1)
    if n > 1:
        T[1][1] = mp.mpf(1)
    for k in range(2, n):
        for j in range(n):
            T[k][j] = (2 * T[k - 1][j - 1] if j > 0 else 0) - T[k - 2][j]
    return [sum(c[k] * T[k][j] for k in range(n)) for j in range(n)]


def tau_exact(K):
    K = mp.mpf(K)
    return (1 - K) * mp.exp(-K) + K ** 2 * mp.e1(K)


def main():
    rows, mids, halfs = [], [], []
    rows.append(cheb_to_monomial(cheb_coeffs(P_exact, mp.mpf(0), mp.mpf(1), DEG)))
    mids.append(mp.mpf(0.5)); halfs.append(mp.mpf(0.5))
    for lo, hi in BINADES:
        a = mp.mpf(0) if hi is None else mp.mpf(1) / hi
        b = mp.mpf(1) / lo
        rows.append(cheb_to_monomial(cheb_coeffs(H_exact, a, b, DEG)))
        mids.append((a + b) / 2); halfs.append((b - a) / 2)

    rows_f = [cheb_to_monomial(cheb_coeffs(P_exact, mp.mpf(0), mp.mpf(1), DEGF))]
    for lo, hi in BINADES:
        a = mp.mpf(0) if hi is None else mp.mpf(1) / hi
        rows_f.append(cheb_to_monomial(cheb_coeffs(H_exact, a, mp.mpf(1) / lo, DEGF)))
    worst_f = 0.0
    for i, (f, (a, b)) in enumerate(zip([P_exact] + [H_exact] * len(BINADES),
                                        [(mp.mpf(0), mp.mpf(1))] + [(mp.mpf(0) if hi is None else mp.mpf(1) / hi,
                                                                     mp.mpf(1) / lo) for lo, hi in BINADES])):
        for u in np.linspace(-1, 1, 41):
            x = (a + b) / 2 + (b - a) / 2 * mp.mpf(float(u))
            p = sum(mp.mpf(float(np.float32(float(c)))) * mp.mpf(float(u)) ** k for k, c in enumerate(rows_f[i]))
            worst_f = max(worst_f, abs(float((p - f(x)) / f(x))))
    print("max relative error of the degree-%d float tables (float-rounded coefficients): %.3e" % (DEGF, worst_f))
    assert worst_f < 1.5e-7

    # verification in float64 arithmetic (same operation order as the device code)
    tab = np.array([[float(v) for v in r] for r in rows])
    mid = np.array([float(v) for v in mids]); inv_half = np.array([float(1 / v) for v in halfs])
    worst = 0.0
    rng = np.random.default_rng(0)
    xs = np.concatenate([10 ** rng.uniform(-6, 0, 400), rng.uniform(1, 2, 300), rng.uniform(2, 64, 600),
                         rng.uniform(64, 700, 200), [1.0, 2.0, 4.0, 64.0, 0.999999, 1e-8]])
    for K in xs:
        if K < 1:
            u = 2 * K - 1
            p = 0.0
            for cf in tab[0][::-1]:
                p = p * u + cf
            e1 = -0.57721566490153286 - np.log(K) + K * p
            tau = (1 - K) * np.exp(-K) + K * K * e1
        else:
            i = min(int(np.floor(np.log2(K))) + 1, 7)
            t = 1.0 / K
            u = (t - mid[i]) * inv_half[i]
            p = 0.0
            for cf in tab[i][::-1]:
                p = p * u + cf
            tau = np.exp(-K) * t * p
        ex = tau_exact(K)
        if ex != 0:
            worst = max(worst, abs(float((mp.mpf(tau) - ex) / ex)))
    print("max relative error of tau(K) in float64 evaluation: %.3e" % worst)
    assert worst < 2e-15

    with open(OUT, "w") as f:
        f.write("// GENERATED by tools/gen_tau_coeffs.py -- do not edit.\n")
        f.write("// Piecewise degree-%d polynomials for the PROSPECT plate transmissivity tau(K).\n" % DEG)
        f.write("#pragma once\n#define SPART_TAU_DEG %d\n#define SPART_TAU_NINT %d\n" % (DEG, len(rows)))
        f.write("static const double SPART_TAU_COEF_H[SPART_TAU_NINT][SPART_TAU_DEG + 1] = {\n")
        for r in rows:
            f.write("  {" + ", ".join(mp.nstr(v, 20, min_fixed=0, max_fixed=0) for v in r) + "},\n")
        f.write("};\n")
        f.write("// FP32 mode: degree-%d interpolants on the same intervals (float coefficients)\n" % DEGF)
        f.write("#define SPART_TAUF_DEG %d\n" % DEGF)
        f.write("static const float SPART_TAUF_COEF_H[SPART_TAU_NINT][SPART_TAUF_DEG + 1] = {\n")
        for r in rows_f:
            f.write("  {" + ", ".join(mp.nstr(v, 10, min_fixed=0, max_fixed=0) + "f" for v in r) + "},\n")
        f.write("};\n")
        f.write("static const double SPART_TAU_MID_H[SPART_TAU_NINT] = {" + ", ".join(mp.nstr(v, 20) for v in mids) + "};\n")
        f.write("static const double SPART_TAU_INVHALF_H[SPART_TAU_NINT] = {" + ", ".join(mp.nstr(1 / v, 20) for v in halfs) + "};\n")
    print("wrote", OUT)


if __name__ == "__main__":
    if "--out" in sys.argv:          # write somewhere else (tests compare with the committed header)
        OUT = Path(sys.argv[sys.argv.index("--out") + 1])
    sys.exit(main())
