#!/usr/bin/env python
"""Generate csrc/math_coeffs.h: polynomial kernels for the bounded-range sine/cosine used by
the leaf-angle iteration (sailh.py:379 is evaluated ~300 times per sample) and the
Gauss-Legendre rule of the hot-spot integral.

sin r = r + r^3 * PS(r^2),  cos r = 1 - r^2/2 + r^4 * PC(r^2)   on |r| <= pi/4,
PS / PC are Chebyshev interpolants in z = r^2 computed in 50-digit arithmetic and converted
to the monomial basis; their truncation error is below 2^-58 (checked below).
"""
from pathlib import Path

import mpmath as mp
import numpy as np

mp.mp.dps = 50
OUT = Path(__file__).resolve().parents[1] / "spart-python_b200" / "csrc" / "math_coeffs.h"


def cheb_monomial(f, a, b, deg):
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * i + 1) / (2 * n)) for i in range(n)]
    fx = [f((a + b) / 2 + (b - a) / 2 * t) for t in nodes]
    c = [2 * sum(fx[i] * mp.cos(mp.pi * j * (2 * i + 1) / (2 * n)) for i in range(n)) / n for j in range(n)]
    c[0] /= 2
    # Chebyshev (in t) -> monomial in t -> monomial in z = mid + half * t
    T = [[mp.mpf(0)] * n for _ in range(n)]
    T[0][0] = mp.mpf(1)
    if n > 1:
        T[1][1] = mp.mpf(1)
    for k in range(2, n):
        for j in range(n):
            T[k][j] = (2 * T[k - 1][j - 1] if j > 0 else 0) - T[k - 2][j]
    mono_t = [sum(c[k] * T[k][j] for k in range(n)) for j in range(n)]
    mid, half = (a + b) / 2, (b - a) / 2
    # substitute t = (z - mid) / half
    out = [mp.mpf(0)] * n
    for j, cj in enumerate(mono_t):
        # (z - mid)^j / half^j
        for i in range(j + 1):
            out[i] += cj * mp.binomial(j, i) * (-mid) ** (j - i) / half ** j
    return out


def PS(z):
    z = mp.mpf(z)
    if z == 0:
        return -mp.mpf(1) / 6
    r = mp.sqrt(z)
    return (mp.sin(r) - r) / (r * z)


def PC(z):
    z = mp.mpf(z)
    if z == 0:
        return mp.mpf(1) / 24
    r = mp.sqrt(z)
    return (mp.cos(r) - 1 + z / 2) / (z * z)


def main():
    zmax = (mp.pi / 4 + mp.mpf("0.01")) ** 2
    ps = cheb_monomial(PS, mp.mpf(0), zmax, 6)
    pc = cheb_monomial(PC, mp.mpf(0), zmax, 6)
    # float64 check
    S = [float(v) for v in ps]
    Cc = [float(v) for v in pc]
    worst_s = worst_c = 0.0
    for r in np.linspace(-0.79, 0.79, 4001):
        z = r * r
        p = 0.0
        for cf in S[::-1]:
            p = p * z + cf
        s = r + (r * z) * p
        p = 0.0
        for cf in Cc[::-1]:
            p = p * z + cf
        c = (1.0 - 0.5 * z) + (z * z) * p
        if r != 0:
            worst_s = max(worst_s, abs(float((mp.mpf(s) - mp.sin(r)) / mp.sin(r))))
        worst_c = max(worst_c, abs(float((mp.mpf(c) - mp.cos(r)) / mp.cos(r))))
    print("max rel err sin %.2e cos %.2e" % (worst_s, worst_c))
    assert worst_s < 3e-16 and worst_c < 3e-16
    x, w = np.polynomial.legendre.leggauss(12)
    xq = [mp.nstr(mp.mpf(float(v)), 20) for v in x]
    mp_x, mp_w = [], []
    for i in range(12):   # refine the nodes in mp: Newton on P_12
        xi = mp.mpf(float(x[i]))
        for _ in range(4):
            xi = xi - mp.legendre(12, xi) / mp.diff(lambda t: mp.legendre(12, t), xi)
        mp_x.append(xi)
        d = mp.diff(lambda t: mp.legendre(12, t), xi)
        mp_w.append(2 / ((1 - xi ** 2) * d ** 2))
    assert abs(sum(mp_w) - 2) < mp.mpf(10) ** -30

    def gauss_legendre(nq):
        xs, ws = [], []
        x0, _ = np.polynomial.legendre.leggauss(nq)
        for i in range(nq):
            xi = mp.mpf(float(x0[i]))
            for _ in range(4):
                xi = xi - mp.legendre(nq, xi) / mp.diff(lambda t: mp.legendre(nq, t), xi)
            d = mp.diff(lambda t: mp.legendre(nq, t), xi)
            xs.append(xi)
            ws.append(2 / ((1 - xi ** 2) * d ** 2))
        assert abs(sum(ws) - 2) < mp.mpf(10) ** -30
        return xs, ws
    gl24_x, gl24_w = gauss_legendre(24)
    gl16_x, gl16_w = gauss_legendre(16)
    gl10_x, gl10_w = gauss_legendre(10)
    gl4_x, gl4_w = gauss_legendre(4)
    gl6_x, gl6_w = gauss_legendre(6)
    # e^r on |r| <= ln2/2 (+1%): degree-11 Chebyshev interpolant, monomial basis
    rmax = mp.log(2) / 2 * mp.mpf("1.01")
    pe = cheb_monomial(lambda r: mp.exp(r), -rmax, rmax, 11)
    E = [float(v) for v in pe]
    worst_e = 0.0
    for r in np.linspace(-0.3466, 0.3466, 4001):
        p_ = 0.0
        for cf in E[::-1]:
            p_ = p_ * r + cf
        worst_e = max(worst_e, abs(float((mp.mpf(p_) - mp.exp(r)) / mp.exp(r))))
    print("max rel err exp kernel %.2e" % worst_e)
    assert worst_e < 3e-16
    # table-driven variant: e^x = 2^k * 2^(j/64) * e^r, |r| <= ln2/128; 2^(j/64) from a 64-entry table,
    # e^r as a degree-5 interpolant
    rmax_t = mp.log(2) / 128 * mp.mpf("1.01")
    def QE(r):
        r = mp.mpf(r)
        if r == 0:
            return mp.mpf(1) / 2
        return (mp.exp(r) - 1 - r) / (r * r)
    pet = cheb_monomial(QE, -rmax_t, rmax_t, 3)      # e^r = 1 + r + r^2 QE(r)
    tab = [mp.power(2, mp.mpf(j) / 64) for j in range(64)]
    ET = [float(v) for v in pet]
    TAB = [float(v) for v in tab]
    l64 = mp.log(2) / 64
    l64_hi = mp.mpf(float(l64))
    l64_lo = mp.mpf(float(l64 - l64_hi))
    inv_l64 = float(64 / mp.log(2))
    worst_t = 0.0
    rng = np.random.default_rng(0)

    def fma(a, b, c):
        return float(mp.mpf(a) * mp.mpf(b) + mp.mpf(c))
    for x_ in np.concatenate([rng.uniform(-700, 700, 3000), rng.uniform(-3, 3, 3000)]):
        x_ = float(x_)
        n_ = float(np.rint(x_ * inv_l64))
        r_ = fma(-n_, float(l64_hi), x_)
        r_ = fma(-n_, float(l64_lo), r_)
        r2 = r_ * r_
        q_ = fma(fma(ET[3], r2, ET[1]), r_, fma(ET[2], r2, ET[0]))
        t_ = TAB[int(n_) & 63]
        val = fma(t_, fma(r2, q_, r_), t_)
        ex = mp.exp(mp.mpf(x_)) / mp.power(2, int(n_) >> 6)
        worst_t = max(worst_t, abs(float((mp.mpf(val) - ex) / ex)))
    print("max rel err table exp %.2e" % worst_t)
    assert worst_t < 2.3e-16
    # log m = 2 s (1 + z PL(z)), s = (m-1)/(m+1), z = s^2, m in [sqrt(1/2), sqrt(2)]
    smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1) * mp.mpf("1.01")

    def PL(z):
        z = mp.mpf(z)
        if z == 0:
            return mp.mpf(1) / 3
        s_ = mp.sqrt(z)
        return (mp.atanh(s_) / s_ - 1) / z
    pl = cheb_monomial(PL, mp.mpf(0), smax ** 2, 8)
    Lc = [float(v) for v in pl]
    worst_l = 0.0
    for m_ in np.linspace(0.7072, 1.4142, 4001):
        s_ = (m_ - 1) / (m_ + 1)
        z = s_ * s_
        p_ = 0.0
        for cf in Lc[::-1]:
            p_ = p_ * z + cf
        val = 2 * s_ + 2 * s_ * z * p_
        ex = mp.log(mp.mpf(m_))
        if abs(m_ - 1) > 1e-3:
            worst_l = max(worst_l, abs(float((mp.mpf(val) - ex) / ex)))
    print("max rel err log kernel %.2e" % worst_l)
    assert worst_l < 4e-16
    ln2 = mp.log(2)
    ln2_hi = mp.mpf(float(ln2))
    ln2_lo = mp.mpf(float(ln2 - ln2_hi))
    pio2 = mp.pi / 2
    hi = mp.mpf(float(pio2))
    lo = mp.mpf(float(pio2 - hi))
    with open(OUT, "w") as f:
        f.write("// GENERATED by tools/gen_math_coeffs.py -- do not edit.\n#pragma once\n")
        f.write("#define SPART_SIN_POLY {" + ", ".join(mp.nstr(v, 20) for v in ps) + "}\n")
        f.write("#define SPART_COS_POLY {" + ", ".join(mp.nstr(v, 20) for v in pc) + "}\n")
        f.write("#define SPART_PIO2_HI %s\n#define SPART_PIO2_LO %s\n" % (mp.nstr(hi, 20), mp.nstr(lo, 20)))
        f.write("#define SPART_TWO_OVER_PI %s\n" % mp.nstr(2 / mp.pi, 20))
        f.write("#define SPART_EXP_POLY {" + ", ".join(mp.nstr(v, 20) for v in pe) + "}\n")
        f.write("#define SPART_EXPT_POLY {" + ", ".join(mp.nstr(v, 20) for v in pet) + "}\n")
        f.write("#define SPART_EXP2_TABLE {" + ", ".join(mp.nstr(v, 20) for v in tab) + "}\n")
        f.write("#define SPART_64_OVER_LN2 %s\n#define SPART_LN2_64_HI %s\n#define SPART_LN2_64_LO %s\n" % (
            mp.nstr(64 / mp.log(2), 20), mp.nstr(l64_hi, 20), mp.nstr(l64_lo, 20)))
        f.write("#define SPART_LOG_POLY {" + ", ".join(mp.nstr(v, 20) for v in pl) + "}\n")
        f.write("#define SPART_LOG2E %s\n#define SPART_LN2_HI %s\n#define SPART_LN2_LO %s\n" % (
            mp.nstr(1 / ln2, 20), mp.nstr(ln2_hi, 20), mp.nstr(ln2_lo, 20)))
        f.write("#define SPART_GL12_X {" + ", ".join(mp.nstr(v, 20) for v in mp_x) + "}\n")
        f.write("#define SPART_GL12_W {" + ", ".join(mp.nstr(v, 20) for v in mp_w) + "}\n")
        f.write("#define SPART_GL6_X {" + ", ".join(mp.nstr(v, 20) for v in gl6_x) + "}\n")
        f.write("#define SPART_GL6_W {" + ", ".join(mp.nstr(v, 20) for v in gl6_w) + "}\n")
        f.write("#define SPART_GL4_X {" + ", ".join(mp.nstr(v, 20) for v in gl4_x) + "}\n")
        f.write("#define SPART_GL4_W {" + ", ".join(mp.nstr(v, 20) for v in gl4_w) + "}\n")
        f.write("#define SPART_GL10_X {" + ", ".join(mp.nstr(v, 20) for v in gl10_x) + "}\n")
        f.write("#define SPART_GL10_W {" + ", ".join(mp.nstr(v, 20) for v in gl10_w) + "}\n")
        f.write("#define SPART_GL16_X {" + ", ".join(mp.nstr(v, 20) for v in gl16_x) + "}\n")
        f.write("#define SPART_GL16_W {" + ", ".join(mp.nstr(v, 20) for v in gl16_w) + "}\n")
        f.write("#define SPART_GL24_X {" + ", ".join(mp.nstr(v, 20) for v in gl24_x) + "}\n")
        f.write("#define SPART_GL24_W {" + ", ".join(mp.nstr(v, 20) for v in gl24_w) + "}\n")
    print("wrote", OUT)


if __name__ == "__main__":
    import sys
    if "--out" in sys.argv:          # write somewhere else (tests compare with the committed header)
        OUT = Path(sys.argv[sys.argv.index("--out") + 1])
    main()
