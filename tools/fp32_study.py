#!/usr/bin/env python
"""CPU study of the FP32-mode leaf + canopy arithmetic (planning tool, not product code).

Emulates prospect_point_f / sailh_point_f of csrc/spart_device_f32.cuh in NumPy float32 (IEEE
single precision like the GPU's FFMA-free evaluation; the SFU approximations add a few ulp on top) on
the config-3 distribution and reports the relative error of the four canopy reflectances against the
float64 oracle for several variants:
  base      the round-1 formulas
  v1        1 - rinf^2 formed as 2 m rinf / sigb, 1 - e2 / 1 - tau e1 by expm1-style forms
  v2        v1 + leaf absorptance 1 - rho - tau carried from a float64 Stokes solve
usage: python tools/fp32_study.py [n]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import spart_oracle as so  # noqa: E402

f32 = np.float32


def stokes(tau, t_alph, t12, t21, N, dt):
    """prospect_5d.py:208-241 from the plate transmissivity, in dtype dt. Returns refl, tran, absorptance."""
    one = dt(1)
    tau, t_alph, t12, t21, N = (np.asarray(x, dtype=dt) for x in (tau, t_alph, t12, t21, N))
    r_alph, r12, r21 = one - t_alph, one - t12, one - t21
    tt21 = tau * t21
    inv_d1 = one / (one - r21 * r21 * tau * tau)
    Ta = t_alph * tt21 * inv_d1
    Ra = r_alph + r21 * tau * Ta
    t = t12 * tt21 * inv_d1
    r = r12 + r21 * tau * t
    Nm1 = N - one
    with np.errstate(all="ignore"):
        D = np.sqrt((one + r + t) * (one + r - t) * (one - r + t) * (one - r - t))
        rq, tq = r * r, t * t
        a = (one + rq - tq + D) / (dt(2) * r)
        b = (one - rq + tq + D) / (dt(2) * t)
        bNm1 = np.exp(Nm1 * np.log(b)).astype(dt)
        bN2 = bNm1 * bNm1
        a2 = a * a
        inv_d2 = one / (a2 * bN2 - one)
        Rsub = a * (bN2 - one) * inv_d2
        Tsub = bNm1 * (a2 - one) * inv_d2
        j = (r + t) >= one
        Tsub0 = t / (t + (one - t) * Nm1)
        Tsub = np.where(j, Tsub0, Tsub)
        Rsub = np.where(j, one - Tsub0, Rsub)
    inv_d3 = one / (one - Rsub * r)
    tran = Ta * Tsub * inv_d3
    refl = Ra + Ta * Rsub * t * inv_d3
    return refl.astype(dt), tran.astype(dt), (one - refl - tran).astype(dt)


def sail(G, rho, tau, rs, dt, variant, absorb=None):
    one, half = dt(1), dt(0.5)
    c = lambda x: np.asarray(x, dtype=dt)
    k, K, bf, LAI, sob, sof = c(G["k"]), c(G["K"]), c(G["bf"]), c(G["LAI"]), c(G["sob"]), c(G["sof"])
    tau_ss, tau_oo, sumpso, pso2w, Z = c(G["tau_ss"]), c(G["tau_oo"]), c(G["sumpso"]), c(G["pso2w"]), c(G["Z"])
    rho, tau, rs = c(rho), c(tau), c(rs)
    sdb, sdf = half * (k + bf), half * (k - bf)
    ddb, ddf = half * (one + bf), half * (one - bf)
    dob, dof = half * (K + bf), half * (K - bf)
    sigb = ddb * rho + ddf * tau
    sigf = ddf * rho + ddb * tau
    sb = sdb * rho + sdf * tau
    sf = sdf * rho + sdb * tau
    vb = dob * rho + dof * tau
    vf = dof * rho + dob * tau
    w = sob * rho + sof * tau
    a = one - sigf
    ab = (one - rho - tau) if absorb is None else c(absorb)
    m = np.sqrt(ab * (a + sigb))
    rinf = (a - m) / sigb if variant in ("base", "v1") else sigb / (a + m)
    rinf2 = rinf * rinf
    e1 = np.exp(-m * LAI).astype(dt)
    e2 = e1 * e1
    inv_km, inv_Km = one / (k + m), one / (K + m)

    def J1(kk, ek):
        d = (kk - m) * LAI
        return np.where(np.abs(d) < dt(2e-2), half * (e1 + ek) * LAI * (one - dt(1 / 12) * d * d), (e1 - ek) / (kk - m))
    J1k, J1K = J1(k, tau_ss), J1(K, tau_oo)
    if variant == "base":
        omr2 = one - rinf2
        J2k = (one - tau_ss * e1) * inv_km
        J2K = (one - tau_oo * e1) * inv_Km
        ome2 = one - e2
    else:
        omr2 = dt(2) * m * rinf / sigb
        J2k = -np.expm1(-(k + m) * LAI).astype(dt) * inv_km
        J2K = -np.expm1(-(K + m) * LAI).astype(dt) * inv_Km
        ome2 = -np.expm1(-dt(2) * m * LAI).astype(dt)
    re = rinf * e1
    inv_den = one / (omr2 * (one + rinf2))
    s1, s2 = sf + rinf * sb, sf * rinf + sb
    v1, v2 = vf + rinf * vb, vf * rinf + vb
    Pss, Qss, Poo, Qoo = s1 * J1k, s2 * J2k, v1 * J1K, v2 * J2K
    tau_dd = omr2 * e1 * inv_den
    rho_dd = rinf * ome2 * inv_den
    tau_sd = (Pss - re * Qss) * inv_den
    tau_do = (Poo - re * Qoo) * inv_den
    rho_sd = (Qss - re * Pss) * inv_den
    rho_do = (Qoo - re * Poo) * inv_den
    T1 = v2 * s1 * (Z - J1k * tau_oo) * inv_Km + v1 * s2 * (Z - J1K * tau_ss) * inv_km
    T2 = -(Qoo * rho_sd + Poo * tau_sd) * rinf
    rho_sod = (T1 + T2) / omr2
    rho_so = rho_sod + w * sumpso
    rs_den = rs / (one - rs * rho_dd)
    rso = rho_so + rs * pso2w + ((tau_sd + tau_ss * rs * rho_dd) * tau_oo + (tau_sd + tau_ss) * tau_do) * rs_den
    rdo = rho_do + (tau_oo + tau_do) * tau_dd * rs_den
    rsd = rho_sd + (tau_ss + tau_sd) * tau_dd * rs_den
    rdd = rho_dd + tau_dd * tau_dd * rs_den
    return np.stack([rso, rdo, rsd, rdd], axis=-1).astype(np.float64), dict(T1=T1, T2=T2, omr2=omr2, m=m)


def main(n=20000):
    opt = so.load_optical()
    P = so.synthetic_params(n, 3, seed=2003)
    sensor = so.load_sensor("LANDSAT8-OLI")
    lo, hi, frac = so.band_sample_points(sensor["wl_smac"].T[0])
    idx = lo
    leaf = P[:, so.CAB:so.CBC + 1]
    refl, tran, _ = so.prospect(leaf, opt, idx)
    rwet, _ = so.bsm(P[:, so.SOIL_B:so.FILM + 1], opt, idx)
    geo = so.sail_geometry(P[:, so.LAI:so.HOT_Q + 1], P[:, so.SZA:so.RAA + 1])
    want = np.stack(so.sailh(rwet, refl, tran, None, None, geo=geo), axis=-1)
    LAI, k, K = geo["LAI"], geo["k"], geo["K"]
    G = dict(k=k, K=K, bf=geo["bf"], LAI=LAI, sob=geo["sob"], sof=geo["sof"], tau_ss=np.exp(-k * LAI),
             tau_oo=np.exp(-K * LAI), sumpso=geo["Pso"][:, :60].sum(1, keepdims=True) * LAI / 60,
             pso2w=geo["Pso"][:, 60:61], Z=(1 - np.exp(-k * LAI) * np.exp(-K * LAI)) / (K + k))
    valid = ((want > 0) & (want < 1)).all(axis=2)

    def report(name, got):
        with np.errstate(all="ignore"):
            e = np.abs(got - want) / np.abs(want)
        e = np.where(valid[..., None], e, 0)
        i = np.unravel_index(np.argmax(e), e.shape)
        print(f"{name:34s} max {e.max():.2e}  p99.99 {np.quantile(e, 0.9999):.2e}  >1e-4: {(e > 1e-4).sum():6d}"
              f"  worst at sample {i[0]} band {i[1]} out {i[2]}: rho+tau={refl[i[0], i[1]] + tran[i[0], i[1]]:.4f}"
              f" LAI={LAI[i[0], 0]:.2f}")

    # the plate transmissivity from the float64 oracle, rounded to float: isolates the Stokes + SAIL algebra
    nr = so._col(opt["nr"], idx)
    t_alph, t12 = so.calculate_tav(40, nr), so.calculate_tav(90, nr)
    t21 = t12 / nr ** 2
    Kall = ((leaf[:, 0:1] * so._col(opt["Kab"], idx) + leaf[:, 4:5] * so._col(opt["Kca"], idx)
             + np.where((leaf[:, 7:8] > 0) | (leaf[:, 8:9] > 0), 0, leaf[:, 1:2]) * so._col(opt["Kdm"], idx)
             + leaf[:, 2:3] * so._col(opt["Kw"], idx) + leaf[:, 3:4] * so._col(opt["Ks"], idx)
             + leaf[:, 5:6] * so._col(opt["Kant"], idx) + leaf[:, 8:9] * so._col(opt["cbc"], idx)
             + leaf[:, 7:8] * so._col(opt["prot"], idx)) / leaf[:, 6:7])
    from scipy.special import exp1
    tau64 = (1 - Kall) * np.exp(-Kall) + Kall ** 2 * exp1(Kall)
    N = leaf[:, 6:7]
    r64, t64, a64 = stokes(tau64, t_alph, t12, t21, N, np.float64)
    print("check stokes64 vs oracle:", np.abs(r64 - refl).max(), np.abs(t64 - tran).max())
    r32, t32, a32 = stokes(tau64.astype(f32), t_alph, t12, t21, N, f32)
    print("float32 Stokes: max rel err refl %.2e tran %.2e absorptance %.2e" % (
        np.abs(r32 - refl).max() / 1, (np.abs(t32 - tran) / tran).max(), (np.abs(a32 - a64) / np.abs(a64))[valid].max()))
    report("f64 leaf optics, f32 SAIL base", sail(G, refl, tran, rwet, f32, "base")[0])
    report("f64 leaf optics, f32 SAIL v1", sail(G, refl, tran, rwet, f32, "v1")[0])
    report("f64 leaf + absorptance, f32 SAIL v1", sail(G, refl, tran, rwet, f32, "v1", absorb=a64)[0])
    report("f32 Stokes, f32 SAIL base", sail(G, r32, t32, rwet, f32, "base")[0])
    report("f32 Stokes, f32 SAIL v1", sail(G, r32, t32, rwet, f32, "v1")[0])
    report("f32 Stokes+absorptance, f32 SAIL v1", sail(G, r32, t32, rwet, f32, "v1", absorb=a32)[0])
    report("f64 leaf optics, f32 SAIL v2", sail(G, refl, tran, rwet, f32, "v2")[0])
    report("f64 leaf + absorptance, f32 SAIL v2", sail(G, refl, tran, rwet, f32, "v2", absorb=a64)[0])
    report("f32 Stokes, f32 SAIL v2", sail(G, r32, t32, rwet, f32, "v2")[0])
    report("f32 Stokes+absorptance, f32 SAIL v2", sail(G, r32, t32, rwet, f32, "v2", absorb=a32)[0])
    # mixed: float64 Stokes where the single plate is near-conservative, float32 elsewhere, both from float32 tau
    tau32 = tau64.astype(f32)
    for thr in (0.0, 0.8, 0.9, 0.95):
        r64m, t64m, a64m = stokes(tau32.astype(np.float64), t_alph, t12, t21, N, np.float64)
        tt = t12 * tau64 * t21 / (1 - (1 - t21) ** 2 * tau64 ** 2)
        rr = (1 - t12) + (1 - t21) * tau64 * tt
        sel = (rr + tt) > thr
        rm = np.where(sel, r64m, r32.astype(np.float64)).astype(f32)
        tm = np.where(sel, t64m, t32.astype(np.float64)).astype(f32)
        am = np.where(sel, a64m, a32.astype(np.float64)).astype(f32)
        report(f"mixed Stokes thr {thr} ({sel.mean():.2f} f64), SAIL v2", sail(G, rm, tm, rwet, f32, "v2", absorb=am)[0])
    # tau itself perturbed by a float rounding of 1 - tau vs of tau
    omt32 = (1 - tau64).astype(f32).astype(np.float64)
    r64b, t64b, a64b = stokes(1 - omt32, t_alph, t12, t21, N, np.float64)
    report("f64 Stokes from f32(1-tau), f32 SAIL v1 +abs", sail(G, r64b, t64b, rwet, f32, "v1", absorb=a64b)[0])
    r64c, t64c, a64c = stokes(tau64.astype(f32).astype(np.float64), t_alph, t12, t21, N, np.float64)
    report("f64 Stokes from f32(tau), f32 SAIL v1 +abs", sail(G, r64c, t64c, rwet, f32, "v1", absorb=a64c)[0])
    report("f64 everything (sanity)", sail(G, refl, tran, rwet, np.float64, "base")[0])


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 20000)
