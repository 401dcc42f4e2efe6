"""Per-stage entry points with the reference's names and return types, evaluated on the GPU:

    PROSPECT_5D(leafbio, optical_params=None) -> LeafOptics     reference prospect_5d.py:117-246
    BSM(soilpar, optical_params=None)         -> SoilOptics     reference bsm.py:17-59
    SAILH(soil, leafopt, canopy, angles)      -> CanopyReflectances   reference sailh.py:14-237
    set_leaf_refl_trans_assumptions / set_soil_refl_trans_assumptions  reference SPART.py:427-470

plus the batched forms prospect_batch / bsm_batch / sailh_batch, so that the reference's unit-test
grids (tests/unit/test_PROSPECT.py, tests/unit/test_SAILH.py) run against the CUDA path.
`optical_params` is accepted for signature compatibility and ignored: the tables live on the GPU.
"""
from dataclasses import dataclass

import numpy as np
import torch

from .engine import NWL_S, default_engine

NWL = 2001
# finite stand-ins for the parameter rows a stage does not use
_DUMMY = np.array([40, 0.01, 0.02, 0, 10, 10, 1.5, 0, 0, 0.5, 0, 100, 20, 25, 0.015, 3, -0.35, -0.15, 0.05,
                   40, 0, 0, 0.3, 0.35, 1.4, 1013.25, 100], dtype=np.float64)


@dataclass
class LeafOptics:
    refl: np.ndarray
    tran: np.ndarray
    kChlrel: np.ndarray


class SoilOptics:
    def __init__(self, refl, refl_dry):
        self.refl = refl
        self.refl_dry = refl_dry


class CanopyReflectances:
    def __init__(self, rso, rdo, rsd, rdd):
        self.rso = rso
        self.rdo = rdo
        self.rsd = rsd
        self.rdd = rdd


def _block(n, **rows):
    """[27, n] parameter block: dummy defaults with the given row ranges overwritten."""
    P = np.repeat(_DUMMY[:, None], n, axis=1)
    for (lo, hi), v in rows.values():
        P[lo:hi] = np.asarray(v, dtype=np.float64).T
    return P


def prospect_batch(leaf):
    """leaf [n, 9] (or [n, 7]) -> refl, tran, kChlrel, each [n, 2001] (NumPy)."""
    leaf = np.atleast_2d(np.asarray(leaf, dtype=np.float64))
    if leaf.shape[1] == 7:
        leaf = np.concatenate([leaf, np.zeros((leaf.shape[0], 2))], axis=1)
    eng = default_engine()
    P = torch.from_numpy(_block(leaf.shape[0], leaf=((0, 9), leaf))).to(eng.device)
    s = eng.forward_spectrum(P)[:, 0:3, :NWL].cpu().numpy()
    return s[:, 0], s[:, 1], s[:, 2]


def bsm_batch(soil, soil_spectrum=None):
    """soil [n, 6] (B lat lon SMp SMC film) -> wet, dry soil reflectance, each [n, 2001]."""
    soil = np.atleast_2d(np.asarray(soil, dtype=np.float64))
    eng = default_engine()
    P = torch.from_numpy(_block(soil.shape[0], soil=((9, 15), soil))).to(eng.device)
    s = eng.forward_spectrum(P, soil_spectrum=soil_spectrum)[:, 3:5, :NWL].cpu().numpy()
    return s[:, 0], s[:, 1]


def sailh_batch(soil_refl, leaf_refl, leaf_tran, canopy, angles):
    """Spectra [2162] (shared) or [n, 2162]; canopy [n, 4]; angles [n, 3] -> [n, 4, 2162]."""
    canopy = np.atleast_2d(np.asarray(canopy, dtype=np.float64))
    angles = np.atleast_2d(np.asarray(angles, dtype=np.float64))
    n = canopy.shape[0]
    eng = default_engine()
    P = torch.from_numpy(_block(n, canopy=((15, 19), canopy), angles=((19, 22), angles))).to(eng.device)
    spec = []
    for x in (soil_refl, leaf_refl, leaf_tran):
        a = np.asarray(x, dtype=np.float64)
        a = a.reshape(-1) if a.size == NWL_S else a.reshape(n, -1)
        if a.shape[-1] != NWL_S:
            raise RuntimeError("Parameter leafopt.refl must be of len 2162 i.e. include thermal specturm. \n This error"
                               " usually occurs if you are feeding the prospect_5d output directly into the SAILH model"
                               " with adding\n the neccessary thermal wavelengths.")
        spec.append(torch.from_numpy(np.ascontiguousarray(a)).to(eng.device))
    if len({t.dim() for t in spec}) != 1:
        spec = [t if t.dim() == 2 else t.expand(n, NWL_S).contiguous() for t in spec]
    return eng.sailh(P, *spec).cpu().numpy()


ATM_FIELDS = ("Ta_s", "Ta_o", "Tg", "Ra_dd", "Ra_so", "Ta_ss", "Ta_sd", "Ta_oo", "Ta_do")


class AtmosphericOptics:
    """Atmospheric reflectance / transmittance arrays of SMAC, each [1, nb] (reference smac.py:216-272)."""

    def __init__(self, Ta_s, Ta_o, Tg, Ra_dd, Ra_so, Ta_ss, Ta_sd, Ta_oo, Ta_do):
        self.Ta_s, self.Ta_o, self.Tg, self.Ra_dd, self.Ra_so = Ta_s, Ta_o, Tg, Ra_dd, Ra_so
        self.Ta_ss, self.Ta_sd, self.Ta_oo, self.Ta_do = Ta_ss, Ta_sd, Ta_oo, Ta_do


def smac_batch(angles, atm, sensor):
    """angles [n, 3], atm [n, 4] (aot550 uo3 uh2o Pa) -> [n, 9, nb] in the order of ATM_FIELDS.
    `sensor` is a shipped sensor name or a sensorinfo dict (its 'SMAC_coef' are the coefficients)."""
    angles = np.atleast_2d(np.asarray(angles, dtype=np.float64))
    atm = np.atleast_2d(np.asarray(atm, dtype=np.float64))
    eng = default_engine()
    P = torch.from_numpy(_block(angles.shape[0], angles=((19, 22), angles), atm=((22, 26), atm))).to(eng.device)
    return eng.smac(P, sensor).cpu().numpy()


def SMAC(angles, atm, sensor):
    """Reference-shaped SMAC(angles, atm, coefs) (smac.py:14-213).  The third argument is the sensor
    name or the sensorinfo dict the coefficients belong to (the folded coefficient tables live on the
    GPU per sensor, so a bare coefficient dict is not accepted)."""
    out = smac_batch([angles.as_row()], [atm.as_row()], sensor)[0]
    return AtmosphericOptics(*[out[i][None, :] for i in range(9)])


def PROSPECT_5D(leafbio, optical_params=None):
    if (leafbio.PROT > 0.0 or leafbio.CBC > 0.0) and leafbio.Cdm > 0:
        print("WARNING: When setting PROT and/or CBC > 0. we\nassume that PROSPECT-PRO was called. Cdm will be\n"
              "therefore set to zero (Cdm = PROT + CBC)")
    refl, tran, kchl = prospect_batch([leafbio.as_row()])
    return LeafOptics(refl[0][:, None], tran[0][:, None], kchl[0][:, None])


def BSM(soilpar, optical_params=None):
    spectrum = np.asarray(soilpar.rdry, dtype=np.float64).reshape(-1) if getattr(soilpar, "rdry_set", False) else None
    wet, dry = bsm_batch([soilpar.as_row()], soil_spectrum=spectrum)
    return SoilOptics(wet[0][:, None], dry[0][:, None])


def set_soil_refl_trans_assumptions(soilopt, spectral=None):
    """Thermal soil reflectance = value at 2400 nm (reference SPART.py:427-442)."""
    r = np.asarray(soilopt.refl, dtype=np.float64).reshape(-1)
    soilopt.refl = np.concatenate([r, np.full(NWL_S - NWL, r[-1])])[:, None]
    return soilopt


def set_leaf_refl_trans_assumptions(leafopt, leafbio, spectral=None):
    """Thermal leaf reflectance / transmittance = rho_thermal / tau_thermal (reference SPART.py:445-470)."""
    leafopt.refl = np.concatenate([np.asarray(leafopt.refl).reshape(-1), np.full(NWL_S - NWL, leafbio.rho_thermal)])[:, None]
    leafopt.tran = np.concatenate([np.asarray(leafopt.tran).reshape(-1), np.full(NWL_S - NWL, leafbio.tau_thermal)])[:, None]
    return leafopt


def SAILH(soil, leafopt, canopy, angles):
    if len(leafopt.refl) != NWL_S:
        raise RuntimeError("Parameter leafopt.refl must be of len 2162 i.e. include thermal specturm. \n This error"
                           " usually occurs if you are feeding the prospect_5d output directly into the SAILH model"
                           " with adding\n the neccessary thermal wavelengths.")
    out = sailh_batch(np.asarray(soil.refl).reshape(-1), np.asarray(leafopt.refl).reshape(-1),
                      np.asarray(leafopt.tran).reshape(-1), [canopy.as_row()], [angles.as_row()])[0]
    return CanopyReflectances(out[0][:, None], out[1][:, None], out[2][:, None], out[3][:, None])
