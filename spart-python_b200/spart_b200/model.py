"""`SPART` -- the reference's single-run class on top of the CUDA path.

Mirrors reference src/SPART/SPART.py:35-269: same constructor, same mutable attributes
(soilpar, leafbio, canopy, atm, angles, sensor, DOY, sensorinfo), `run(debug=False)` returning
the same DataFrame, and the result attributes R_TOC / R_TOA / L_TOA ([1, nb]) plus
soilopt / leafopt / canopyopt.  Every `run()` is a fresh evaluation of all stages on the GPU:
the reference's dirty-flag cache (and its stale-SMAC bug, SPART.py:186-187 vs 226) is not
reproduced.
"""
from types import SimpleNamespace

import numpy as np
import pandas as pd
import torch

from . import tables as T
from .engine import _sensor_digest, default_engine
from .params import pack_params


class SpectralBands:
    """Wavelength grids (reference SPART.py:272-315)."""

    def __init__(self):
        self.wlP = np.arange(400, 2401, 1)
        self.wlE = np.arange(400, 751, 1)
        self.WlF = np.arange(640, 851, 1)
        self.wlO = np.arange(400, 2401, 1)
        self.wlT = np.concatenate([np.arange(2500, 15001, 100), np.arange(16000, 50001, 1000)])
        self.wlS = np.concatenate([self.wlO, self.wlT])
        self.wlPAR = np.arange(400, 701, 1)
        self.nwlP = len(self.wlP)
        self.nwlT = len(self.wlT)
        self.IwlP = np.arange(0, self.nwlP, 1)
        self.IwlT = np.arange(self.nwlP, self.nwlP + self.nwlT, 1)


def load_optical_parameters():
    return T.load_optical()


def load_sensor_info(sensor):
    return T.load_sensor_info(sensor)


class SPART:
    def __init__(self, soilpar, leafbio, canopy, atm, angles, sensor, DOY):
        self.soilpar = soilpar
        self.leafbio = leafbio
        self.canopy = canopy
        self.atm = atm
        self.angles = angles
        self.sensor = sensor
        self.DOY = DOY
        self.spectral = SpectralBands()
        self.sensorinfo = load_sensor_info(sensor)       # FileNotFoundError for unknown sensors
        self._loaded_sensor = sensor
        self._shipped_info = self.sensorinfo
        self._shipped_digest = _sensor_digest(self.sensorinfo)
        self._spec = None

    def _params(self):
        return pack_params(self.soilpar, self.leafbio, self.canopy, self.atm, self.angles, self.DOY)

    def _soil_spectrum(self):
        if getattr(self.soilpar, "rdry_set", False):
            return np.asarray(self.soilpar.rdry, dtype=np.float64).reshape(-1)
        return None

    def _sensor_key(self):
        # like the reference, a changed `sensor` attribute does not reload sensorinfo; a user
        # may however assign or edit the sensorinfo dict (SPART.py:95 is a plain attribute): every
        # array the hot path uses is compared with the shipped table, an edited dict is a custom sensor
        # (the engine keys those by content)
        if self.sensorinfo is self._shipped_info:
            same = self._shipped_digest == _sensor_digest(self.sensorinfo)
        else:
            same = False
        return self._loaded_sensor if same else self.sensorinfo

    def run(self, debug=False):
        eng = default_engine()
        leaf = self.leafbio
        if (leaf.PROT > 0.0 or leaf.CBC > 0.0) and leaf.Cdm > 0:
            print("WARNING: When setting PROT and/or CBC > 0. we\n"
                  "assume that PROSPECT-PRO was called. Cdm will be\n"
                  "therefore set to zero (Cdm = PROT + CBC)")
        key = self._sensor_key()
        p = self._params()
        soil = self._soil_spectrum()
        if getattr(self.canopy, "lidf_set", False):      # an assigned leaf inclination distribution: device path
            dev = torch.from_numpy(p).to(eng.device)
            out = eng.forward_bands(dev, key, soil_spectrum=soil, lidf=self.canopy.lidf[:, 0]).cpu().numpy()[0]
        else:
            out = eng.forward_bands_host(p, key, soil_spectrum=soil)[0]            # [nb, 3]
        self.R_TOC = out[:, 0][None, :].copy()
        self.R_TOA = out[:, 1][None, :].copy()
        self.L_TOA = out[:, 2][None, :].copy()
        self._spec = None
        _, st = eng.sensor(key, soil)          # the context that just ran (cached)
        table = pd.DataFrame(zip(st.band_id, self.L_TOA[0], self.R_TOA[0], self.R_TOC[0]),
                             index=st.wl_smac, columns=["Band", "L_TOA", "R_TOA", "R_TOC"])
        if debug:
            table["rsoil"] = np.interp(st.wl_smac, self.spectral.wlS, self.soilopt.refl[:, 0])
        return table

    # ---- leafopt / soilopt / canopyopt (SPART.py:192-214), evaluated on demand ----------
    def _spectra(self):
        if self._spec is None:
            eng = default_engine()
            p = torch.from_numpy(self._params()).to(eng.device)
            lidf = self.canopy.lidf[:, 0] if getattr(self.canopy, "lidf_set", False) else None
            self._spec = eng.forward_spectrum(p, soil_spectrum=self._soil_spectrum(),
                                              rho_thermal=getattr(self.leafbio, "rho_thermal", 0.01),
                                              tau_thermal=getattr(self.leafbio, "tau_thermal", 0.01),
                                              lidf=lidf)[0].cpu().numpy()
        return self._spec

    @property
    def leafopt(self):
        s = self._spectra()
        return SimpleNamespace(refl=s[0][:, None], tran=s[1][:, None], kChlrel=s[2][:2001, None])

    @property
    def soilopt(self):
        s = self._spectra()
        return SimpleNamespace(refl=s[3][:, None], refl_dry=s[4][:2001, None])

    @property
    def canopyopt(self):
        s = self._spectra()
        return SimpleNamespace(rso=s[5][:, None], rdo=s[6][:, None], rsd=s[7][:, None], rdd=s[8][:, None])
