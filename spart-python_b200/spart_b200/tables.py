"""Host-side preparation of the immutable tables uploaded to the GPU once per context.

Everything here is sample-independent set-up work (a few thousand NumPy operations per
context); no per-sample arithmetic ever runs on the host.

* `leaf_soil_constants()` builds the [17, 2001] per-wavelength table of
  `include/spart_b200.h::SpartTables` from the optical tables of the reference
  (reference src/SPART/model_parameters/optical_params.pkl, exported verbatim to
  data/optical.npz by tools/export_tables.py).
* `SensorTables` holds what the hot path needs of one sensor_information/<sensor>.pkl:
  np.interp knots for the band centres (SPART.py:216-223), host-folded SMAC constants
  (smac.py:44-92, 149-167) and the SRF-convolved extraterrestrial irradiance
  (SPART.py:358-396).  Coefficients keep the dtype they have in the reference's pickle while
  the sample-independent sub-expressions are folded, because the float32 Sentinel-2
  coefficients make NumPy evaluate exactly those sub-expressions in float32
  (SURVEY.md section 8(a), row a11).
"""
from dataclasses import dataclass
from pathlib import Path

import numpy as np

DATA_DIR = Path(__file__).resolve().parent / "data"

NWL = 2001
NLC = 17
NSMAC = 60

SENSOR_NAMES = (
    "TerraAqua-MODIS", "LANDSAT4-TM", "LANDSAT5-TM", "LANDSAT7-ETM", "LANDSAT8-OLI",
    "Sentinel2A-MSI", "Sentinel2B-MSI", "Sentinel3A-OLCI", "Sentinel3B-OLCI",
)

# rows of SpartSensor.smac -- keep in sync with enum SmacRow in csrc/spart_device.cuh
SMAC_ROWS = (
    "ah2o", "nh2o", "ao3", "no3",
    "ao2", "no2", "npo2", "aco2", "nco2", "npco2", "ach4", "nch4", "npch4",
    "ano2", "nno2", "npno2", "aco", "nco", "npco",
    "a0s", "a1s", "a2s", "a3s", "a0T", "a1T", "a2T", "a3T",
    "taur", "a0taup", "a1taup",
    "wo", "ak2", "ak", "opb", "omb", "opb2", "omb2", "ww", "g3", "d3", "h3",
    "a0P", "a1P", "a2P", "a3P", "a4P",
    "Rest1", "Rest2", "Rest3", "Rest4", "Resr1", "Resr2", "Resr3",
    "Resa1", "Resa2", "Resa3", "Resa4",
    "resr2taur", "akd3",
)


def load_optical(data_dir=DATA_DIR):
    with np.load(Path(data_dir) / "optical.npz") as z:
        return {k: z[k] for k in z.files}


def interface_transmissivity(alpha_deg, n):
    """Average transmissivity of a dielectric interface for isotropic incidence within a
    cone of half-angle `alpha_deg` (Stern 1964 / Allen 1973), as evaluated by
    calculate_tav (reference prospect_5d.py:249-311).  `n` may be an array."""
    n = np.asarray(n, dtype=np.float64)
    s2 = np.sin(alpha_deg * (np.pi / 180)) ** 2
    n2 = n ** 2
    npl = n2 + 1
    nmi = n2 - 1
    a = (n + 1) * (n + 1) / 2
    k = -(n2 - 1) * (n2 - 1) / 4
    half = s2 - npl / 2
    root = 0 if alpha_deg == 90 else np.sqrt(half * half + k)
    b = root - half
    # s-polarisation and p-polarisation parts
    ts = (k ** 2 / (6 * b ** 3) + k / b - b / 2) - (k ** 2 / (6 * a ** 3) + k / a - a / 2)
    qb = 2 * npl * b - nmi ** 2
    qa = 2 * npl * a - nmi ** 2
    tp = (-2 * n2 * (b - a) / (npl ** 2)
          + -2 * n2 * npl * np.log(b / a) / (nmi ** 2)
          + n2 * (1 / b - 1 / a) / 2
          + 16 * n2 ** 2 * (n2 ** 2 + 1) * np.log(qb / qa) / (npl ** 3 * nmi ** 2)
          + 16 * n2 ** 3 * (1 / qb - 1 / qa) / npl ** 3)
    return (ts + tp) / (2 * s2)


def leaf_soil_constants(opt=None):
    """[NLC, NWL] float64 table, rows as documented in include/spart_b200.h."""
    opt = opt or load_optical()
    col = lambda k: np.asarray(opt[k], dtype=np.float64).reshape(NWL)
    nr, nw = col("nr"), col("nw")
    lc = np.empty((NLC, NWL), dtype=np.float64)
    for i, k in enumerate(("Kab", "Kca", "Kdm", "Kw", "Ks", "Kant", "cbc", "prot")):
        lc[i] = col(k)
    t12 = interface_transmissivity(90, nr)
    lc[8] = interface_transmissivity(40, nr)                     # prospect_5d.py:200
    lc[9] = t12                                                  # prospect_5d.py:202
    lc[10] = t12 / (nr ** 2)                                     # prospect_5d.py:204
    gsv = np.asarray(opt["GSV"], dtype=np.float64)
    lc[11], lc[12], lc[13] = gsv[:, 0], gsv[:, 1], gsv[:, 2]
    lc[14] = interface_transmissivity(90, 2 / nw) / interface_transmissivity(90, 2)   # bsm.py:111
    lc[15] = 1 - interface_transmissivity(90, nw) / nw ** 2      # bsm.py:115
    lc[16] = 1 - interface_transmissivity(40, nw)                # bsm.py:119
    return np.ascontiguousarray(lc)


def nearest_index(values, grid):
    """Index of the nearest `grid` entry for each value, defined exactly as
    argmin(|grid - v|) over the whole grid (get_closest_index, SPART.py:381-387): first
    minimum on ties, 0 for NaN, and 0 for the uninitialised ~1e306 padding values of the
    Sentinel-2 SRF tables (all distances round to the same number).  Set-up time only."""
    grid = np.asarray(grid, dtype=np.float64).reshape(-1)
    v = np.asarray(values, dtype=np.float64)
    flat = v.reshape(-1)
    out = np.empty(flat.shape[0], dtype=np.int64)
    with np.errstate(invalid="ignore", over="ignore"):
        for s0 in range(0, flat.shape[0], 1024):
            d = np.abs(grid[:, None] - flat[None, s0:s0 + 1024])
            out[s0:s0 + 1024] = np.argmin(d, axis=0)
    return out.reshape(v.shape)


@dataclass
class SensorTables:
    name: str
    n_bands: int
    wl_smac: np.ndarray        # [nb] band-centre wavelengths in the pickle's dtype (DataFrame index)
    band_id: list              # [nb] str (DataFrame 'Band' column)
    wl_lo: np.ndarray          # [nb] int32
    wl_hi: np.ndarray          # [nb] int32
    wl_frac: np.ndarray        # [nb] float64
    smac: np.ndarray           # [NSMAC, nb] float64
    conv_ea: np.ndarray        # [nb] float64
    srf_index: np.ndarray      # [n_srf, nb] nearest 1-nm index of every SRF sample
    srf_weight: np.ndarray     # [n_srf, nb]
    srf_len: np.ndarray        # [nb] int32 number of non-zero SRF weights on the 1-nm grid
    srf_idx: np.ndarray        # [sum(srf_len)] int32 wavelength indices
    srf_w: np.ndarray          # [sum(srf_len)] normalised weights


def fold_smac(coef):
    """[NSMAC, nb] float64 from the 49 SMAC coefficient arrays ([1, nb], native dtype)."""
    c = {k: np.asarray(v) for k, v in coef.items()}
    nb = c["wo"].shape[1]
    f64 = lambda x: np.asarray(x, dtype=np.float64).reshape(nb)
    wo, gc = c["wo"], c["gc"]
    # sample-independent sub-expressions of smac.py:149-167 in the coefficients' own dtype
    ak2 = (1 - wo) * (3 - wo * 3 * gc)
    ak = np.sqrt(ak2)
    b = 2 * ak / (3 - wo * 3 * gc)
    folded = {
        "ak2": ak2, "ak": ak, "opb": 1 + b, "omb": 1 - b, "opb2": (1 + b) ** 2, "omb2": (1 - b) ** 2,
        "ww": wo / 4, "g3": wo * 3 * gc, "d3": 3 - wo * 3 * gc, "h3": (1 - wo) * 3 * gc,
        "resr2taur": c["Resr2"] * c["taur"],          # smac.py:184 evaluates Resr2 * taur first
        "akd3": f64(ak) / f64(3 - wo * 3 * gc),       # cp1 = c1 * ak / (3 - wo*3*gc), smac.py:166
    }
    # u**n with u = Peq**p * m is evaluated on the device as exp(n*p*ln Peq + n*ln m)
    for gas in ("o2", "co2", "ch4", "no2", "co"):
        folded["np" + gas] = f64(c["n" + gas]) * f64(c["p" + gas])
    out = np.zeros((NSMAC, nb), dtype=np.float64)
    for i, key in enumerate(SMAC_ROWS):
        out[i] = f64(folded[key] if key in folded else c[key])
    return out


def build_sensor(name, info, opt=None):
    """SensorTables from a sensorinfo mapping with the reference's keys
    ('SMAC_coef', 'wl_smac', 'wl_srf_smac', 'p_srf_smac', 'band_id_smac')."""
    opt = opt or load_optical()
    wl_smac = np.asarray(info["wl_smac"]).reshape(-1)
    x = wl_smac.astype(np.float64)
    nb = x.shape[0]
    if np.any(x < 400) or np.any(x > 2400):
        raise ValueError(f"sensor {name!r}: band centres must lie in 400..2400 nm")
    klo = np.floor(x).astype(np.int64) - 400
    frac = x - (klo + 400)
    khi = np.where(frac == 0, klo, klo + 1)
    idx = nearest_index(info["wl_srf_smac"], opt["wl_Ea"])
    p = np.asarray(info["p_srf_smac"], dtype=np.float64)
    ea = np.asarray(opt["Ea"], dtype=np.float64).reshape(-1)
    with np.errstate(all="ignore"):
        conv_ea = np.sum(ea[idx] * p, axis=0) / np.sum(p, axis=0)
    # dense SRF weights on the 1-nm grid: W[l, b] = sum of p over the SRF samples whose nearest
    # wavelength is l, divided by sum(p) (what calculate_spectral_convolution evaluates)
    W = np.zeros((NWL, nb))
    with np.errstate(all="ignore"):
        pn = p / np.sum(p, axis=0)[None, :]
    for b in range(nb):
        np.add.at(W[:, b], idx[:, b], np.nan_to_num(pn[:, b], nan=0.0))
    W[np.abs(W) < 1e-200] = 0.0      # uninitialised ~1e-310 padding weights of the Sentinel-2 tables: no effect
    ln = np.zeros(nb, dtype=np.int32)
    ichunks, wchunks = [], []
    for b in range(nb):
        nz = np.nonzero(W[:, b])[0]
        ln[b] = nz.size
        ichunks.append(nz.astype(np.int32))
        wchunks.append(W[nz, b])
    srf_idx = np.ascontiguousarray(np.concatenate(ichunks)) if ichunks else np.zeros(0, dtype=np.int32)
    srf_w = np.ascontiguousarray(np.concatenate(wchunks)) if wchunks else np.zeros(0)
    return SensorTables(
        srf_len=ln, srf_idx=srf_idx, srf_w=srf_w,
        name=name, n_bands=nb, wl_smac=wl_smac, band_id=[str(b) for b in info["band_id_smac"]],
        wl_lo=klo.astype(np.int32), wl_hi=khi.astype(np.int32), wl_frac=np.ascontiguousarray(frac),
        smac=fold_smac(info["SMAC_coef"]), conv_ea=np.ascontiguousarray(conv_ea),
        srf_index=idx, srf_weight=p,
    )


def load_sensor_info(name, data_dir=DATA_DIR):
    """The reference's sensorinfo dict for a shipped sensor.  Unknown names raise
    FileNotFoundError like the reference's open() (SPART.py:421-422)."""
    path = Path(data_dir) / "sensors" / f"{name}.npz"
    if not path.exists():
        raise FileNotFoundError(f"[Errno 2] No such file or directory: '{path}'")
    with np.load(path) as z:
        coef = {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith("SMAC_coef.")}
        return {
            "SMAC_coef": coef,
            "wl_smac": z["wl_smac"],
            "wl_srf_smac": z["wl_srf_smac"],
            "p_srf_smac": z["p_srf_smac"],
            "band_id_smac": [str(b) for b in z["band_id_smac"]],
        }


def synthetic_fullspectrum_sensorinfo(data_dir=DATA_DIR):
    """2001-band synthetic sensor (band centres 400..2400 nm) used for full 1-nm
    R_TOC/R_TOA/L_TOA output: SMAC coefficients linearly interpolated in wavelength from
    TerraAqua-MODIS, single-wavelength top-hat SRF.  The
    returned dict can be assigned to `SPART.sensorinfo` of the reference unchanged."""
    modis = load_sensor_info("TerraAqua-MODIS", data_dir)
    wl_m = modis["wl_smac"].T[0].astype(np.float64)
    order = np.argsort(wl_m)
    grid = np.arange(400, 2401, 1).astype(np.float64)
    coef = {k: np.interp(grid, wl_m[order], v[0].astype(np.float64)[order])[None, :]
            for k, v in modis["SMAC_coef"].items()}
    return {
        "SMAC_coef": coef,
        "wl_smac": grid[:, None],
        "wl_srf_smac": grid[None, :].copy(),
        "p_srf_smac": np.ones((1, NWL)),
        "band_id_smac": [f"{int(w)} nm" for w in grid],
    }
