"""Host-side driver of the CUDA path: one `Engine` per GPU.

PyTorch is used for device memory, streams and (in distributed.py) NCCL only; all
arithmetic happens in libspart_b200.so.  A context (immutable device tables) is created
lazily per sensor and cached.
"""
import threading

import numpy as np
import torch

from . import _lib
from . import tables as T

NPAR = 27
NOUT = 3
NSPEC = 9
NWL_S = 2162
MAX_SAMPLES_PER_CALL = 65535 * 128    # grid.y limit of band_kernel, see spart_forward_bands

_PRECISION = {"fp64": _lib.FP64, 64: _lib.FP64, "fp32": _lib.FP32, 32: _lib.FP32}


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.SpartError("spart_b200 needs a CUDA device (B200); there is no CPU fallback")


class Engine:
    """Owns the C contexts of one GPU."""

    def __init__(self, device=None):
        _require_cuda()
        self.lib = _lib.load()
        if self.lib.spart_device_count() <= 0:
            raise _lib.SpartError("libspart_b200: no CUDA device visible")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device.index if isinstance(device, torch.device) else int(device)))
        self._opt = T.load_optical()
        self._lc = T.leaf_soil_constants(self._opt)
        self._ctx = {}        # key -> (ctx handle, SensorTables, keep-alive arrays)
        self._lock = threading.Lock()

    # ---- contexts ------------------------------------------------------------------
    def sensor(self, sensor, soil_spectrum=None):
        """(ctx, SensorTables) for a shipped sensor name or a reference-style sensorinfo dict.
        soil_spectrum: optional user dry-soil reflectance [2001] (SoilParametersFromFile); such a
        context carries the spectrum as its first soil vector and must be run with
        SPART_FLAG_SOIL_SPECTRUM."""
        skey = sensor if isinstance(sensor, str) else ("custom", id(sensor))
        soil = None
        if soil_spectrum is not None:
            soil = np.ascontiguousarray(np.asarray(soil_spectrum, dtype=np.float64).reshape(-1))
            if soil.shape[0] != T.NWL:
                raise ValueError("soil_spectrum must have 2001 values (400..2400 nm)")
        key = (skey, None if soil is None else soil.tobytes())
        with self._lock:
            hit = self._ctx.get(key)
            if hit is not None:
                return hit[0], hit[1]
            info = T.load_sensor_info(sensor) if isinstance(sensor, str) else sensor
            st = T.build_sensor(sensor if isinstance(sensor, str) else "custom", info, self._opt)
            lc = self._lc
            if soil is not None:
                lc = self._lc.copy()
                lc[11], lc[12], lc[13] = soil, 0.0, 0.0
            tabs = _lib.SpartTables(n_wl=T.NWL, lc=_lib.as_double_ptr(lc))
            smac = np.ascontiguousarray(st.smac)
            cs = _lib.SpartSensor(n_bands=st.n_bands, wl_lo=_lib.as_int32_ptr(st.wl_lo),
                                  wl_hi=_lib.as_int32_ptr(st.wl_hi), wl_frac=_lib.as_double_ptr(st.wl_frac),
                                  smac=_lib.as_double_ptr(smac), conv_ea=_lib.as_double_ptr(st.conv_ea),
                                  srf_len=_lib.as_int32_ptr(st.srf_len), srf_idx=_lib.as_int32_ptr(st.srf_idx),
                                  srf_w=_lib.as_double_ptr(st.srf_w))
            handle = _lib.c_void_p()
            _lib.check(self.lib.spart_create(_lib.byref(tabs), _lib.byref(cs), 1, self.device.index,
                                             _lib.byref(handle)), "spart_create")
            self._ctx[key] = (handle, st, (smac, sensor, lc))
            return handle, st

    def close(self):
        with self._lock:
            for handle, _, _ in self._ctx.values():
                self.lib.spart_destroy(handle)
            self._ctx.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device path ---------------------------------------------------------------
    def _prep(self, params):
        if not (isinstance(params, torch.Tensor) and params.is_cuda and params.dtype == torch.float64
                and params.dim() == 2 and params.shape[0] == NPAR
                and (params.shape[1] <= 1 or params.stride(1) == 1)):
            raise ValueError("params must be a CUDA float64 tensor [27, n] with contiguous rows")
        if params.device != self.device:
            raise ValueError(f"params on {params.device}, engine on {self.device}")
        return params, params.shape[1], (params.stride(0) if params.shape[1] > 1 else max(params.shape[1], 1))

    def forward_bands(self, params, sensor, out=None, precision="fp64", uniform_geometry=False,
                      soil_spectrum=None, band_mode="interp"):
        """params: CUDA float64 [27, n] -> CUDA float64 [n, nb, 3] = (R_TOC, R_TOA, L_TOA).
        Asynchronous on the current torch stream.  uniform_geometry=True asserts that the three
        angle rows are constant over the batch (SPART_FLAG_UNIFORM_GEOMETRY).  soil_spectrum: dry
        soil reflectance [2001] used instead of the B/lat/lon soil vectors.  band_mode: "interp"
        (the reference: np.interp at the band centre) or "srf" (SRF-weighted band means)."""
        if band_mode not in ("interp", "srf"):
            raise ValueError("band_mode must be 'interp' or 'srf'")
        handle, st = self.sensor(sensor, soil_spectrum)
        params, n, ld = self._prep(params)
        if out is None:
            out = torch.empty((n, st.n_bands, NOUT), dtype=torch.float64, device=self.device)
        elif not (out.is_cuda and out.dtype == torch.float64 and out.is_contiguous()
                  and tuple(out.shape) == (n, st.n_bands, NOUT)):
            raise ValueError("out must be a contiguous CUDA float64 tensor [n, nb, 3]")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        prec = _PRECISION[precision]
        flags = (_lib.FLAG_UNIFORM_GEOMETRY if uniform_geometry else 0) | (
            _lib.FLAG_SOIL_SPECTRUM if soil_spectrum is not None else 0) | (
            _lib.FLAG_SRF_BANDS if band_mode == "srf" else 0)
        for s0 in range(0, n, MAX_SAMPLES_PER_CALL):
            m = min(MAX_SAMPLES_PER_CALL, n - s0)
            ws = torch.empty(self.lib.spart_workspace_bytes(handle, m) // 8, dtype=torch.float64, device=self.device)
            _lib.check(self.lib.spart_forward_bands(handle, 0, params.data_ptr() + 8 * s0, m, ld, prec, flags,
                                                    ws.data_ptr(), out.data_ptr() + 8 * s0 * st.n_bands * NOUT,
                                                    stream), "spart_forward_bands")
        return out

    def forward_bands_multi(self, params, sensors, outs=None, precision="fp64", uniform_geometry=False,
                            soil_spectrum=None, band_mode="interp"):
        """One batch evaluated for several sensors: the sensor-independent per-sample kernels run
        once and every further sensor only runs the band kernel (SPART_FLAG_REUSE_RECORD).
        Returns a list of CUDA float64 [n, nb_i, 3] tensors."""
        params, n, ld = self._prep(params)
        if n > MAX_SAMPLES_PER_CALL:
            raise ValueError(f"forward_bands_multi handles at most {MAX_SAMPLES_PER_CALL} samples per call")
        ctxs = [self.sensor(s, soil_spectrum) for s in sensors]
        if outs is None:
            outs = [torch.empty((n, st.n_bands, NOUT), dtype=torch.float64, device=self.device) for _, st in ctxs]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        base = (_lib.FLAG_UNIFORM_GEOMETRY if uniform_geometry else 0) | (
            _lib.FLAG_SOIL_SPECTRUM if soil_spectrum is not None else 0) | (
            _lib.FLAG_SRF_BANDS if band_mode == "srf" else 0)
        ws = torch.empty(self.lib.spart_workspace_bytes(ctxs[0][0], n) // 8, dtype=torch.float64, device=self.device)
        for i, ((handle, st), out) in enumerate(zip(ctxs, outs)):
            flags = base | (_lib.FLAG_REUSE_RECORD if i > 0 else 0)
            _lib.check(self.lib.spart_forward_bands(handle, 0, params.data_ptr(), n, ld, _PRECISION[precision], flags,
                                                    ws.data_ptr(), out.data_ptr(), stream), "spart_forward_bands")
        return outs

    def forward_spectrum(self, params, out=None, soil_spectrum=None):
        """params: CUDA float64 [27, n] -> CUDA float64 [n, 9, 2162]: leaf refl, leaf tran,
        kChlrel, soil refl, soil refl dry, rso, rdo, rsd, rdd."""
        handle, _ = self.sensor("Sentinel2A-MSI", soil_spectrum)   # any context carries the wavelength tables
        params, n, ld = self._prep(params)
        if out is None:
            out = torch.empty((n, NSPEC, NWL_S), dtype=torch.float64, device=self.device)
        ws = torch.empty(self.lib.spart_workspace_bytes(handle, n) // 8, dtype=torch.float64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        flags = _lib.FLAG_SOIL_SPECTRUM if soil_spectrum is not None else 0
        _lib.check(self.lib.spart_forward_spectrum(handle, params.data_ptr(), n, ld, flags, ws.data_ptr(),
                                                   out.data_ptr(), stream), "spart_forward_spectrum")
        return out

    def smac(self, params, sensor, out=None):
        """SMAC alone.  params: CUDA float64 [27, n] (angle and atmosphere rows used)
        -> CUDA float64 [n, 9, nb]: Ta_s, Ta_o, Tg, Ra_dd, Ra_so, Ta_ss, Ta_sd, Ta_oo, Ta_do."""
        handle, st = self.sensor(sensor)
        params, n, ld = self._prep(params)
        if out is None:
            out = torch.empty((n, 9, st.n_bands), dtype=torch.float64, device=self.device)
        ws = torch.empty(self.lib.spart_workspace_bytes(handle, n) // 8, dtype=torch.float64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.spart_smac(handle, 0, params.data_ptr(), n, ld, ws.data_ptr(), out.data_ptr(), stream),
                   "spart_smac")
        return out

    def sailh(self, params, soil_refl, leaf_refl, leaf_tran, out=None):
        """SAILH on caller-supplied spectra.  params: CUDA float64 [27, n] (canopy and angle rows
        used); spectra: CUDA float64 [2162] (shared by all samples) or [n, 2162].
        -> CUDA float64 [n, 4, 2162] = rso, rdo, rsd, rdd."""
        handle, _ = self.sensor("Sentinel2A-MSI")
        params, n, ld = self._prep(params)
        specs = [soil_refl, leaf_refl, leaf_tran]
        shared = all(x.dim() == 1 for x in specs)
        for x in specs:
            ok = x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and (
                tuple(x.shape) == (NWL_S,) if shared else tuple(x.shape) == (n, NWL_S))
            if not ok:
                raise RuntimeError("Parameter leafopt.refl must be of len 2162 i.e. include thermal specturm "
                                   "(CUDA float64, all three spectra [2162] or all [n, 2162])")
        if out is None:
            out = torch.empty((n, 4, NWL_S), dtype=torch.float64, device=self.device)
        ws = torch.empty(self.lib.spart_workspace_bytes(handle, n) // 8, dtype=torch.float64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.spart_sailh(handle, params.data_ptr(), n, ld, soil_refl.data_ptr(), leaf_refl.data_ptr(),
                                        leaf_tran.data_ptr(), 0 if shared else NWL_S, ws.data_ptr(), out.data_ptr(),
                                        stream), "spart_sailh")
        return out

    def leafangles(self, ab):
        """[n, 2] (LIDFa, LIDFb) host array -> [n, 13] lidf host array."""
        ab = np.ascontiguousarray(np.asarray(ab, dtype=np.float64).reshape(-1, 2).T)
        n = ab.shape[1]
        d = torch.from_numpy(ab).to(self.device)
        out = torch.empty((n, 13), dtype=torch.float64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.spart_leafangles(d.data_ptr(), n, n, out.data_ptr(), stream), "spart_leafangles")
        return out.cpu().numpy()

    # ---- host path -----------------------------------------------------------------
    def forward_bands_host(self, params, sensor, out=None, precision="fp64", uniform_geometry=False,
                           soil_spectrum=None, band_mode="interp"):
        """params: host float64 [27, n] (NumPy array or CPU tensor, ideally pinned) ->
        host float64 [n, nb, 3].  H2D, kernels and D2H are pipelined inside the C library."""
        handle, st = self.sensor(sensor, soil_spectrum)
        p = params.numpy() if isinstance(params, torch.Tensor) else np.asarray(params)
        if p.dtype != np.float64 or p.ndim != 2 or p.shape[0] != NPAR or (p.shape[1] > 1 and p.strides[1] != 8):
            raise ValueError("params must be a host float64 array [27, n] with contiguous rows")
        n = p.shape[1]
        ld = p.strides[0] // 8 if n > 1 else max(n, 1)
        if out is None:
            out = np.empty((n, st.n_bands, NOUT), dtype=np.float64)
        o = out.numpy() if isinstance(out, torch.Tensor) else out
        if o.dtype != np.float64 or not o.flags.c_contiguous or o.shape != (n, st.n_bands, NOUT):
            raise ValueError("out must be a C-contiguous host float64 array [n, nb, 3]")
        with torch.cuda.device(self.device):
            if band_mode not in ("interp", "srf"):
                raise ValueError("band_mode must be 'interp' or 'srf'")
            flags = (_lib.FLAG_UNIFORM_GEOMETRY if uniform_geometry else 0) | (
                _lib.FLAG_SOIL_SPECTRUM if soil_spectrum is not None else 0) | (
                _lib.FLAG_SRF_BANDS if band_mode == "srf" else 0)
            _lib.check(self.lib.spart_forward_bands_host(handle, 0, p.ctypes.data, n, ld, _PRECISION[precision],
                                                         flags, o.ctypes.data), "spart_forward_bands_host")
        return out

    def profile_enable(self, sensor, on=True):
        handle, _ = self.sensor(sensor)
        _lib.check(self.lib.spart_profile_enable(handle, 1 if on else 0), "spart_profile_enable")

    def profile_read(self, sensor):
        """Summed CUDA-event durations of the two kernels since the last read:
        {'lidf_ms', 'geometry_ms', 'band_ms', 'calls'}."""
        handle, _ = self.sensor(sensor)
        ms = (_lib.c_double * _lib.NKERNELS)()
        c = _lib.c_int64()
        _lib.check(self.lib.spart_profile_read(handle, ms, _lib.byref(c)), "spart_profile_read")
        return {"lidf_ms": ms[0], "geometry_ms": ms[1], "band_ms": ms[2], "calls": c.value}

    def measure_peaks(self):
        a, b = _lib.c_double(), _lib.c_double()
        _lib.check(self.lib.spart_measure_peaks(self.device.index, _lib.byref(a), _lib.byref(b)),
                   "spart_measure_peaks")
        return {"fp64_tflops": a.value, "fp32_tflops": b.value}

    def launch_count(self):
        return int(self.lib.spart_launch_count())


_default = {}
_default_lock = threading.Lock()


def default_engine(device=None):
    """Process-wide engine of a device (created on first use)."""
    _require_cuda()
    idx = torch.cuda.current_device() if device is None else (
        device.index if isinstance(device, torch.device) else int(device))
    with _default_lock:
        if idx not in _default:
            _default[idx] = Engine(idx)
        return _default[idx]
