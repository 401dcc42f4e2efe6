/*
 * spart_b200.h -- C ABI of libspart_b200.so, the B200 (sm_100a) implementation of the
 * SPART forward model (BSM soil -> PROSPECT-5D/PRO leaf -> SAILH canopy -> SMAC atmosphere
 * -> sensor bands) evaluated over large batches of parameter sets.
 *
 * The reference (wirrell/SPART-python) has no FFI layer: its boundary is the Python call
 * surface `SPART(soilpar, leafbio, canopy, atm, angles, sensor, DOY).run()`
 * (reference src/SPART/SPART.py:83-95, 162-269).  Each entry point below names the part of
 * that surface it replaces.  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, a negative SPART_E* code
 *     for an argument error, or a positive cudaError_t value for a CUDA failure;
 *     spart_last_error() returns a thread-local description of the last failure;
 *   - no exception ever crosses this boundary;
 *   - `*_dev` pointers are device pointers owned by the caller (e.g. PyTorch's caching
 *     allocator); the library never allocates per call on the device paths and all work is
 *     enqueued asynchronously on the caller's stream (a cudaStream_t passed as void*);
 *   - a context is immutable after spart_create and may be used concurrently from several
 *     host threads; create one context per GPU;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Batch parameter layout ("params"): struct-of-arrays, double, [SPART_NPAR][ld] with
 * ld >= n, row order
 *   0..8   leaf    Cab Cdm Cw Cs Cca Cant N PROT CBC   (LeafBiology, prospect_5d.py:73-81)
 *   9..14  soil    B lat lon SMp SMC film              (SoilParameters, bsm.py:269-287)
 *   15..18 canopy  LAI LIDFa LIDFb q                   (CanopyStructure, sailh.py:340-348)
 *   19..21 angles  sol_angle obs_angle rel_angle, deg  (Angles, sailh.py:298-301)
 *   22..25 atm     aot550 uo3 uh2o Pa                  (AtmosphericProperties, smac.py:307-317)
 *   26     DOY                                         (SPART.__init__, SPART.py:83)
 *
 * Broadcast rows: `broadcast_rows` is a bit mask over these 27 rows.  A set bit r says that row r
 * is constant over the batch (a look-up table with fixed geometry, fixed SMC / film, PROSPECT-5D
 * without PROT / CBC ...): only params[r * ld + 0] is ever read, the host path copies one element
 * of that row instead of n, and when bits 19, 20 and 21 are all set -- one sun / observer geometry
 * for the whole batch BY CONSTRUCTION -- the library evaluates the sample-independent volume
 * scattering terms (_volscatt, sailh.py:401-446) and the geometry-only sub-expressions of SMAC
 * (smac.py:125-201) once per thread block instead of once per sample.  Results agree with the
 * general path to a few ulp (sums are re-associated).  When bits 16 and 17 (LIDFa, LIDFb) are both
 * set the leaf inclination distribution (calculate_leafangles, sailh.py:351-398) is iterated once for
 * the whole batch (bit-identical results).  There is no unchecked "the caller promises" flag: a row
 * is either read per sample or read once.
 */
#ifndef SPART_B200_H
#define SPART_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPART_ABI_VERSION 5

#define SPART_NPAR 27       /* rows of a parameter batch                                   */
#define SPART_NWL 2001      /* 400..2400 nm, 1 nm (SpectralBands.wlP, SPART.py:303)        */
#define SPART_NWL_S 2162    /* + 161 thermal wavelengths (SpectralBands.wlS, SPART.py:310) */
#define SPART_NLC 17        /* per-wavelength constants, see SpartTables                   */
#define SPART_NSMAC 60      /* per-band host-folded SMAC constants, see SpartSensor        */
#define SPART_NOUT 3        /* R_TOC, R_TOA, L_TOA                                         */
#define SPART_NSPEC 9       /* planes of spart_forward_spectrum                            */
#define SPART_NKERNELS 3    /* kernels of spart_forward_bands: leaf angles, geometry, bands */

enum {
  SPART_OK = 0,
  SPART_EINVAL = -1,   /* bad argument (null pointer, negative size, unknown sensor index) */
  SPART_ENODEV = -2,   /* no usable CUDA device                                            */
  SPART_ENOMEM = -3    /* host allocation failed                                           */
};

enum { SPART_FP64 = 64, SPART_FP32 = 32 };

/* flags of spart_forward_bands (bit 0 was ABI 4's unchecked uniform-geometry promise; it is now
 * derived from broadcast_rows and the value 1 is rejected) */
enum {
  /* the context was created with a user-supplied dry-soil spectrum in row 11 of SpartTables.lc
   * (rows 12, 13 zero): B / lat / lon are ignored and rdry = that spectrum, as with the reference's
   * SoilParametersFromFile (bsm.py:42-43, 155-226) */
  SPART_FLAG_SOIL_SPECTRUM = 2,
  /* band values of the four canopy reflectances are SRF-weighted means over the band
   * (calculate_spectral_convolution, SPART.py:358-396, applied to canopyopt) instead of the
   * reference's np.interp sample at the band centre (SPART.py:216-223); SPART_FP64 only */
  SPART_FLAG_SRF_BANDS = 4,
  /* the workspace still holds the per-sample record of a previous spart_forward_bands call for
   * the same params / n / precision / soil flag (any sensor): skip the per-sample kernels and run
   * only the band kernel.  This is how one batch is evaluated for several sensors (e.g.
   * Sentinel-2A and -2B) while paying for the sensor-independent work once. */
  SPART_FLAG_REUSE_RECORD = 8,
  /* SPART_FP32 only: params are float [SPART_NPAR][ld] and the results are written as float, which
   * halves the bytes on every copy and on the final gather of a sharded run.  The model is
   * evaluated at the float-rounded inputs. */
  SPART_FLAG_F32_IO = 16,
  /* compact result: out = [n][n_bands][2] (R_TOC, R_TOA) followed by etscale[n], the per-sample
   * extraterrestrial scale cf(DOY) cos(sza) / pi (SPART.py:345-353).  L_TOA is then rebuilt by the
   * consumer bit for bit as (conv_ea[b] * etscale[s]) * R_TOA[s][b] -- the product the kernels
   * themselves form (SPART.py:252); conv_ea is the SpartSensor member.  Two thirds of the bytes. */
  SPART_FLAG_COMPACT_OUT = 32,
  /* SPART_FP64, device path: the workspace already holds a caller-supplied leaf inclination
   * distribution (spart_set_lidf on the same workspace and stream): LIDFa / LIDFb are ignored, the
   * leaf-angle kernel is skipped and the 13 values enter k, K, bf, sob, sof as they are -- what the
   * reference does with an assigned CanopyStructure.lidf (sailh.py:81-97 read canopy.lidf). */
  SPART_FLAG_USER_LIDF = 64
};

/* elements of the result buffer of spart_forward_bands for n samples, nb bands and these flags */
#define SPART_OUT_ELEMS(n, nb, flags) \
  (((flags) & SPART_FLAG_COMPACT_OUT) ? (size_t)(n) * (size_t)(nb) * 2 + (size_t)(n) : (size_t)(n) * (size_t)(nb) * 3)

typedef struct SpartCtx SpartCtx;

/* Sample-independent per-wavelength constants, host pointer, double [SPART_NLC][SPART_NWL]:
 *   0 Kab 1 Kca 2 Kdm 3 Kw 4 Ks 5 Kant 6 cbc 7 prot          optical_params.pkl (SPART.py:399-406)
 *   8 tav(40,nr)  9 tav(90,nr)  10 tav(90,nr)/nr^2            prospect_5d.py:200-205
 *   11..13 GSV[:,0..2]                                         bsm.py:45-52
 *   14 tav(90,2/nw)/tav(90,2)  15 1-tav(90,nw)/nw^2  16 1-tav(40,nw)   bsm.py:110-119
 * calculate_tav (prospect_5d.py:249-311) only ever receives table arguments, so these are
 * folded once on the host. */
typedef struct {
  int32_t n_wl;            /* must be SPART_NWL */
  const double* lc;        /* [SPART_NLC][n_wl] */
} SpartTables;

/* One sensor = the content of sensor_information/<name>.pkl the hot path uses
 * (SPART.py:216-232, 358-396), with sample-independent sub-expressions folded on the host. */
typedef struct {
  int32_t n_bands;
  const int32_t* wl_lo;    /* [n_bands] index into 400..2400 nm of the knot at/below wl_smac      */
  const int32_t* wl_hi;    /* [n_bands] upper knot (== wl_lo when wl_smac hits a knot exactly)     */
  const double* wl_frac;   /* [n_bands] wl_smac - knot(wl_lo): np.interp weight (SPART.py:220-223) */
  const double* smac;      /* [SPART_NSMAC][n_bands] folded SMAC coefficients (smac.py:44-92)      */
  const double* conv_ea;   /* [n_bands] SRF-convolved Ea (SPART.py:358-396 applied to ETpar['Ea']) */
  /* optional (all three NULL = not available): spectral response per band for
   * SPART_FLAG_SRF_BANDS, already reduced to the 1-nm grid with the nearest-index rule of
   * get_closest_index (SPART.py:381-387) and normalised by sum(p_srf): band b has srf_len[b]
   * non-zero weights; its (wavelength index, weight) pairs follow those of band b-1 in
   * srf_idx / srf_w */
  const int32_t* srf_len;  /* [n_bands] */
  const int32_t* srf_idx;  /* [sum(srf_len)] index into 400..2400 nm */
  const double* srf_w;     /* [sum(srf_len)] */
} SpartSensor;

/* ABI version of the loaded library (== SPART_ABI_VERSION of the header it was built from). */
int spart_abi_version(void);

/* Thread-local text of the last error returned on this thread ("" if none). */
const char* spart_last_error(void);

/* Number of CUDA devices visible (0 when there is no driver/GPU). Never fails. */
int spart_device_count(void);

/* Replaces SPART.__init__'s table loads (SPART.py:92-95): uploads the immutable tables and
 * all sensors to `device` once.  *out must be released with spart_destroy. */
int spart_create(const SpartTables* tables, const SpartSensor* sensors, int32_t n_sensors,
                 int32_t device, SpartCtx** out);
int spart_destroy(SpartCtx* ctx);

/* Device scratch the caller must provide to the *_dev entry points for a batch of n samples. */
size_t spart_workspace_bytes(const SpartCtx* ctx, int64_t n);

/* Replaces the per-sample loop over SPART(...).run() (SPART.py:162-269): for every sample
 * s < n and band b of sensor `sensor`, out_dev[(s * n_bands + b) * 3 + {0,1,2}] =
 * {R_TOC, R_TOA, L_TOA} (or the compact layout of SPART_FLAG_COMPACT_OUT).  params_dev:
 * [SPART_NPAR][ld] (see top) of double, or of float with SPART_FLAG_F32_IO; out_dev has the same
 * element type.  broadcast_rows: see top.  precision: SPART_FP64 or SPART_FP32 (arithmetic type of
 * the spectral / atmosphere stage).  Asynchronous on `stream`; makes the context's device current
 * for the duration of the call and restores the caller's. */
int spart_forward_bands(const SpartCtx* ctx, int32_t sensor, const void* params_dev,
                        int64_t n, int64_t ld, uint32_t broadcast_rows, int32_t precision,
                        int32_t flags, void* workspace_dev, void* out_dev, void* stream);

/* Same computation with HOST buffers: params_host [SPART_NPAR][ld] and out_host (layout and
 * element type as above, SPART_OUT_ELEMS elements) are ordinary host memory.  The parameters travel
 * in spans of up to 128 Ki samples (one 2-D copy per run of per-sample rows) into three device
 * buffers; each span is evaluated in place in chunks of up to 64 Ki samples whose results flow back
 * as they complete.  One stream per copy direction, one per output slot, tied by events, so both
 * copy engines and the SMs run side by side.  Pinned or registered buffers (cudaHostAlloc /
 * cudaHostRegister / torch pin_memory) are
 * DMA'd directly; pageable buffers (plain NumPy arrays) are staged through internal pinned
 * buffers by a small pool of copy threads (SPART_HOST_THREADS, default 12), because an asynchronous
 * copy on pageable memory degenerates to a synchronous single-threaded driver copy.  Broadcast
 * rows move one element.  Returns when out_host is complete; on failure no copy is left in
 * flight.  This is the drop-in for a caller that holds NumPy arrays. */
int spart_forward_bands_host(SpartCtx* ctx, int32_t sensor, const void* params_host,
                             int64_t n, int64_t ld, uint32_t broadcast_rows, int32_t precision,
                             int32_t flags, void* out_host);

/* Stores a caller-supplied leaf inclination distribution, lidf_dev [n][13] of double (the 13 classes
 * of sailh.py:49, each row summing to 1), in the workspace of a batch of n samples; the following
 * spart_forward_bands call on that workspace passes SPART_FLAG_USER_LIDF.  Asynchronous on stream. */
int spart_set_lidf(const double* lidf_dev, int64_t n, void* workspace_dev, void* stream);

/* Replaces the leafopt / soilopt / canopyopt attributes of a SPART object after run()
 * (SPART.py:192-214, 427-470): full 2162-wavelength spectra.
 * out_dev: double [n][SPART_NSPEC][SPART_NWL_S], planes
 *   0 leaf refl  1 leaf tran  2 kChlrel (0 beyond 2400 nm)  3 soil refl (wet)  4 soil refl dry
 *   (value at 2400 nm beyond)  5 rso  6 rdo  7 rsd  8 rdd.
 * rho_thermal / tau_thermal: leaf reflectance / transmittance beyond 2400 nm
 * (LeafBiology.rho_thermal / tau_thermal, prospect_5d.py:82-83, SPART.py:461-466; the reference's
 * default is 0.01 each).  flags: 0, SPART_FLAG_SOIL_SPECTRUM and / or SPART_FLAG_USER_LIDF. */
int spart_forward_spectrum(const SpartCtx* ctx, const double* params_dev, int64_t n, int64_t ld,
                           int32_t flags, double rho_thermal, double tau_thermal,
                           void* workspace_dev, double* out_dev, void* stream);

/* Replaces SMAC(angles, atm, coefs) (smac.py:14-213) for sensor `sensor`.  params_dev:
 * [SPART_NPAR][ld]; only the angle and atmosphere rows 19..25 matter, the others must merely be
 * finite.  out_dev: double [n][9][n_bands] in the field order of the reference's
 * AtmosphericOptics: Ta_s, Ta_o, Tg, Ra_dd, Ra_so, Ta_ss, Ta_sd, Ta_oo, Ta_do. */
int spart_smac(const SpartCtx* ctx, int32_t sensor, const double* params_dev, int64_t n, int64_t ld,
               void* workspace_dev, double* out_dev, void* stream);

/* Replaces SAILH(soil, leafopt, canopy, angles) on caller-supplied spectra (sailh.py:14-237).
 * params_dev: [SPART_NPAR][ld]; only the canopy and angle rows 15..21 matter, the others must
 * merely be finite.  soil_refl / leaf_refl / leaf_tran: double spectra of SPART_NWL_S (2162)
 * wavelengths, sample s at offset s * spectra_stride (spectra_stride = 0: one spectrum shared by
 * all samples).  out_dev: double [n][4][2162] = rso, rdo, rsd, rdd. */
int spart_sailh(const SpartCtx* ctx, const double* params_dev, int64_t n, int64_t ld,
                const double* soil_refl_dev, const double* leaf_refl_dev, const double* leaf_tran_dev,
                int64_t spectra_stride, void* workspace_dev, double* out_dev, void* stream);

/* Replaces CanopyStructure.__init__'s calculate_leafangles (sailh.py:340-398): the 13-class
 * leaf inclination distribution for n (LIDFa, LIDFb) pairs.  ab_dev: double [2][ld];
 * out_dev: double [n][13].  Needs no context. */
int spart_leafangles(const double* ab_dev, int64_t n, int64_t ld, double* out_dev, void* stream);

/* Consumer of a LUT (new; the reference has no retrieval step): for each of m observed band
 * vectors obs_dev [m][n_bands] (float) the index of the LUT entry lut_dev [n][n_bands] (float, e.g.
 * one output column of spart_forward_bands cast to float) with the smallest weighted squared
 * distance sum_b (sqrt_w[b] (obs_b - lut_b))^2, ties to the lowest index.  weights_dev: sqrt of the
 * band weights [n_bands] or NULL (all 1).  workspace_dev: spart_lut_workspace_bytes(m) bytes.
 * Outputs best_index_dev int64 [m] (= index_offset + local index), best_cost_dev float [m].
 * Sharded tables: every GPU searches its own slice with index_offset = first global entry of the
 * slice and packed_dev != NULL; the call then leaves one word (cost bits << 32 | global index)
 * per observation in packed_dev [m] instead of unpacking it, the caller min-reduces the words
 * over the GPUs (ncclMin on 64-bit integers: costs are >= 0, so integer order = (cost, index)
 * order, ties to the lowest global index) and calls spart_lut_unpack.  The full table never has
 * to be gathered on one GPU.  index_offset + n < 2^32. */
size_t spart_lut_workspace_bytes(int64_t m);
int spart_lut_nearest(const float* lut_dev, int64_t n, int32_t n_bands, const float* obs_dev, int64_t m,
                      const float* weights_dev, int64_t index_offset, void* workspace_dev,
                      int64_t* best_index_dev, float* best_cost_dev, unsigned long long* packed_dev,
                      void* stream);
/* The same search on the tensor cores (3xTF32 mma.sync over the band axis, n_bands <= 30): the entry is
 * chosen by an approximate comparison (error ~1e-6 |obs||entry|, so near-ties may resolve differently),
 * its reported cost is recomputed exactly in FP32.  Same arguments as spart_lut_nearest. */
int spart_lut_nearest_tc(const float* lut_dev, int64_t n, int32_t n_bands, const float* obs_dev, int64_t m,
                         const float* weights_dev, int64_t index_offset, void* workspace_dev,
                         int64_t* best_index_dev, float* best_cost_dev, unsigned long long* packed_dev,
                         void* stream);
int spart_lut_unpack(const unsigned long long* packed_dev, int64_t m, int64_t* best_index_dev,
                     float* best_cost_dev, void* stream);

/* Per-kernel timing of spart_forward_bands with CUDA events recorded on the caller's stream
 * (used by bench.py for the roofline).  After spart_profile_enable(ctx, 1) every
 * spart_forward_bands call records events around its SPART_NKERNELS kernels; spart_profile_read
 * waits for them and returns in kernel_ms[0..2] the summed durations [ms] of the leaf-angle
 * kernel, the per-sample geometry kernel and the per-(sample, band) kernel over `*calls` calls,
 * then clears the list. */
int spart_profile_enable(SpartCtx* ctx, int32_t on);
int spart_profile_read(SpartCtx* ctx, double* kernel_ms, int64_t* calls);

/* Micro-benchmarks used as roofline denominators by bench.py: dependent-free DFMA / FFMA
 * chains on all SMs.  Results in TFLOP/s (FMA = 2 flop). */
int spart_measure_peaks(int32_t device, double* fp64_tflops, double* fp32_tflops);

/* The FP64 rate of DEPENDENT DFMA chains on register operands, one chain per warp, eight warps per
 * scheduler (a degree-6 Horner step): on the B200 FP64 instructions of different warps issue every
 * 3 cycles, not 2, so this is ~2/3 of spart_measure_peaks' FP64 figure -- the issue ceiling of the
 * leaf-angle / geometry / band kernels, reported by bench.py beside the DFMA-chain peak. */
int spart_measure_fp64_chain(int32_t device, double* tflops);

/* Number of kernel launches this library has enqueued on this thread since load. */
int64_t spart_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SPART_B200_H */
