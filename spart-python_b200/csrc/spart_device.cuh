// spart_device.cuh -- device-side physics of the SPART forward model for sm_100a.
//
// Written from the model equations as implemented by the reference (file:line citations
// refer to wirrell/SPART-python, src/SPART/...).  Everything here is pure register math on
// one (sample, wavelength) or one sample; the kernels in spart_kernels.cu decide the thread
// mapping and the memory traffic.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <math_constants.h>
#include <stdint.h>

#include "math_coeffs.h"
#include "tau_coeffs.h"

namespace spart {

// ---- layouts shared with the host -----------------------------------------------------
enum ParamRow {
  P_CAB = 0, P_CDM, P_CW, P_CS, P_CCA, P_CANT, P_N, P_PROT, P_CBC,
  P_B, P_LAT, P_LON, P_SMP, P_SMC, P_FILM,
  P_LAI, P_LIDFA, P_LIDFB, P_Q,
  P_SZA, P_VZA, P_RAA,
  P_AOT, P_UO3, P_UH2O, P_PA,
  P_DOY, P_COUNT
};

// per-wavelength constants (SpartTables.lc rows)
enum LcRow {
  LC_KAB = 0, LC_KCA, LC_KDM, LC_KW, LC_KS, LC_KANT, LC_CBC, LC_PROT,
  LC_TALPH, LC_T12, LC_T21,
  LC_GSV0, LC_GSV1, LC_GSV2,
  LC_SOILC1, LC_SOILP, LC_SOILRW,
  LC_COUNT
};

// per-band folded SMAC constants (SpartSensor.smac rows)
enum SmacRow {
  SM_AH2O = 0, SM_NH2O, SM_AO3, SM_NO3,
  SM_AO2, SM_NO2, SM_NPO2,
  SM_ACO2, SM_NCO2, SM_NPCO2,
  SM_ACH4, SM_NCH4, SM_NPCH4,
  SM_ANO2, SM_NNO2, SM_NPNO2,
  SM_ACO, SM_NCO, SM_NPCO,
  SM_A0S, SM_A1S, SM_A2S, SM_A3S,
  SM_A0T, SM_A1T, SM_A2T, SM_A3T,
  SM_TAUR, SM_A0TAUP, SM_A1TAUP,
  SM_WO, SM_AK2, SM_AK, SM_OPB, SM_OMB, SM_OPB2, SM_OMB2, SM_WW, SM_G3, SM_D3, SM_H3,
  SM_A0P, SM_A1P, SM_A2P, SM_A3P, SM_A4P,
  SM_REST1, SM_REST2, SM_REST3, SM_REST4,
  SM_RESR1, SM_RESR2, SM_RESR3,
  SM_RESA1, SM_RESA2, SM_RESA3, SM_RESA4,
  SM_RESR2TAUR,                 // Resr2 * taur: a pure-coefficient product (float32 for Sentinel-2)
  SM_AKD3,                      // ak / (3 - wo 3 gc)
  SM_USED,
  SM_COUNT = 60
};

// per-sample record written by the sample kernel, read by the band / spectrum kernels
enum RecRow {
  R_F1 = 0, R_F2, R_F3,        // BSM soil-vector weights (bsm.py:49-51)
  R_MU, R_EMU,                 // Poisson mean (SMp-5)/SMC and exp(-mu) (bsm.py:101,121)
  R_K_SUN, R_K_OBS, R_BF, R_SOB, R_SOF,   // k, K, bf, sob, sof (sailh.py:93-97)
  R_TAUSS, R_TAUOO,            // exp(-k LAI), exp(-K LAI) (sailh.py:200-201)
  R_SUMPSO, R_PSO2W,           // sum(Pso[0:60])*iLAI and Pso[60] (sailh.py:216,219)
  R_US, R_UV, R_M, R_PEQ,      // SMAC geometry / pressure scalars (smac.py:98-102)
  R_LO3, R_LH2O, R_LM, R_LPEQ, // ln(uo3 m), ln(uh2o m), ln m, ln Peq
  R_CKSI, R_KSID, R_RAYPH,     // scattering angle terms (smac.py:129-141)
  R_ETSCALE,                   // cf(DOY) cos(sza)/pi (SPART.py:345-353)
  R_INVUS, R_INVUV,            // 1/us, 1/uv
  R_INV1PUS, R_INV1PUV,        // 1/(1+us), 1/(1+uv) (smac.py:125-126)
  R_AA3,                       // us uv / (us + uv) (smac.py:173)
  R_Z,                         // (1 - tau_ss tau_oo) / (K + k) (sailh.py:203)
  R_COUNT
};

#define SPART_PI 3.141592653589793238462643383279502884
#define SPART_DEG2RAD (SPART_PI / 180.0)

__constant__ double c_tau_coef[SPART_TAU_NINT][SPART_TAU_DEG + 1];
__constant__ double c_tau_mid[SPART_TAU_NINT];
__constant__ double c_tau_invhalf[SPART_TAU_NINT];

// Gauss-Legendre rules on [-1, 1] for the hot-spot integral (tools/gen_math_coeffs.py):
// 6-point rule for the narrow last layer
#define SPART_NQ1 6
__constant__ double c_gl6_x[SPART_NQ1] = SPART_GL6_X;
__constant__ double c_gl6_w[SPART_NQ1] = SPART_GL6_W;
// 16-point rule for the two graded panels
#define SPART_NQ2 16
__constant__ double c_glp_x[SPART_NQ2] = SPART_GL16_X;
__constant__ double c_glp_w[SPART_NQ2] = SPART_GL16_W;

// A parameter batch as the kernels see it: [P_COUNT][ld] of T (double; float with SPART_FLAG_F32_IO).
// Rows whose bit is set in `bc` are constant over the batch ("broadcast rows"): only their element 0
// is ever read, which all lanes load from the same address.
// `pb` is the base the broadcast rows are read from: the same as `p` for a caller's batch; the host-buffer path
// evaluates chunks of a larger device-resident span (p = span + chunk offset) whose broadcast elements sit at the
// start of the span's rows (pb = span).
template <typename T>
struct ParamsT {
  const T* p;
  int64_t ld;
  uint32_t bc;
  const T* pb;
  __host__ __device__ ParamsT(const T* p_, int64_t ld_, uint32_t bc_, const T* pb_ = nullptr)
      : p(p_), ld(ld_), bc(bc_), pb(pb_ ? pb_ : p_) {}
  __host__ __device__ __forceinline__ const T* ptr(int row, int64_t s) const {
    return ((bc >> row) & 1u) ? pb + row * ld : p + row * ld + s;
  }
  __device__ __forceinline__ T at(int row, int64_t s) const { return __ldg(ptr(row, s)); }
};
using Params = ParamsT<double>;

// ---- bounded-range sine / cosine ------------------------------------------------------------
// |x| is at most a few pi here (leaf-angle iteration), so a two-term Cody-Waite reduction by
// pi/2 is exact to the last bit and no large-argument path is needed.  Accuracy ~1 ulp
// (tools/gen_math_coeffs.py checks the polynomial kernels against 50-digit values).
// Polynomial coefficients live in constant memory so that DFMA reads them as c[bank][offset]
// operands; as literals the compiler rebuilds each one with two moves on every call.
__constant__ double c_sin_poly[7] = SPART_SIN_POLY;
__constant__ double c_cos_poly[7] = SPART_COS_POLY;
__constant__ double c_trig_red[4] = {SPART_TWO_OVER_PI, 6755399441055744.0, SPART_PIO2_HI, SPART_PIO2_LO};

__device__ __forceinline__ void sincos_small(double x, double& sn, double& cs) {
  const double* PS = c_sin_poly;
  const double* PC = c_cos_poly;
  // round-to-nearest of x * 2/pi with the 1.5 * 2^52 trick: no F2I / I2F conversions
  const double magic = c_trig_red[1];
  const double t = fma(x, c_trig_red[0], magic);
  const int q = __double2loint(t);
  const double qd = t - magic;
  double r = fma(-qd, c_trig_red[2], x);
  r = fma(-qd, c_trig_red[3], r);
  const double z = r * r;
  double ps = PS[6], pc = PC[6];
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    ps = fma(ps, z, PS[i]);
    pc = fma(pc, z, PC[i]);
  }
  const double s0 = fma(r * z, ps, r);
  const double c0 = fma(z * z, pc, fma(-0.5, z, 1.0));
  const double sa = (q & 1) ? c0 : s0;
  const double ca = (q & 1) ? s0 : c0;
  // quadrant signs applied on the sign bit (integer pipe) instead of FP64 negations
  sn = __hiloint2double(__double2hiint(sa) ^ ((q & 2) << 30), __double2loint(sa));
  cs = __hiloint2double(__double2hiint(ca) ^ (((q + 1) & 2) << 30), __double2loint(ca));
}

// x != 0 for a stored coefficient, tested on the integer pipe (no coefficient is subnormal)
__device__ __forceinline__ bool nonzero_coef(double x) { return (__double2hiint(x) << 1) != 0; }

// ---- polynomial evaluation -------------------------------------------------------------------
// poly_eval<DEG, WAYS>: WAYS = 1 is a single Horner chain; WAYS = 2 / 4 evaluate the coefficients
// of equal index mod WAYS as independent Horner chains in x^WAYS (a few extra operations, 1/WAYS
// of the dependency depth) -- the band kernels run 4-5 warps per scheduler and lose more to the
// latency of dependent DFMAs than to their count.
#ifndef SPART_TAU_SPLIT
#define SPART_TAU_SPLIT 2
#endif
#ifndef SPART_EXP_SPLIT
#define SPART_EXP_SPLIT 2
#endif
#ifndef SPART_LOG_SPLIT
#define SPART_LOG_SPLIT 2
#endif
template <int DEG, int WAYS>
__device__ __forceinline__ double poly_eval(const double* c, double x) {
  if constexpr (WAYS == 1) {
    double p = c[DEG];
#pragma unroll
    for (int i = DEG - 1; i >= 0; --i) p = fma(p, x, c[i]);
    return p;
  } else {
    double xw = x * x;
    if constexpr (WAYS == 4) xw = xw * xw;
    double q[WAYS];
#pragma unroll
    for (int w = 0; w < WAYS; ++w) {
      const int top = DEG - ((DEG - w) % WAYS);   // largest index <= DEG congruent to w
      double p = c[top];
#pragma unroll
      for (int i = top - WAYS; i >= 0; i -= WAYS) p = fma(p, xw, c[i]);
      q[w] = p;
    }
    if constexpr (WAYS == 2) {
      return fma(q[1], x, q[0]);
    } else {
      const double x2 = x * x;
      return fma(fma(q[3], x, q[2]), x2, fma(q[1], x, q[0]));
    }
  }
}

// ---- reciprocal without the slow path ---------------------------------------------------------
// `1.0 / x` compiles to MUFU.RCP64H + 5 DFMA + a range check that branches to an out-of-line
// fix-up for subnormal / huge operands (~12 instructions and two basic-block boundaries per
// division).  Every denominator on the band path is a normal number of moderate magnitude, so
// the seed (good to ~2^-20) and one third-order step suffice: MUFU + 3 DFMA, 0.51 ulp measured
// (tools/micro/math_check.cu).  0 -> inf/NaN and NaN -> NaN as with a true division; subnormal or
// > 2^1022 denominators are not supported.  SPART_RCP_ORDER3 = 0 selects two Newton steps.
#ifndef SPART_FAST_RCP
#define SPART_FAST_RCP 1
#endif
#ifndef SPART_RCP_ORDER3
#define SPART_RCP_ORDER3 1
#endif
__device__ __forceinline__ double rcp_fast(double x) {
#if !SPART_FAST_RCP
  return 1.0 / x;
#else
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#if SPART_RCP_ORDER3
  // the seed is good to ~2^-20: one third-order step r (1 + e + e^2) leaves e^3 ~ 2^-60
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
#else
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
#endif
#endif
}

// sqrt from the rsqrt seed and one third-order step (0.76 ulp measured), branch free: sqrt(0) = 0,
// negative / NaN -> NaN.  (+inf and subnormal arguments are not supported.)
__device__ __forceinline__ double sqrt_fast(double x) {
#if !SPART_FAST_RCP
  return sqrt(x);
#else
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = (x == 0.0) ? 0.0 : r;
#if SPART_RCP_ORDER3
  // g = x r ~ sqrt(x), e = 1 - x r^2; sqrt(x) = g (1 - e)^(-1/2) = g (1 + e/2 + 3 e^2/8 + O(e^3))
  const double g = x * r;
  const double e = fma(-g, r, 1.0);
  const double t = fma(0.375, e, 0.5) * e;
  return fma(g, t, g);
#else
  double g = x * r, h = 0.5 * r;
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-h, g, 0.5);
  return fma(g, e, g);
#endif
#endif
}

// ---- exp / log with constant-bank coefficients ----------------------------------------------
// libdevice's exp/log rebuild each of their ~13 polynomial coefficients with two move
// instructions per call (46 / ~100 SASS instructions per call, 15 / ~25 of them FP64).  These
// versions read the coefficients as c[bank][offset] operands of the DFMAs: 19 / ~32
// instructions, same <= 1 ulp accuracy class (tools/gen_math_coeffs.py).  Arguments outside
// the plain range are handled branch-free with NaN-preserving selects (see each function).
__constant__ double c_exp_poly[12] = SPART_EXP_POLY;
__constant__ double c_log_poly[9] = SPART_LOG_POLY;
__constant__ double c_explog_red[4] = {SPART_LOG2E, 6755399441055744.0, SPART_LN2_HI, SPART_LN2_LO};


#ifndef SPART_FAST_EXP
#define SPART_FAST_EXP 1
#endif
#ifndef SPART_ESTRIN
#define SPART_ESTRIN 0
#endif
#ifndef SPART_FAST_LOG
#define SPART_FAST_LOG 1
#endif

// e^x = 2^k * 2^(j/64) * e^r with |r| <= ln2/128: 2^(j/64) comes from a 64-entry shared-memory
// table (lanes index it independently, which the constant cache would serialise) and e^r - 1 from
// r + r^2 q(r) with a cubic q: 10 FP64 operations per call instead of 17 for the table-free kernel
// (degree-11 polynomial), <= 1 ulp (tools/gen_math_coeffs.py).  Every kernel that calls one of the
// exp_* functions must run exp_table_load() and a __syncthreads() first.
#ifndef SPART_EXP_TABLE
#define SPART_EXP_TABLE 1
#endif
#ifndef SPART_EXP_INT_SCALE
#define SPART_EXP_INT_SCALE 1
#endif
#ifndef SPART_EXP_INT_CLAMP
#define SPART_EXP_INT_CLAMP 1
#endif
__constant__ double c_exp2_tab[64] = SPART_EXP2_TABLE;
__constant__ double c_expt_poly[4] = SPART_EXPT_POLY;
__constant__ double c_expt_red[4] = {SPART_64_OVER_LN2, 6755399441055744.0, SPART_LN2_64_HI, SPART_LN2_64_LO};
__shared__ double s_exp2_tab[64];

__device__ __forceinline__ void exp_table_load() {
#if SPART_EXP_TABLE
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_exp2_tab[i] = c_exp2_tab[i];
#endif
}

// core shared by all variants: returns p ~ e^x / 2^k, p in [1, 2), and k.
__device__ __forceinline__ double exp_core(double x, int& k) {
#if SPART_EXP_TABLE
  const double magic = c_expt_red[1];
  const double t = fma(x, c_expt_red[0], magic);
  const int nq = __double2loint(t);
  const double nd = t - magic;
  double r = fma(-nd, c_expt_red[2], x);
  r = fma(-nd, c_expt_red[3], r);
  const double T = s_exp2_tab[nq & 63];
  k = nq >> 6;
  const double r2 = r * r;
  const double q = fma(fma(c_expt_poly[3], r2, c_expt_poly[1]), r, fma(c_expt_poly[2], r2, c_expt_poly[0]));
  return fma(T, fma(r2, q, r), T);
#else
  const double magic = c_explog_red[1];
  const double t = fma(x, c_explog_red[0], magic);
  k = __double2loint(t);
  const double kd = t - magic;
  double r = fma(-kd, c_explog_red[2], x);
  r = fma(-kd, c_explog_red[3], r);
  return poly_eval<11, SPART_EXP_SPLIT>(c_exp_poly, r);
#endif
}

// p * 2^k for p in [0.99, 2) and |k| <= 1011 (the clamped variants): the exponent field is adjusted
// on the integer pipe instead of a DMUL.  NaN stays NaN: a NaN argument gives the canonical NaN in
// exp_core's first FMA, whose low word -- and with it k -- is 0.
__device__ __forceinline__ double exp_scale(double p, int k) {
#if SPART_EXP_INT_SCALE
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
  return p * __hiloint2double((k + 1023) << 20, 0);
#endif
}

// Full-range exp, branch free: the argument is clamped to [-745.2, 709.8] with NaN-preserving
// selects (exp(-inf) = 0, exp(+inf) = inf, denormal results rounded correctly by the two-step
// scaling, NaN in -> NaN out).
__device__ __forceinline__ double exp_fast(double x) {
#if !SPART_FAST_EXP
  return exp(x);
#endif
  double xc = x;
  xc = (x < -745.2) ? -745.2 : xc;
  xc = (x > 709.8) ? 709.8 : xc;
  int k;
  const double p = exp_core(xc, k);
  const int k1 = k >> 1, k2 = k - k1;
  const double s1 = __hiloint2double((k1 + 1023) << 20, 0);
  const double s2 = __hiloint2double((k2 + 1023) << 20, 0);
  return (p * s1) * s2;
}

// exp with the argument clamped to [-700, 700] by NaN-preserving selects and one-step scaling:
// results below e^-700 ~ 1e-304 come out as 1e-304 instead of 0 / subnormal and overflow
// saturates at e^700; used where such values are physically irrelevant (transmittances,
// Poisson weights).  NaN in -> NaN out.
__device__ __forceinline__ double exp_clamp(double x) {
#if !SPART_FAST_EXP
  return exp(x);
#endif
#if SPART_EXP_INT_CLAMP
  // |x| > 700 (infinities included, NaN excluded) tested and replaced on the high word, integer pipe.
  // Arguments are arithmetic results, so a NaN is the canonical 0x7ff8... pattern.
  const int hx = __double2hiint(x);
  const unsigned ax = (unsigned)hx & 0x7fffffffu;
  const bool big = (ax - 0x4085E001u) <= (0x7FF00000u - 0x4085E001u);
  const double xc = __hiloint2double(big ? (int)(0x4085E000u | ((unsigned)hx & 0x80000000u)) : hx,
                                     big ? 0 : __double2loint(x));
#else
  double xc = x;
  xc = (x < -700.0) ? -700.0 : xc;
  xc = (x > 700.0) ? 700.0 : xc;
#endif
  int k;
  const double p = exp_core(xc, k);
  return exp_scale(p, k);
}

// exp_clamp for arguments that are never large and positive (-tau/mu, -m LAI, ...): only the lower
// clamp is needed.
__device__ __forceinline__ double exp_neg(double x) {
#if !SPART_FAST_EXP
  return exp(x);
#endif
#if SPART_EXP_INT_CLAMP
  // x < -700 (or -inf) on the high word: as an unsigned number it is then above that of -700.0
  // (a canonical NaN, 0x7ff8..., is below and passes through)
  const int hx = __double2hiint(x);
  const bool big = (unsigned)hx > 0xC085E000u;
  const double xc = __hiloint2double(big ? (int)0xC085E000u : hx, big ? 0 : __double2loint(x));
#else
  const double xc = (x < -700.0) ? -700.0 : x;
#endif
  int k;
  const double p = exp_core(xc, k);
  return exp_scale(p, k);
}

// exp for call sites that guarantee |x| <= 700 for finite inputs (NaN still propagates).
__device__ __forceinline__ double exp_bounded(double x) {
#if !SPART_FAST_EXP
  return exp(x);
#endif
  int k;
  const double p = exp_core(x, k);
  return exp_scale(p, k);
}

__device__ __forceinline__ double log_fast(double x) {
#if !SPART_FAST_LOG
  return log(x);
#endif
  const int hi = __double2hiint(x);
  int e = (hi >> 20) - 1023;
  int mh = (hi & 0x000fffff) | 0x3ff00000;
  if (mh >= 0x3ff6a09f) {        // mantissa above sqrt(2): halve it so that m is in [0.707, 1.414)
    mh -= 0x00100000;
    e += 1;
  }
  const double m = __hiloint2double(mh, __double2loint(x));
  const double s = (m - 1.0) * rcp_fast(m + 1.0);
  const double z = s * s;
  const double p = poly_eval<8, SPART_LOG_SPLIT>(c_log_poly, z);
  const double s2 = s + s;
  const double l = fma(s2 * z, p, s2);     // 2 atanh(s) = ln m
  // int -> double without I2F: 2^52 + 2^31 + e, minus the bias
  const double ed = __hiloint2double(0x43300000, e ^ 0x80000000) - 4503601774854144.0;
  const double res = fma(ed, c_explog_red[2], fma(ed, c_explog_red[3], l));
  // special arguments, branch free: +inf / NaN -> x + x, 0 -> -inf, negative -> NaN
  // (positive subnormals are not supported: they return x + x)
  const bool plain = (unsigned)(hi - 0x00100000) < 0x7fe00000u;
  // classified on the bit pattern (integer pipe): +-0 -> -inf, negative -> NaN, +inf / NaN -> x + x
  const bool zero = (((unsigned)hi << 1) | (unsigned)__double2loint(x)) == 0u;
  const double special = zero ? -CUDART_INF : ((hi < 0) ? CUDART_NAN : x + x);
  return plain ? res : special;
}

// ---- PROSPECT plate transmissivity -----------------------------------------------------
// tau(K) = (1-K) e^-K + K^2 E1(K)  (prospect_5d.py:182-196).  The reference integrates
// e^-t/t numerically per wavelength; here E1 comes from piecewise polynomials generated in
// 60-digit arithmetic (tools/gen_tau_coeffs.py).  `tab` points at a shared-memory copy of
// the coefficient table so that lanes in different intervals do not serialise on the
// constant cache.
struct TauTable {
  double coef[SPART_TAU_NINT][SPART_TAU_DEG + 1];
  double mid[SPART_TAU_NINT];
  double invhalf[SPART_TAU_NINT];
};

__device__ __forceinline__ void load_tau_table(TauTable* s) {
  double* dst = reinterpret_cast<double*>(s);
  const int ncoef = SPART_TAU_NINT * (SPART_TAU_DEG + 1);
  for (int i = threadIdx.x; i < ncoef + 2 * SPART_TAU_NINT; i += blockDim.x) {
    double v;
    if (i < ncoef) v = (&c_tau_coef[0][0])[i];
    else if (i < ncoef + SPART_TAU_NINT) v = c_tau_mid[i - ncoef];
    else v = c_tau_invhalf[i - ncoef - SPART_TAU_NINT];
    dst[i] = v;
  }
}

__device__ __forceinline__ double plate_tau(double K, const TauTable* tab) {
  // caller guarantees K > 0
  const double emk = exp_fast(-K);
  int idx;
  double u, t = rcp_fast(K);
  if (K < 1.0) {
    idx = 0;
    u = 2.0 * K - 1.0;
  } else {
    int e = (__double2hiint(K) >> 20) - 1023;   // floor(log2 K) for K >= 1
    idx = min(e + 1, SPART_TAU_NINT - 1);
    u = (t - tab->mid[idx]) * tab->invhalf[idx];
  }
  const double* c = tab->coef[idx];
#if SPART_ESTRIN
  static_assert(SPART_TAU_DEG == 16, "Estrin evaluation below is written for degree 16");
  const double u2 = u * u;
  const double a0 = fma(c[1], u, c[0]), a1 = fma(c[3], u, c[2]), a2 = fma(c[5], u, c[4]), a3 = fma(c[7], u, c[6]);
  const double a4 = fma(c[9], u, c[8]), a5 = fma(c[11], u, c[10]), a6 = fma(c[13], u, c[12]), a7 = fma(c[15], u, c[14]);
  const double u4 = u2 * u2;
  const double b0 = fma(a1, u2, a0), b1 = fma(a3, u2, a2), b2 = fma(a5, u2, a4), b3 = fma(a7, u2, a6);
  const double u8 = u4 * u4;
  const double d0 = fma(b1, u4, b0), d1 = fma(b3, u4, b2);
  const double p = fma(c[16], u8 * u8, fma(d1, u8, d0));
#else
  const double p = poly_eval<SPART_TAU_DEG, SPART_TAU_SPLIT>(c, u);
#endif
  if (K < 1.0) {
    const double e1 = fma(K, p, -0.57721566490153286061 - log_fast(K));
    return (1.0 - K) * emk + K * K * e1;
  }
  return emk * t * p;
}

// ---- PROSPECT-5D / PROSPECT-PRO at one wavelength (prospect_5d.py:117-246) --------------
// Divisions that share a denominator are done as one reciprocal and multiplications; this
// changes individual roundings by <= 1 ulp with respect to the reference's operation order.
struct LeafPar {
  double Cab, Cca, Cdm, Cw, Cs, Cant, CBC, PROT, N, invN;
};

__device__ __forceinline__ LeafPar load_leaf(const Params& P, int64_t s) {
  LeafPar L;
  L.Cab = P.at(P_CAB, s);
  L.Cdm = P.at(P_CDM, s);
  L.Cw = P.at(P_CW, s);
  L.Cs = P.at(P_CS, s);
  L.Cca = P.at(P_CCA, s);
  L.Cant = P.at(P_CANT, s);
  L.N = P.at(P_N, s);
  L.PROT = P.at(P_PROT, s);
  L.CBC = P.at(P_CBC, s);
  // PROSPECT-PRO switch, prospect_5d.py:148-155
  if ((L.PROT > 0.0 || L.CBC > 0.0) && L.Cdm > 0.0) L.Cdm = 0.0;
  L.invN = rcp_fast(L.N);
  return L;
}

__device__ __forceinline__ void leaf_from_tau(double tau, double t_alph, double t12, double t21, double N,
                                              double& refl, double& tran);

// lc: the SPART_NLC constants of this wavelength.  Returns refl, tran (and kChlrel).
template <bool kWantKchl>
__device__ __forceinline__ void prospect_point(const LeafPar& L, const double* lc, const TauTable* tab,
                                               double& refl, double& tran, double& kchl) {
  const double Ksum = L.Cab * lc[LC_KAB] + L.Cca * lc[LC_KCA] + L.Cdm * lc[LC_KDM] + L.Cw * lc[LC_KW] +
                      L.Cs * lc[LC_KS] + L.Cant * lc[LC_KANT] + L.CBC * lc[LC_CBC] + L.PROT * lc[LC_PROT];
  const double Kall = Ksum * L.invN;
  double tau = 1.0;
  kchl = 0.0;
  if (Kall > 0.0) {
    tau = plate_tau(Kall, tab);
    if (kWantKchl) kchl = L.Cab * lc[LC_KAB] / (Kall * L.N);
  }
  leaf_from_tau(tau, lc[LC_TALPH], lc[LC_T12], lc[LC_T21], L.N, refl, tran);
}

// Leaf reflectance / transmittance from the plate transmissivity tau (prospect_5d.py:208-241).  Also
// used by the FP32 mode for near-conservative plates, where the Stokes system is ill-conditioned in
// 1 - r - t and single precision cannot hold the leaf absorptance 1 - refl - tran.
__device__ __forceinline__ void leaf_from_tau(double tau, double t_alph, double t12, double t21, double N,
                                              double& refl, double& tran) {
  const double r_alph = 1.0 - t_alph, r12 = 1.0 - t12, r21 = 1.0 - t21;

  // one plate, prospect_5d.py:208-214
  const double tt21 = tau * t21;
  const double inv_d1 = rcp_fast(1.0 - r21 * r21 * tau * tau);
  const double Ta = t_alph * tt21 * inv_d1;
  const double Ra = r_alph + r21 * tau * Ta;
  const double t = t12 * tt21 * inv_d1;
  const double r = r12 + r21 * tau * t;

  // Stokes system for the remaining N-1 plates, prospect_5d.py:219-230
  double Rsub, Tsub;
  const double Nm1 = N - 1.0;
  if (r + t >= 1.0) {  // zero absorption, prospect_5d.py:233-235
    Tsub = t * rcp_fast(t + (1.0 - t) * Nm1);
    Rsub = 1.0 - Tsub;
  } else {
    const double D = sqrt_fast((1.0 + r + t) * (1.0 + r - t) * (1.0 - r + t) * (1.0 - r - t));
    const double rq = r * r, tq = t * t;
    const double a = (1.0 + rq - tq + D) * rcp_fast(2.0 * r);
    const double b = (1.0 - rq + tq + D) * rcp_fast(2.0 * t);
    // b ** (N - 1): b >= 1 and |(N-1) ln b| is O(1), so exp_fast(y ln b) is accurate to a few ulp;
    // the exact cases of pow are kept (y == 0 -> 1, b == inf -> inf).
    const double bNm1 = (Nm1 == 0.0) ? 1.0 : exp_clamp(Nm1 * log_fast(b));
    const double bN2 = bNm1 * bNm1;
    const double a2 = a * a;
    const double inv_d2 = rcp_fast(a2 * bN2 - 1.0);
    Rsub = a * (bN2 - 1.0) * inv_d2;
    Tsub = bNm1 * (a2 - 1.0) * inv_d2;
  }
  const double inv_d3 = rcp_fast(1.0 - Rsub * r);  // prospect_5d.py:239-241
  tran = Ta * Tsub * inv_d3;
  refl = Ra + Ta * Rsub * t * inv_d3;
}

// ---- BSM soil at one wavelength (bsm.py:49-52, 99-124) ----------------------------------
struct SoilPar {
  double f1, f2, f3, mu, emu, film;
};

__device__ __forceinline__ void bsm_point(const SoilPar& S, const double* lc, double& rwet, double& rdry) {
  rdry = S.f1 * lc[LC_GSV0] + S.f2 * lc[LC_GSV1] + S.f3 * lc[LC_GSV2];
  rwet = rdry;
  if (S.mu > 0.0) {
    const double rbac = 1.0 - (1.0 - rdry) * (rdry * lc[LC_SOILC1] + 1.0 - rdry);
    const double p = lc[LC_SOILP], Rw = lc[LC_SOILRW];
    const double tw1 = exp_neg(-2.0 * lc[LC_KW] * S.film);
    double fk = S.emu;           // Poisson weight k = 0
    double acc = rdry * fk;
    double tw = 1.0;
    const double g = (1.0 - Rw) * (1.0 - p);
#pragma unroll
    for (int k = 1; k <= 6; ++k) {
      tw *= tw1;                 // exp_clamp(-2 kw film k)
      fk = fk * S.mu * (1.0 / (double)k);
      const double x = tw * rbac;
      acc += (Rw + g * x * rcp_fast(1.0 - p * x)) * fk;
    }
    rwet = acc;
  }
}

// ---- SAILH four-stream solution at one wavelength (sailh.py:99-105, 142-233) ------------
struct CanopyGeo {
  double LAI, k, K, bf, sob, sof, tau_ss, tau_oo, sumpso, pso2w, Z;
};

__device__ __forceinline__ double sail_J1(double m, double k, double LAI, double em, double ek) {
  // calcJ1 at x = -1 (sailh.py:154-170); em = exp_fast(-m LAI), ek = exp_fast(-k LAI)
  if (fabs((m - k) * LAI) < 1e-6) {
    return 0.5 * (em + ek) * LAI * (1.0 - (1.0 / 12.0) * (k - m) * (k - m) * LAI * LAI);
  }
  return (em - ek) * rcp_fast(k - m);
}

__device__ __forceinline__ double sail_J1_nb(double m, double k, double LAI, double em, double ek);

template <bool kBranchFree = false>
__device__ __forceinline__ void sailh_point(const CanopyGeo& G, double rho, double tau, double rs, double& rso,
                                            double& rdo, double& rsd, double& rdd) {
  const double k = G.k, K = G.K, bf = G.bf, LAI = G.LAI;
  const double sdb = 0.5 * (k + bf), sdf = 0.5 * (k - bf);
  const double ddb = 0.5 * (1.0 + bf), ddf = 0.5 * (1.0 - bf);
  const double dob = 0.5 * (K + bf), dof = 0.5 * (K - bf);

  const double sigb = ddb * rho + ddf * tau;
  const double sigf = ddf * rho + ddb * tau;
  const double sb = sdb * rho + sdf * tau;
  const double sf = sdf * rho + sdb * tau;
  const double vb = dob * rho + dof * tau;
  const double vf = dof * rho + dob * tau;
  const double w = G.sob * rho + G.sof * tau;
  const double a = 1.0 - sigf;
  const double m = sqrt_fast(a * a - sigb * sigb);
  const double rinf = (a - m) * rcp_fast(sigb);
  const double rinf2 = rinf * rinf;

  const double e1 = exp_neg(-m * LAI);
  const double e2 = e1 * e1;
  const double tau_ss = G.tau_ss, tau_oo = G.tau_oo;
  const double inv_km = rcp_fast(k + m), inv_Km = rcp_fast(K + m);
  const double J1k = kBranchFree ? sail_J1_nb(m, k, LAI, e1, tau_ss) : sail_J1(m, k, LAI, e1, tau_ss);
  const double J2k = (1.0 - tau_ss * e1) * inv_km;   // calcJ2 at x = 0 (sailh.py:172-177)
  const double J1K = kBranchFree ? sail_J1_nb(m, K, LAI, e1, tau_oo) : sail_J1(m, K, LAI, e1, tau_oo);
  const double J2K = (1.0 - tau_oo * e1) * inv_Km;
  const double re = rinf * e1;
  const double inv_den = rcp_fast(1.0 - rinf2 * rinf2);

  const double s1 = sf + rinf * sb, s2 = sf * rinf + sb;
  const double v1 = vf + rinf * vb, v2 = vf * rinf + vb;
  const double Pss = s1 * J1k, Qss = s2 * J2k;
  const double Poo = v1 * J1K, Qoo = v2 * J2K;
  const double Z = G.Z;

  const double tau_dd = (1.0 - rinf2) * e1 * inv_den;
  const double rho_dd = rinf * (1.0 - e2) * inv_den;
  const double tau_sd = (Pss - re * Qss) * inv_den;
  const double tau_do = (Poo - re * Qoo) * inv_den;
  const double rho_sd = (Qss - re * Pss) * inv_den;
  const double rho_do = (Qoo - re * Poo) * inv_den;

  const double T1 = v2 * s1 * (Z - J1k * tau_oo) * inv_Km + v1 * s2 * (Z - J1K * tau_ss) * inv_km;
  const double T2 = -(Qoo * rho_sd + Poo * tau_sd) * rinf;
  const double rho_sod = (T1 + T2) * rcp_fast(1.0 - rinf2);
  const double rho_so = rho_sod + w * G.sumpso;

  const double rs_den = rs * rcp_fast(1.0 - rs * rho_dd);
  rso = rho_so + rs * G.pso2w + ((tau_sd + tau_ss * rs * rho_dd) * tau_oo + (tau_sd + tau_ss) * tau_do) * rs_den;
  rdo = rho_do + (tau_oo + tau_do) * tau_dd * rs_den;
  rsd = rho_sd + (tau_ss + tau_sd) * tau_dd * rs_den;
  rdd = rho_dd + tau_dd * tau_dd * rs_den;
}

// ---- leaf inclination distribution (sailh.py:351-398) ------------------------------------
// One step of the reference's dcum fixed-point iteration (sailh.py:378-382).  The iteration
// is stopped at |dx| <= 1e-8 and its result depends on the number of steps taken, so it is
// reproduced step for step; sin 2x is formed as 2 sin x cos x (<= 1 ulp from the direct value).
__device__ __forceinline__ bool dcum_step(double a, double b, double theta2, double& x, double& y) {
  double s, c;
  sincos_small(x, s, c);
  y = s * fma(b, c, a);           // a sin x + 0.5 b sin 2x
  const double dx = fma(0.5, y, 0.5 * (theta2 - x));   // 0.5 (y - x + theta2)
  x += dx;
  return !(fabs(dx) > 1e-8);      // converged (NaN input also stops)
}


// ---- per-class geometry (sailh.py:401-446) ---------------------------------------------
__device__ __forceinline__ void volscatt_class(double sin_tts, double cos_tts, double sin_tto, double cos_tto,
                                               double psi_rad, double sin_psi, double cos_psi, double sin_ttli,
                                               double cos_ttli, double& chi_s, double& chi_o, double& frho,
                                               double& ftau) {
  const double Cs = cos_ttli * cos_tts, Ss = sin_ttli * sin_tts;
  const double Co = cos_ttli * cos_tto, So = sin_ttli * sin_tto;
  const double As = fmax(Ss, Cs), Ao = fmax(So, Co);
  // the quotient is exactly -1 whenever Cs >= Ss (As == Cs) -- acos amplifies a 1-ulp deviation from -1 to
  // 1e-8, so that case is a select; otherwise -Cs / Ss comes from the fast reciprocal (<= 1 ulp; chi_s, chi_o are
  // stationary in bts, bto there, and the reference's own correctly rounded quotient is as ill-conditioned)
  // For the flat leaf classes (leaf angle + zenith angle <= 90 degrees) C >= S holds in every lane of the warp:
  // beta = acos(-1) = pi and sin beta = 0 exactly, no reciprocal, acos and square root (warp-uniform branch; the
  // calling lanes of a warp are converged here)
  double zs = -1.0, zo = -1.0, bts = SPART_PI, bto = SPART_PI, sbts = 0.0, sbto = 0.0;
  const unsigned active = __activemask();
  if (!__all_sync(active, Cs >= Ss)) {
    zs = (Cs >= Ss) ? -1.0 : -Cs * rcp_fast(Ss);
    bts = acos(zs);
    // sin(acos z) = sqrt(1 - z^2) (>= 0 on [0, pi]); differs from sin of the rounded angle by < 2e-16 absolute
    sbts = sqrt_fast(fma(-zs, zs, 1.0));
  }
  if (!__all_sync(active, Co >= So)) {
    zo = (Co >= So) ? -1.0 : -Co * rcp_fast(So);
    bto = acos(zo);
    sbto = sqrt_fast(fma(-zo, zo, 1.0));
  }
  chi_o = 2.0 / SPART_PI * ((bto - SPART_PI / 2.0) * Co + sbto * So);
  chi_s = 2.0 / SPART_PI * ((bts - SPART_PI / 2.0) * Cs + sbts * Ss);
  const double delta1 = fabs(bts - bto);
  const double delta2 = SPART_PI - fabs(bts + bto - SPART_PI);
  const double Tot = psi_rad + delta1 + delta2;
  const double bt1 = fmin(psi_rad, delta1);
  const double bt3 = fmax(psi_rad, delta2);
  const double bt2 = Tot - bt1 - bt3;
  // delta1 <= delta2, so (bt1, bt2, bt3) is (psi, delta1, delta2) sorted and the reference's cos(bt1), sin(bt2),
  // cos(bt3) are sines / cosines of psi, bts - bto and bts + bto.  The latter follow from cos / sin of bts, bto
  // (zs, sbts, zo, sbto) by the addition theorems -- no sincos per class (3 x 13 per sample before); the
  // values differ from sin / cos of the rounded angles by < 1e-14 absolute (NumPy study over 2e5 geometries).
  const double cc = zs * zo, ss = sbts * sbto, sc = sbts * zo, cs = zs * sbto;
  const double cos_d1 = cc + ss, sin_d1 = fabs(sc - cs);      // delta1 = |bts - bto|
  const double cos_d2 = cc - ss, sin_d2 = fabs(sc + cs);      // delta2 = pi - |bts + bto - pi|
  const bool lo = psi_rad <= delta1, hi = psi_rad > delta2;
  const double c1 = lo ? cos_psi : cos_d1;
  const double c3 = hi ? cos_psi : cos_d2;
  const double s2 = lo ? sin_d1 : (hi ? sin_d2 : sin_psi);
  const double T1 = 2.0 * Cs * Co + Ss * So * cos_psi;
  const double T2 = s2 * (2.0 * As * Ao + Ss * So * c1 * c3);
  const double Jmin = bt2 * T1 - T2;
  const double Jplus = (SPART_PI - bt2) * T1 + T2;
  frho = fmax(0.0, Jplus * (1.0 / (2.0 * SPART_PI * SPART_PI)));
  ftau = fmax(0.0, -Jmin * (1.0 / (2.0 * SPART_PI * SPART_PI)));
}

// ---- hot-spot integrals (sailh.py:116-135, 216-219) --------------------------------------
// pso(x) = exp_fast(A x + Cq (1 - e^{alpha x})),  A = (K+k) LAI,  Cq = sqrt(Kk) LAI / alpha.
// The reference needs sum_{j<60} Pso[j] * LAI/60 = LAI * int_{-1}^{0} pso dx and
// Pso[60] = 60 * int_{-1-1/60}^{-1} pso dx, each Pso[j] from one QUADPACK call.
// Here the first integral is split at x = -L, L = min(1, 40/alpha, 40/(A - sqrt(Kk) LAI)):
//   * below -L either e^{alpha x} < e^-40 (pso is a pure exponential, integrated in closed
//     form) or pso itself is < e^-40 of its peak (dropped);
//   * [-L, 0] is covered by two graded panels, [-L/5, 0] and [-L, -L/5], of a 16-point
//     Gauss-Legendre rule each.  The integrand is analytic and varies fastest next to the hot spot
//     at x = 0; with the split at L/5 the 32 nodes agree with 16 panels x 24 nodes to 3e-15 on the
//     benchmark distributions and to 4e-13 for q down to 1e-4, LAI up to 20 and zenith angles up
//     to 85 degrees (two equal panels of 24 nodes: 2e-15 / 4e-12; the study is described in
//     DESIGN.md).
// All lanes run the same trip counts (no divergence); cost 64 + 12 exp instead of 557.
#define SPART_NP 2
#define SPART_HOTSPOT_SPLIT 0.2
__device__ __forceinline__ void hotspot_integrals(double K, double k, double LAI, double q, double dso,
                                                  double& sumpso_ilai, double& pso2w) {
  const double A0 = (K + k) * LAI;
  const double S = sqrt_fast(K * k) * LAI;
  const double Amin = A0 - S;
  double A = A0, Cq = 0.0, alpha = 0.0;
  if (dso != 0.0) {
    alpha = (dso * rcp_fast(q)) * 2.0 * rcp_fast(k + K);
    Cq = S * rcp_fast(alpha);
  } else {
    A = Amin;   // sailh.py:127
  }
  double L = 1.0;
  if (alpha > 0.0) L = fmin(L, 40.0 * rcp_fast(alpha));
  if (Amin > 0.0) L = fmin(L, 40.0 * rcp_fast(Amin));
  double total = 0.0;
#pragma unroll 1
  for (int j = 0; j < SPART_NP; ++j) {
    // panel 0 = [-SPLIT L, 0], panel 1 = [-L, -SPLIT L]
    const double hw = (j == 0 ? 0.5 * SPART_HOTSPOT_SPLIT : 0.5 * (1.0 - SPART_HOTSPOT_SPLIT)) * L;   // half width
    const double xc = (j == 0 ? -0.5 * SPART_HOTSPOT_SPLIT : -0.5 * (1.0 + SPART_HOTSPOT_SPLIT)) * L;  // centre
    double acc = 0.0;
#pragma unroll 4
    for (int i = 0; i < SPART_NQ2; ++i) {
      const double x = fma(hw, c_glp_x[i], xc);
      const double ea = exp_bounded(alpha * x);                 // alpha x in [-40, 0]
      const double arg = fma(A, x, Cq * (1.0 - ea));
      acc = fma(c_glp_w[i], exp_bounded(arg), acc);             // arg in [-80, 0]: A L <= 40 A / Amin <= 80
    }
    total = fma(hw, acc, total);
  }
  if (L < 1.0 && alpha * L >= 40.0 * (1.0 - 1e-12)) {  // analytic pure-exponential remainder
    // int_{-1}^{-L} e^{Cq + A x} dx = e^{Cq - A L} (1 - e^{-A (1 - L)}) / A; for A (1 - L) < 1e-4 -- bare
    // soil, LAI = 0, included -- the quotient is taken from its series (truncation < 5e-14 relative)
    const double w = 1.0 - L, xr = A * w;
    const double f = (xr < 1e-4) ? w * fma(xr, fma(xr, 1.0 / 6.0, -0.5), 1.0)
                                 : (1.0 - exp_fast(-xr)) * rcp_fast(A);
    total += exp_fast(Cq - A * L) * f;
  }
  sumpso_ilai = total * LAI;

  // Pso[60]: mean over [-1 - 1/60, -1] (sailh.py:219)
  const double dx = 1.0 / 60.0;
  const double xc = -1.0 - 0.5 * dx;
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < SPART_NQ1; ++i) {      // the layer is 1/60 wide: a 6-point rule resolves it to 1e-14
    const double x = fma(0.5 * dx, c_gl6_x[i], xc);
    const double arg = fma(A, x, Cq * (1.0 - exp_fast(alpha * x)));
    acc = fma(c_gl6_w[i], exp_fast(arg), acc);
  }
  pso2w = 0.5 * acc;
}

// ---- branch-free variants for the paired-knot path of band_kernel -----------------------------------------
// A band whose centre lies between two grid wavelengths (all MODIS bands, half of the OLCI bands; the
// Sentinel-2 / Landsat-8 tables have integer centres) is evaluated at two knots.  The two evaluations are independent; written as straight-line code without branches they
// sit in one basic block and the compiler interleaves them, which gives every warp two independent FP64 chains
// (FP64 instructions issued back to back by one warp keep the pipe's 2-cycle cadence, those of different warps
// follow each other every 3 cycles, tools/micro/fp64_latency.cu).  MEASURED: no gain (SPART_BAND_PAIR, off by
// default) -- with register operands the pipe runs at ~2.9 cycles per instruction either way.  Each function performs the operations of its
// branching original on the lanes the original would have executed them on (the other form's result is
// computed and dropped by a select), so the results are bit-identical.
__device__ __forceinline__ double plate_tau_nb(double K, const TauTable* tab) {
  // caller guarantees K > 0
  const double emk = exp_neg(-K);
  const double t = rcp_fast(K);
  const bool small = K < 1.0;
  const int e = (__double2hiint(K) >> 20) - 1023;   // floor(log2 K) for K >= 1
  const int idx = small ? 0 : min(e + 1, SPART_TAU_NINT - 1);
  const double u = small ? 2.0 * K - 1.0 : (t - tab->mid[idx]) * tab->invhalf[idx];
  const double p = poly_eval<SPART_TAU_DEG, SPART_TAU_SPLIT>(tab->coef[idx], u);
  const double lg = log_fast(small ? K : 1.0);
  const double e1 = fma(K, p, -0.57721566490153286061 - lg);
  const double r_small = (1.0 - K) * emk + K * K * e1;
  const double r_big = emk * t * p;
  return small ? r_small : r_big;
}

__device__ __forceinline__ void leaf_from_tau_nb(double tau, double t_alph, double t12, double t21, double N,
                                                 double& refl, double& tran) {
  const double r_alph = 1.0 - t_alph, r12 = 1.0 - t12, r21 = 1.0 - t21;
  const double tt21 = tau * t21;
  const double inv_d1 = rcp_fast(1.0 - r21 * r21 * tau * tau);
  const double Ta = t_alph * tt21 * inv_d1;
  const double Ra = r_alph + r21 * tau * Ta;
  const double t = t12 * tt21 * inv_d1;
  const double r = r12 + r21 * tau * t;
  const double Nm1 = N - 1.0;
  const bool zero_abs = r + t >= 1.0;                  // prospect_5d.py:233-235
  const double Tz = t * rcp_fast(t + (1.0 - t) * Nm1);
  const double D = sqrt_fast((1.0 + r + t) * (1.0 + r - t) * (1.0 - r + t) * (1.0 - r - t));
  const double rq = r * r, tq = t * t;
  const double a = (1.0 + rq - tq + D) * rcp_fast(2.0 * r);
  const double b = (1.0 - rq + tq + D) * rcp_fast(2.0 * t);
  const double bNm1 = (Nm1 == 0.0) ? 1.0 : exp_clamp(Nm1 * log_fast(b));
  const double bN2 = bNm1 * bNm1;
  const double a2 = a * a;
  const double inv_d2 = rcp_fast(a2 * bN2 - 1.0);
  const double Rsub = zero_abs ? 1.0 - Tz : a * (bN2 - 1.0) * inv_d2;
  const double Tsub = zero_abs ? Tz : bNm1 * (a2 - 1.0) * inv_d2;
  const double inv_d3 = rcp_fast(1.0 - Rsub * r);
  tran = Ta * Tsub * inv_d3;
  refl = Ra + Ta * Rsub * t * inv_d3;
}

__device__ __forceinline__ void prospect_point_nb(const LeafPar& L, const double* lc, const TauTable* tab,
                                                  double& refl, double& tran) {
  const double Ksum = L.Cab * lc[LC_KAB] + L.Cca * lc[LC_KCA] + L.Cdm * lc[LC_KDM] + L.Cw * lc[LC_KW] +
                      L.Cs * lc[LC_KS] + L.Cant * lc[LC_KANT] + L.CBC * lc[LC_CBC] + L.PROT * lc[LC_PROT];
  const double Kall = Ksum * L.invN;
  const bool pos = Kall > 0.0;
  const double tk = plate_tau_nb(pos ? Kall : 1.0, tab);
  leaf_from_tau_nb(pos ? tk : 1.0, lc[LC_TALPH], lc[LC_T12], lc[LC_T21], L.N, refl, tran);
}

// BSM at two wavelengths (same sample): bsm_point twice, side by side
__device__ __forceinline__ void bsm_pair(const SoilPar& S, const double* lc0, const double* lc1, double& rwet0,
                                         double& rwet1) {
  const double rdry0 = S.f1 * lc0[LC_GSV0] + S.f2 * lc0[LC_GSV1] + S.f3 * lc0[LC_GSV2];
  const double rdry1 = S.f1 * lc1[LC_GSV0] + S.f2 * lc1[LC_GSV1] + S.f3 * lc1[LC_GSV2];
  rwet0 = rdry0;
  rwet1 = rdry1;
  if (S.mu > 0.0) {
    const double rbac0 = 1.0 - (1.0 - rdry0) * (rdry0 * lc0[LC_SOILC1] + 1.0 - rdry0);
    const double rbac1 = 1.0 - (1.0 - rdry1) * (rdry1 * lc1[LC_SOILC1] + 1.0 - rdry1);
    const double p0 = lc0[LC_SOILP], Rw0 = lc0[LC_SOILRW], p1 = lc1[LC_SOILP], Rw1 = lc1[LC_SOILRW];
    const double tw10 = exp_neg(-2.0 * lc0[LC_KW] * S.film), tw11 = exp_neg(-2.0 * lc1[LC_KW] * S.film);
    double fk = S.emu;
    double acc0 = rdry0 * fk, acc1 = rdry1 * fk;
    double tw0 = 1.0, tw1 = 1.0;
    const double g0 = (1.0 - Rw0) * (1.0 - p0), g1 = (1.0 - Rw1) * (1.0 - p1);
#pragma unroll
    for (int k = 1; k <= 6; ++k) {
      tw0 *= tw10;
      tw1 *= tw11;
      fk = fk * S.mu * (1.0 / (double)k);
      const double x0 = tw0 * rbac0, x1 = tw1 * rbac1;
      acc0 += (Rw0 + g0 * x0 * rcp_fast(1.0 - p0 * x0)) * fk;
      acc1 += (Rw1 + g1 * x1 * rcp_fast(1.0 - p1 * x1)) * fk;
    }
    rwet0 = acc0;
    rwet1 = acc1;
  }
}

__device__ __forceinline__ double sail_J1_nb(double m, double k, double LAI, double em, double ek) {
  const bool close = fabs((m - k) * LAI) < 1e-6;
  const double s = 0.5 * (em + ek) * LAI * (1.0 - (1.0 / 12.0) * (k - m) * (k - m) * LAI * LAI);
  const double d = (em - ek) * rcp_fast(k - m);
  return close ? s : d;
}

// ---- SMAC atmosphere at one band (smac.py:94-207) + TOC->TOA (SPART.py:235-252) ---------
// u^n is evaluated as exp_fast(n ln u) with the logarithms taken once per sample; gases whose
// coefficient a is zero for this band contribute exp_fast(0) = 1 and are skipped (warp-uniform
// branch); exp_fast(-taup/aa_i) are products of exp_fast(-taup/us), exp_fast(-taup/uv), exp_fast(+-ak taup).
struct AtmSample {
  double us, uv, m, Peq, lo3, lh2o, lm, lpeq, cksi, ksiD, ray_phase, taup550;
  double inv_us, inv_uv, inv_1pus, inv_1puv, aa3;
};

struct AtmOptics {   // AtmosphericOptics of the reference (smac.py:216-272)
  double Ta_s, Ta_o, Tg, Ra_dd, Ra_so, Ta_ss, Ta_sd, Ta_oo, Ta_do;
};

__device__ __forceinline__ AtmOptics smac_band(const AtmSample& S, const double* c) {
  const double us = S.us, uv = S.uv, m = S.m, Peq = S.Peq, taup550 = S.taup550;
  const double inv_us = S.inv_us, inv_uv = S.inv_uv;
  const double taup = c[SM_A0TAUP] + c[SM_A1TAUP] * taup550;

  // gaseous transmission, smac.py:105-119
  double gsum = 0.0;
  if (nonzero_coef(c[SM_AO3])) gsum += c[SM_AO3] * exp_clamp(c[SM_NO3] * S.lo3);
  if (nonzero_coef(c[SM_AH2O])) gsum += c[SM_AH2O] * exp_clamp(c[SM_NH2O] * S.lh2o);
  if (nonzero_coef(c[SM_AO2])) gsum += c[SM_AO2] * exp_clamp(fma(c[SM_NPO2], S.lpeq, c[SM_NO2] * S.lm));
  if (nonzero_coef(c[SM_ACO2])) gsum += c[SM_ACO2] * exp_clamp(fma(c[SM_NPCO2], S.lpeq, c[SM_NCO2] * S.lm));
  if (nonzero_coef(c[SM_ACH4])) gsum += c[SM_ACH4] * exp_clamp(fma(c[SM_NPCH4], S.lpeq, c[SM_NCH4] * S.lm));
  if (nonzero_coef(c[SM_ANO2])) gsum += c[SM_ANO2] * exp_clamp(fma(c[SM_NPNO2], S.lpeq, c[SM_NNO2] * S.lm));
  if (nonzero_coef(c[SM_ACO])) gsum += c[SM_ACO] * exp_clamp(fma(c[SM_NPCO], S.lpeq, c[SM_NCO] * S.lm));
  const double tg = exp_clamp(gsum);

  const double s = c[SM_A0S] * Peq + c[SM_A3S] + c[SM_A1S] * taup550 + c[SM_A2S] * taup550 * taup550;
  const double tnum = c[SM_A2T] * Peq + c[SM_A3T];
  const double ttetas = c[SM_A0T] + c[SM_A1T] * taup550 * inv_us + tnum * S.inv_1pus;
  const double ttetav = c[SM_A0T] + c[SM_A1T] * taup550 * inv_uv + tnum * S.inv_1puv;

  const double cksi = S.cksi, ksiD = S.ksiD;
  const double taur = c[SM_TAUR];
  const double inv_usuv = inv_us * inv_uv;
  const double rr = taur * S.ray_phase * inv_usuv;
  const double ray_ref = 0.25 * rr * Peq;    // smac.py:142-143
  const double taurz = taur * Peq;

  const double ksi2 = ksiD * ksiD;
  const double aer_phase = c[SM_A0P] + c[SM_A1P] * ksiD + c[SM_A2P] * ksi2 + c[SM_A3P] * (ksi2 * ksiD) +
                           c[SM_A4P] * (ksi2 * ksi2);
  const double wo = c[SM_WO], ak2 = c[SM_AK2], ak = c[SM_AK];
  const double opb = c[SM_OPB], omb = c[SM_OMB], g3 = c[SM_G3], h3 = c[SM_H3], akd3 = c[SM_AKD3];

  const double us2 = us * us;
  const double inv_q = rcp_fast(1.0 - ak2 * us2);
  const double e = -0.75 * us2 * wo * inv_q;
  const double f = -0.25 * h3 * us2 * wo * inv_q;
  const double dp = e * inv_us * (1.0 / 3.0) + us * f;
  const double d = e + f;
  const double eak = exp_clamp(ak * taup);
  const double emak = rcp_fast(eak);
  const double inv_delta = rcp_fast(eak * c[SM_OPB2] - emak * c[SM_OMB2]);
  const double ss = us * inv_q;
  const double q1 = 2.0 + 3.0 * us + h3 * us * (1.0 + 2.0 * us);
  const double q2 = 2.0 - 3.0 * us - h3 * us * (1.0 - 2.0 * us);
  const double Eu = exp_neg(-taup * inv_us), Ev = exp_neg(-taup * inv_uv);
  const double q3 = q2 * Eu;
  const double wsd = c[SM_WW] * ss * inv_delta;
  const double c1 = wsd * (q1 * eak * opb + q3 * omb);
  const double c2 = -wsd * (q1 * emak * omb + q3 * opb);
  const double cp1 = c1 * akd3;
  const double cp2 = -c2 * akd3;
  const double g3uv = g3 * uv;
  const double z = d - g3uv * dp + wo * aer_phase * 0.25;
  const double x = c1 - g3uv * cp1;
  const double y = c2 - g3uv * cp2;
  const double aa1 = uv * rcp_fast(1.0 + ak * uv);
  const double aa2 = uv * rcp_fast(1.0 - ak * uv);
  const double aer_ref1 = x * aa1 * (1.0 - Ev * emak);   // exp_clamp(-taup/aa1) = exp_clamp(-taup/uv - ak taup)
  const double aer_ref2 = y * aa2 * (1.0 - Ev * eak);
  const double aer_ref3 = z * S.aa3 * (1.0 - Ev * Eu);
  const double aer_ref = (aer_ref1 + aer_ref2 + aer_ref3) * inv_usuv;

  const double Res_ray = c[SM_RESR1] + c[SM_RESR2TAUR] * S.ray_phase * inv_usuv + c[SM_RESR3] * (rr * rr);
  const double ta = taup * m * cksi;
  const double Res_aer = (c[SM_RESA1] + c[SM_RESA2] * ta + c[SM_RESA3] * (ta * ta)) + c[SM_RESA4] * (ta * ta * ta);
  const double tautot = taup + taurz;
  const double tt = tautot * m * cksi;
  const double Res_6s = (c[SM_REST1] + c[SM_REST2] * tt + c[SM_REST3] * (tt * tt)) + c[SM_REST4] * (tt * tt * tt);
  const double atm_ref = ray_ref - Res_ray + aer_ref - Res_aer + Res_6s;

  const double ta_ss = exp_neg(-tautot * inv_us);
  const double ta_oo = exp_neg(-tautot * inv_uv);
  AtmOptics O;
  O.Ta_s = ttetas;
  O.Ta_o = ttetav;
  O.Tg = tg;
  O.Ra_dd = s;
  O.Ra_so = atm_ref;
  O.Ta_ss = ta_ss;
  O.Ta_sd = ttetas - ta_ss;
  O.Ta_oo = ta_oo;
  O.Ta_do = ttetav - ta_oo;
  return O;
}

__device__ __forceinline__ void smac_toa_band(const AtmSample& S, const double* c, double conv_ea, double etscale,
                                              double rv_so, double rv_do, double rv_dd, double rv_sd,
                                              double& R_TOC, double& R_TOA, double& L_TOA) {
  const AtmOptics O = smac_band(S, c);
  const double tg = O.Tg, ta_ss = O.Ta_ss, ta_sd = O.Ta_sd, ta_oo = O.Ta_oo, ta_do = O.Ta_do;
  // SPART.py:243-252
  const double ra_dd = O.Ra_dd, ra_so = O.Ra_so;
  const double inv_ms = rcp_fast(1.0 - rv_dd * ra_dd);
  const double rtoa0 = ra_so + ta_ss * rv_so * ta_oo;
  const double rtoa1 = (ta_sd * rv_do + ta_ss * rv_sd * ra_dd * rv_do) * ta_oo * inv_ms;
  const double rtoa2 = (ta_ss * rv_sd + ta_sd * rv_dd) * ta_do * inv_ms;
  R_TOC = (ta_ss * rv_so + ta_sd * rv_do) * rcp_fast(ta_ss + ta_sd);
  R_TOA = tg * (rtoa0 + rtoa1 + rtoa2);
  L_TOA = (conv_ea * etscale) * R_TOA;
}

// ---- SMAC with block-uniform geometry (SPART_FLAG_UNIFORM_GEOMETRY) -------------------------
// When every sample shares the sun / observer angles, all sub-expressions of smac_band() that
// depend only on the geometry and the band coefficients are formed once per block and band
// (smac_fold_geometry, one thread per band) and kept in shared memory; the per-sample work keeps
// only what depends on pressure, aerosol load and the gas columns.  The sums are re-associated
// with respect to smac_band() (a few ulp), which is why this is a separate code path.
enum UgIndex {
  UG_TS0 = 0, UG_TS1, UG_TS2,       // ttetas = TS0 + TS1 taup550 + TS2 Peq
  UG_TV0, UG_TV1, UG_TV2,           // ttetav
  UG_RAYREF,                        // ray_ref / Peq
  UG_RESRAY,                        // Res_ray
  UG_Q1OPB, UG_Q2OMB, UG_Q1OMB, UG_Q2OPB,
  UG_X1W, UG_Y2W, UG_Z3,            // weights of the three aerosol reflectance terms
  UG_MCKSI, UG_INVUS, UG_INVUV,
  UG_GO2, UG_GCO2, UG_GCH4, UG_GNO2, UG_GCO,   // n ln m of the uniformly mixed gases
  UG_COUNT
};

struct AtmGeometry {   // the geometry-only members of AtmSample
  double us, uv, m, lm, cksi, ksiD, ray_phase, inv_us, inv_uv, inv_1pus, inv_1puv, aa3;
};

__device__ __forceinline__ void smac_fold_geometry(const AtmGeometry& S, const double* c, double* u) {
  const double us = S.us, uv = S.uv, inv_us = S.inv_us, inv_uv = S.inv_uv;
  u[UG_TS0] = c[SM_A0T] + c[SM_A3T] * S.inv_1pus;
  u[UG_TS1] = c[SM_A1T] * inv_us;
  u[UG_TS2] = c[SM_A2T] * S.inv_1pus;
  u[UG_TV0] = c[SM_A0T] + c[SM_A3T] * S.inv_1puv;
  u[UG_TV1] = c[SM_A1T] * inv_uv;
  u[UG_TV2] = c[SM_A2T] * S.inv_1puv;
  const double taur = c[SM_TAUR];
  const double inv_usuv = inv_us * inv_uv;
  const double rr = taur * S.ray_phase * inv_usuv;
  u[UG_RAYREF] = 0.25 * rr;
  u[UG_RESRAY] = c[SM_RESR1] + c[SM_RESR2TAUR] * S.ray_phase * inv_usuv + c[SM_RESR3] * (rr * rr);
  const double ksiD = S.ksiD, ksi2 = ksiD * ksiD;
  const double aer_phase = c[SM_A0P] + c[SM_A1P] * ksiD + c[SM_A2P] * ksi2 + c[SM_A3P] * (ksi2 * ksiD) +
                           c[SM_A4P] * (ksi2 * ksi2);
  const double wo = c[SM_WO], ak2 = c[SM_AK2], ak = c[SM_AK];
  const double opb = c[SM_OPB], omb = c[SM_OMB], g3 = c[SM_G3], h3 = c[SM_H3], akd3 = c[SM_AKD3];
  const double us2 = us * us;
  const double inv_q = 1.0 / (1.0 - ak2 * us2);
  const double e = -0.75 * us2 * wo * inv_q;
  const double f = -0.25 * h3 * us2 * wo * inv_q;
  const double dp = e * inv_us * (1.0 / 3.0) + us * f;
  const double d = e + f;
  const double q1 = 2.0 + 3.0 * us + h3 * us * (1.0 + 2.0 * us);
  const double q2 = 2.0 - 3.0 * us - h3 * us * (1.0 - 2.0 * us);
  u[UG_Q1OPB] = q1 * opb;
  u[UG_Q2OMB] = q2 * omb;
  u[UG_Q1OMB] = q1 * omb;
  u[UG_Q2OPB] = q2 * opb;
  const double wss = c[SM_WW] * us * inv_q;
  const double g3uv = g3 * uv;
  const double z = d - g3uv * dp + wo * aer_phase * 0.25;
  const double aa1 = uv / (1.0 + ak * uv);
  const double aa2 = uv / (1.0 - ak * uv);
  u[UG_X1W] = wss * (1.0 - g3uv * akd3) * aa1 * inv_usuv;
  u[UG_Y2W] = wss * (1.0 + g3uv * akd3) * aa2 * inv_usuv;
  u[UG_Z3] = z * S.aa3 * inv_usuv;
  u[UG_MCKSI] = S.m * S.cksi;
  u[UG_INVUS] = inv_us;
  u[UG_INVUV] = inv_uv;
  u[UG_GO2] = c[SM_NO2] * S.lm;
  u[UG_GCO2] = c[SM_NCO2] * S.lm;
  u[UG_GCH4] = c[SM_NCH4] * S.lm;
  u[UG_GNO2] = c[SM_NNO2] * S.lm;
  u[UG_GCO] = c[SM_NCO] * S.lm;
}

struct AtmColumn {   // the per-sample members of AtmSample
  double Peq, lo3, lh2o, lpeq, taup550;
};

__device__ __forceinline__ void smac_toa_band_uniform(const AtmColumn& S, const double* c, const double* u,
                                                      double conv_ea, double etscale, double rv_so, double rv_do,
                                                      double rv_dd, double rv_sd, double& R_TOC, double& R_TOA,
                                                      double& L_TOA) {
  const double Peq = S.Peq, taup550 = S.taup550;
  const double taup = c[SM_A0TAUP] + c[SM_A1TAUP] * taup550;

  double gsum = 0.0;
  if (nonzero_coef(c[SM_AO3])) gsum += c[SM_AO3] * exp_clamp(c[SM_NO3] * S.lo3);
  if (nonzero_coef(c[SM_AH2O])) gsum += c[SM_AH2O] * exp_clamp(c[SM_NH2O] * S.lh2o);
  if (nonzero_coef(c[SM_AO2])) gsum += c[SM_AO2] * exp_clamp(fma(c[SM_NPO2], S.lpeq, u[UG_GO2]));
  if (nonzero_coef(c[SM_ACO2])) gsum += c[SM_ACO2] * exp_clamp(fma(c[SM_NPCO2], S.lpeq, u[UG_GCO2]));
  if (nonzero_coef(c[SM_ACH4])) gsum += c[SM_ACH4] * exp_clamp(fma(c[SM_NPCH4], S.lpeq, u[UG_GCH4]));
  if (nonzero_coef(c[SM_ANO2])) gsum += c[SM_ANO2] * exp_clamp(fma(c[SM_NPNO2], S.lpeq, u[UG_GNO2]));
  if (nonzero_coef(c[SM_ACO])) gsum += c[SM_ACO] * exp_clamp(fma(c[SM_NPCO], S.lpeq, u[UG_GCO]));
  const double tg = exp_clamp(gsum);

  const double ra_dd = c[SM_A0S] * Peq + c[SM_A3S] + c[SM_A1S] * taup550 + c[SM_A2S] * taup550 * taup550;
  const double ttetas = fma(u[UG_TS2], Peq, fma(u[UG_TS1], taup550, u[UG_TS0]));
  const double ttetav = fma(u[UG_TV2], Peq, fma(u[UG_TV1], taup550, u[UG_TV0]));

  const double eak = exp_clamp(c[SM_AK] * taup);
  const double emak = rcp_fast(eak);
  const double inv_delta = rcp_fast(eak * c[SM_OPB2] - emak * c[SM_OMB2]);
  const double Eu = exp_neg(-taup * u[UG_INVUS]), Ev = exp_neg(-taup * u[UG_INVUV]);
  const double c1 = fma(u[UG_Q1OPB], eak, u[UG_Q2OMB] * Eu);
  const double c2 = fma(u[UG_Q1OMB], emak, u[UG_Q2OPB] * Eu);
  const double t1 = u[UG_X1W] * c1 * (1.0 - Ev * emak);
  const double t2 = u[UG_Y2W] * c2 * (1.0 - Ev * eak);
  const double aer_ref = fma(inv_delta, t1 - t2, u[UG_Z3] * (1.0 - Ev * Eu));

  const double mcksi = u[UG_MCKSI];
  const double ta = taup * mcksi;
  const double Res_aer = (c[SM_RESA1] + c[SM_RESA2] * ta + c[SM_RESA3] * (ta * ta)) + c[SM_RESA4] * (ta * ta * ta);
  const double tautot = fma(c[SM_TAUR], Peq, taup);
  const double tt = tautot * mcksi;
  const double Res_6s = (c[SM_REST1] + c[SM_REST2] * tt + c[SM_REST3] * (tt * tt)) + c[SM_REST4] * (tt * tt * tt);
  const double ra_so = u[UG_RAYREF] * Peq - u[UG_RESRAY] + aer_ref - Res_aer + Res_6s;

  const double ta_ss = exp_neg(-tautot * u[UG_INVUS]);
  const double ta_oo = exp_neg(-tautot * u[UG_INVUV]);
  const double ta_sd = ttetas - ta_ss, ta_do = ttetav - ta_oo;

  // SPART.py:243-252
  const double inv_ms = rcp_fast(1.0 - rv_dd * ra_dd);
  const double rtoa0 = ra_so + ta_ss * rv_so * ta_oo;
  const double rtoa1 = (ta_sd * rv_do + ta_ss * rv_sd * ra_dd * rv_do) * ta_oo * inv_ms;
  const double rtoa2 = (ta_ss * rv_sd + ta_sd * rv_dd) * ta_do * inv_ms;
  R_TOC = (ta_ss * rv_so + ta_sd * rv_do) * rcp_fast(ta_ss + ta_sd);
  R_TOA = tg * (rtoa0 + rtoa1 + rtoa2);
  L_TOA = (conv_ea * etscale) * R_TOA;
}

}  // namespace spart
