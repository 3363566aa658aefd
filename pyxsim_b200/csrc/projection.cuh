// projection.cuh -- device code of the photon projection: absorption decisions, position scatter,
// sky transform (shared by the pipeline kernels and the stand-alone natives).
//
// Reference: the projection tile loop, pyxsim/photon_list.py:686-799, calling
//   doppler_shift            pyxsim/lib/sky_functions.pyx:20-34
//   absorb_photons           pyxsim/spectral_models.py:556-583
//   scatter_events           sky_functions.pyx:40-147
//   scatter_events_allsky    sky_functions.pyx:185-250
//   pixel_to_cel             sky_functions.pyx:153-179
//
// Random numbers: photons 2m and 2m+1 of a cell (global id gid) share their Philox blocks:
//   STREAM_ENERGY (gid, m): words (x,y) / (z,w)  energy of photon 2m / 2m+1
//   STREAM_PROJ_A (gid, m): x / y  absorption uniform (top 32 bits) of photon 2m / 2m+1,
//                           z / w  third scatter word of photon 2m / 2m+1
//   STREAM_PROJ_B (gid, m): (x,y) / (z,w)  first two scatter words of photon 2m / 2m+1
//   STREAM_PROJ_D (gid, m): x / y  internal-absorption uniform of photon 2m / 2m+1
//   STREAM_PROJ_C (gid, d): per photon: sigma_pos smoothing, all-sky particle kernels
//   STREAM_PROJ_R (gid, d): per photon, drawn only when needed: the low 32 bits of an absorption
//                           uniform whose top 32 bits do not decide (probability ~2^-32 per test)
// so the events of a cell are a pure function of (seed, gid) -- independent of chunking, of the
// number of ranks and of whether generation and projection run fused or as two stages.
#pragma once
#include "common.cuh"

constexpr double PXB_C_KMS = 299792.458;          // clight in km/s (utils.py:10-11)
constexpr double PXB_PI = 3.14159265358979323846;
constexpr double PXB_RAD2DEG = 180.0 / PXB_PI;

// ---------------------------------------------------------------------------------------
// absorption: transmission(E) = exp(-sigma(E) * nH)
// ---------------------------------------------------------------------------------------
// Morrison & McCammon (1983) piecewise polynomial, as used by soxs.get_wabs_absorb (the
// third-party routine behind WabsModel, spectral_models.py:633-635): sigma(E) =
// (c0 + c1 E + c2 E^2) / E^3 * 1e-24 cm^2.
// Coefficients live in global memory (L1-resident, 16-byte records) rather than __constant__:
// the piece index differs per lane, and divergent constant-bank reads serialise.
struct __align__(16) WabsPiece { double c0, c1, c2, pad; };
__device__ const WabsPiece g_wabs[14] = {
    {17.3, 608.1, -2150.0, 0.0},  {34.6, 267.9, -476.1, 0.0}, {78.1, 18.8, 4.3, 0.0},
    {71.4, 66.8, -51.4, 0.0},     {95.5, 145.8, -61.1, 0.0},  {308.9, -380.6, 294.0, 0.0},
    {120.6, 169.3, -47.7, 0.0},   {141.3, 146.8, -31.5, 0.0}, {202.7, 104.7, -17.0, 0.0},
    {342.7, 18.7, 0.0, 0.0},      {352.2, 18.7, 0.0, 0.0},    {433.9, -2.4, 0.75, 0.0},
    {629.0, 30.9, 0.0, 0.0},      {701.2, 25.2, 0.0, 0.0}};

__device__ __forceinline__ double wabs_tau_per_nh(double e) {
  // piece index = number of inner edges <= e (edges 0.1 ... 8.331 keV; below 0.1 the first
  // piece, above 10 the last): thirteen independent compares against immediates
  const int k = (e >= 0.1) + (e >= 0.284) + (e >= 0.4) + (e >= 0.532) + (e >= 0.707) +
                (e >= 0.867) + (e >= 1.303) + (e >= 1.84) + (e >= 2.471) + (e >= 3.21) +
                (e >= 4.038) + (e >= 7.111) + (e >= 8.331);
  const double2* q = reinterpret_cast<const double2*>(&g_wabs[k]);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  const double inv = 1.0 / e;
  const double s = (a.x + e * (a.y + e * b.x)) * (inv * inv * inv);
  return s * 1.0e-2;  // 1e-24 cm^2 * 1e22 cm^-2
}

// np.interp(e, emid, sigma, left=0, right=0): spectral_models.py:560
__device__ __forceinline__ double table_tau_per_nh(const pxb_project_params& P, double e) {
  const int64_t n = P.abs_n;
  const double* __restrict__ xs = P.abs_energy;
  const double* __restrict__ ys = P.abs_sigma;
  if (n < 1 || !(e >= __ldg(xs)) || !(e <= __ldg(xs + n - 1))) return 0.0;
  int64_t lo;
  if (P.abs_grid == 0) {
    lo = 0;
    int64_t hi = n - 1;  // xs[lo] <= e <= xs[hi]
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(xs + mid) <= e) lo = mid; else hi = mid;
    }
  } else {
    // log-spaced nodes: a single-precision logarithm is enough for the GUESS of the interval
    // (the fp64 log10 cost 75 instructions per photon); the stored nodes decide
    const double xx = (P.abs_grid == 2) ? (double)(__log2f((float)e) * 0.30102999566398120f) : e;
    lo = (int64_t)((xx - P.abs_x0) * P.abs_inv_dx);
    lo = lo < 0 ? 0 : (lo > n - 2 ? n - 2 : lo);
    // the analytic index may be off by a node or two: fix up against the stored nodes
    while (lo > 0 && __ldg(xs + lo) > e) --lo;
    while (lo < n - 2 && __ldg(xs + lo + 1) <= e) ++lo;
  }
  if (n == 1) return __ldg(ys);
  const double x0 = __ldg(xs + lo), x1 = __ldg(xs + lo + 1);
  const double y0 = __ldg(ys + lo), y1 = __ldg(ys + lo + 1);
  return y0 + (y1 - y0) / (x1 - x0) * (e - x0);
}

__device__ __forceinline__ double tau_per_nh(const pxb_project_params& P, double e) {
  return P.absorb_kind == PXB_ABS_WABS ? wabs_tau_per_nh(e) : table_tau_per_nh(P, e);
}

__device__ __forceinline__ double doppler_factor(double beta_n, double beta2) {
  return (1.0 - beta_n) / sqrt(1.0 - beta2);  // sky_functions.pyx:31
}

// inverse gnomonic projection; closed form of sky_functions.pyx:167-179
// (xi = -x, eta = y: RA = RA0 + atan2(xi, cos d0 - eta sin d0),
//  Dec = atan2(sin d0 + eta cos d0, hypot(xi, cos d0 - eta sin d0)))
__device__ __forceinline__ void tan_to_cel(double x, double y, double ra0_rad, double sin_d0,
                                           double cos_d0, double& ra_deg, double& dec_deg) {
  const double den = cos_d0 - y * sin_d0;
  const double num = sin_d0 + y * cos_d0;
  ra_deg = (ra0_rad + atan2(-x, den)) * PXB_RAD2DEG;
  dec_deg = atan2(num, sqrt(x * x + den * den)) * PXB_RAD2DEG;
}

// atan(t) for small |t| by its Maclaurin series (terms to t^15: < 1 ulp for |t| <= 1/16)
__device__ __forceinline__ double atan_small(double t) {
  const double z = t * t;
  double p = -1.0 / 15.0;
  p = p * z + 1.0 / 13.0;
  p = p * z - 1.0 / 11.0;
  p = p * z + 1.0 / 9.0;
  p = p * z - 1.0 / 7.0;
  p = p * z + 1.0 / 5.0;
  p = p * z - 1.0 / 3.0;
  return t + t * z * p;
}

// the same series cut after t^9: exact to fp64 rounding for |t| <= 1/64
__device__ __forceinline__ double atan_tiny(double t) {
  const double z = t * t;
  double p = 1.0 / 9.0;
  p = p * z - 1.0 / 7.0;
  p = p * z + 1.0 / 5.0;
  p = p * z - 1.0 / 3.0;
  return t + t * z * p;
}

// inverse gnomonic projection for the fused kernel.  Same map as tan_to_cel, arranged so that
// both angles are small-argument arctangents (no atan2, no cancellation):
//   RA  - RA0  = atan(-x / den),                      den = cos d0 - y sin d0
//   Dec - Dec0 = atan((y - s sin d0) / (1 + s cos d0)), s = x^2 / (hypot(x, den) + den)
// Wide fields (|arg| > 1/16 or den <= 0) take the general formula.
__device__ __forceinline__ void tan_to_cel_fast(double x, double y, double ra0_rad, double d0_rad,
                                                double sin_d0, double cos_d0, double& ra_deg,
                                                double& dec_deg) {
  const double den = cos_d0 - y * sin_d0;
  {
    // Within ~0.9 degrees of the centre: one reciprocal, no square root.  With r = 1/den,
    // t1 = -x r and q = t1^2:  s = hypot(x, den) - den = (x^2 r / 2)(1 - q/4 + q^2/8 - 5q^3/64 +
    // 7q^4/128) and 1/(1 + s cos d0) by its geometric series (s <= 1.3e-4); both truncation errors
    // are below 2e-16 relative, the arctangent series can stop at t^9 (next term < 8e-20).
    const double r = 1.0 / den;
    const double t1 = -x * r;
    if (den > 0.0 && fabs(t1) <= 0.015625 && fabs(y) <= 0.015625) {
      const double q = t1 * t1;
      const double s = 0.5 * x * x * r * (1.0 + q * (-0.25 + q * (0.125 + q * (-5.0 / 64.0 + q * (7.0 / 128.0)))));
      const double e = s * cos_d0;
      const double t2 = (y - s * sin_d0) * (1.0 + e * (-1.0 + e * (1.0 + e * (-1.0 + e))));
      if (fabs(t2) <= 0.015625) {
        ra_deg = (ra0_rad + atan_tiny(t1)) * PXB_RAD2DEG;
        dec_deg = (d0_rad + atan_tiny(t2)) * PXB_RAD2DEG;
        return;
      }
    }
  }
  const double h = sqrt(x * x + den * den);
  // three quotients x^2/(h+den), -x/den, num2/den2 share one reciprocal
  const double hd = h + den;
  const double den2 = hd + x * x * cos_d0;          // (1 + s cos d0) * (h + den)
  const double num2 = y * hd - x * x * sin_d0;      // (y - s sin d0) * (h + den)
  const double inv = 1.0 / (den * den2);
  const double t1 = -x * den2 * inv;
  const double t2 = num2 * den * inv;
  if (den > 0.0 && fabs(t1) <= 0.0625 && fabs(t2) <= 0.0625) {
    ra_deg = (ra0_rad + atan_small(t1)) * PXB_RAD2DEG;
    dec_deg = (d0_rad + atan_small(t2)) * PXB_RAD2DEG;
  } else {
    tan_to_cel(x, y, ra0_rad, sin_d0, cos_d0, ra_deg, dec_deg);
  }
}

// u < exp(-tau) without evaluating exp when the bounds 1 - tau <= exp(-tau) <= 1 - tau + tau^2/2
// already decide (tau >= 0)
__device__ __forceinline__ bool survives(double u, double tau) {
  // partial sums of the alternating series bracket exp(-tau) for tau >= 0:
  //   S3 = 1 - t + t^2/2 - t^3/6 <= exp(-t) <= S3 + t^4/24
  if (tau < 1.0) {
    const double s3 = 1.0 - tau * (1.0 - tau * (0.5 - tau * (1.0 / 6.0)));
    if (u < s3) return true;
    const double t2 = tau * tau;
    if (u >= s3 + t2 * t2 * (1.0 / 24.0)) return false;
  } else if (tau < 80.0) {
    // single-precision estimate with a safety margin decides all but ~1e-5 of the cases
    const double p32 = (double)__expf(-(float)tau);
    if (u < p32 * (1.0 - 1.0e-5)) return true;
    if (u > p32 * (1.0 + 1.0e-5)) return false;
  }
  return u < exp(-tau);
}

// ---- single-precision first look at the wabs decision ------------------------------------
// u < exp(-sigma(E) nH) is decided in fp32 whenever the fp32 estimate is further from u than a
// rigorous error bound; only the remaining ~1e-5 of the photons evaluate the fp64 formula.  The
// decision is therefore the fp64 one in every case -- this is a speed-up, not an approximation.
//  * piece index: rounding to float is monotone, so (float)E vs (float)edge orders like E vs
//    edge unless the two floats are EQUAL (then fp64 decides).  The index comes from a
//    128-entry table over [2^-4, 2^4) keV in steps of 1/16 octave (float bits >> 19); every
//    sub-interval holds at most one edge and no edge is a sub-interval boundary (checked by
//    tests/test_cpu_abi.py against the edge list).
//  * tau32: float coefficients, Horner, approximate reciprocal: relative error < 4e-6
//    (condition number of the piece polynomials <= 5); __expf: 2^-22 + 1.2e-7 |x| relative.
//    Bound used: p32 * (2e-5 tau + 1e-6), i.e. > 1.6x the worst case, plus 2e-7 for (float)u.
constexpr int WABS_LUT_N = 128;
constexpr int WABS_LUT_BASE = (127 - 4) << 4;   // float bits >> 19 of 2^-4
struct WabsFast {
  float2 lut[WABS_LUT_N];   // x: the edge inside the sub-interval (+inf: none); y: piece index bits
  float4 coef[14];          // c0, c1, c2 of the piece in float
};

__device__ __forceinline__ void wabs_fast_init(WabsFast& F) {
  const double edges[13] = {0.1, 0.284, 0.4, 0.532, 0.707, 0.867, 1.303, 1.84, 2.471, 3.21,
                            4.038, 7.111, 8.331};
  for (int i = threadIdx.x; i < WABS_LUT_N; i += blockDim.x) {
    const float lo = __int_as_float((WABS_LUT_BASE + i) << 19);
    const float hi = __int_as_float((WABS_LUT_BASE + i + 1) << 19);
    int k = 0;
    float inside = __int_as_float(0x7f800000);
#pragma unroll
    for (int j = 0; j < 13; ++j) {
      const float ef = (float)edges[j];
      if (ef <= lo) ++k;
      else if (ef < hi) inside = ef;
    }
    F.lut[i] = make_float2(inside, __int_as_float(k));
  }
  for (int i = threadIdx.x; i < 14; i += blockDim.x)
    F.coef[i] = make_float4((float)g_wabs[i].c0, (float)g_wabs[i].c1, (float)g_wabs[i].c2, 0.0f);
}

// u < exp(-wabs_tau_per_nh(e) * nH), decided in fp32 when that is safe
__device__ __forceinline__ bool wabs_survives(const WabsFast& F, double u, double e, double nH) {
  const float ef = (float)e;
  int idx = (__float_as_int(ef) >> 19) - WABS_LUT_BASE;
  idx = max(0, min(idx, WABS_LUT_N - 1));
  const float2 le = F.lut[idx];
  if (ef != le.x) {   // (false for NaN too: falls through to fp64)
    const int k = __float_as_int(le.y) + (ef > le.x ? 1 : 0);
    const float4 c = F.coef[k];
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(ef));
    const float tau = (c.x + ef * (c.y + ef * c.z)) * (inv * inv * inv) * (1.0e-2f * (float)nH);
    const float p = __expf(-tau);
    const float m = 2.0e-5f * tau + 1.2e-6f;
    const float uf = (float)u;
    // + 1e-37 absolute: below 2^-126 the exponential (and uf) are subnormal and only
    // absolutely accurate; for normal p the extra 1e-37 is far inside the relative slack
    if (uf < p * (1.0f - m) - 1.0e-37f) return true;
    if (uf > p * (1.0f + m) + 1.0e-37f) return false;
  }
  return survives(u, wabs_tau_per_nh(e) * nH);
}

// ---- decisions from random WORDS ---------------------------------------------------------
// The absorption uniform is U = (w + w2 / 2^32) / 2^32 with 64 random bits; only the top word w
// is drawn up front.  U lies in [w, w+1) / 2^32: when both ends fall on the same side of the
// transmission the decision is made; otherwise (probability 2^-32) the low word w2 is drawn from
// the photon's own refinement stream and U < p is evaluated in full.  Decisions are therefore
// those of a 64-bit uniform -- exact, at the cost of one 32-bit word per photon.
__device__ __forceinline__ bool survives_w(uint32_t w, double tau, const PhiloxKeys& RK,
                                           uint64_t gid, uint64_t d) {
  const double ulo = (double)w * 2.3283064365386963e-10;       // w / 2^32, exact
  const double uhi = ulo + 2.3283064365386963e-10;             // (w + 1) / 2^32, exact
  // cheap rigorous bounds plo <= exp(-tau) <= phi first (tau >= 0)
  double plo = 0.0, phi = 1.0;
  if (tau < 1.0) {
    plo = 1.0 - tau * (1.0 - tau * (0.5 - tau * (1.0 / 6.0)));
    const double t2 = tau * tau;
    phi = plo + t2 * t2 * (1.0 / 24.0);
  } else if (tau < 80.0) {
    const double p32 = (double)__expf(-(float)tau);
    plo = p32 * (1.0 - 1.0e-5);
    phi = p32 * (1.0 + 1.0e-5);
  }
  if (uhi <= plo) return true;
  if (ulo >= phi) return false;
  const double p = exp(-tau);
  if (uhi <= p) return true;
  if (ulo >= p) return false;
  const Philox4 rr = pxb_rand4(RK, STREAM_PROJ_R, gid, d);
  return fma((double)rr.x, 5.421010862427522e-20, ulo) < p;    // w2 / 2^64 + w / 2^32
}

// wabs: the same decision with a single-precision first look (see WabsFast above)
__device__ __forceinline__ bool wabs_survives_w(const WabsFast& F, uint32_t w, double e, double nH,
                                                const PhiloxKeys& RK, uint64_t gid, uint64_t d) {
  const float ef = (float)e;
  int idx = (__float_as_int(ef) >> 19) - WABS_LUT_BASE;
  idx = max(0, min(idx, WABS_LUT_N - 1));
  const float2 le = F.lut[idx];
  if (ef != le.x) {   // (false for NaN too: falls through to fp64)
    const int k = __float_as_int(le.y) + (ef > le.x ? 1 : 0);
    const float4 c = F.coef[k];
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(ef));
    const float tau = (c.x + ef * (c.y + ef * c.z)) * (inv * inv * inv) * (1.0e-2f * (float)nH);
    const float p = __expf(-tau);
    const float m = 2.0e-5f * tau + 1.2e-6f;
    // uf <= w / 2^32 (conversion rounds toward zero, off by < 2^-23 relative); U < w/2^32 + 2^-32
    const float uf = __uint2float_rz(w) * 2.3283064e-10f;
    if (fmaf(uf, 1.0f + 3.6e-7f, 2.4e-10f) < p * (1.0f - m) - 1.0e-37f) return true;
    if (uf > p * (1.0f + m) + 1.0e-37f) return false;
  }
  return survives_w(w, wabs_tau_per_nh(e) * nH, RK, gid, d);
}

struct SkyConst {
  double sin_d0, cos_d0, ra0_rad, d0_rad, inv_DA;
};

static inline SkyConst make_sky_const(const pxb_project_params* P) {
  SkyConst K;
  K.d0_rad = P->sky_center[1] * (PXB_PI / 180.0);
  K.sin_d0 = sin(K.d0_rad);
  K.cos_d0 = cos(K.d0_rad);
  K.ra0_rad = P->sky_center[0] * (PXB_PI / 180.0);
  K.inv_DA = P->D_A_kpc != 0.0 ? 1.0 / P->D_A_kpc : 0.0;
  return K;
}

__device__ __forceinline__ double u32_centered(uint32_t w) {
  // (w + 0.5) / 2^32 - 0.5: uniform in (-1/2, 1/2) with 32-bit resolution.  w goes into the top
  // mantissa bits of a double in [1, 2) (= 1 + w/2^32, exact); one exact addition of
  // -1.5 + 2^-33 finishes it -- a single fp64 instruction.
  const double d = __hiloint2double((int)(0x3FF00000u | (w >> 12)), (int)(w << 20));
  return d + (-1.5 + 1.0 / 8589934592.0);
}
// w / 2^32 in [0, 1), exact
__device__ __forceinline__ double u32_unit(uint32_t w) {
  return __hiloint2double((int)(0x3FF00000u | (w >> 12)), (int)(w << 20)) - 1.0;
}
// (w0 2^32 + w2 + 1) / 2^64 in (0, 1]  (safe for log; 64-bit resolution near zero)
__device__ __forceinline__ double u64_open0(uint32_t w0, uint32_t w2) {
  return fma((double)w2 + 1.0, 5.421010862427522e-20, u32_unit(w0));
}

// what the scatter step needs of a cell, one 32-byte sector per cell
struct __align__(32) CellGeom {
  double cx;  // centre . x_hat
  double cy;  // centre . y_hat
  double cz;  // centre . z_hat (the reference's `los`; also the 3rd all-sky coordinate)
  double w;   // scatter width: dx (cells), dx/2 (particles, external observer)
};

__device__ __forceinline__ CellGeom ld_geom(const CellGeom* p) {
  const double2* q = reinterpret_cast<const double2*>(p);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  return CellGeom{a.x, a.y, b.x, b.y};
}

// scatter + sky transform of one surviving photon (external observer) from its random words
// (w0, w1: STREAM_PROJ_B pair; w2: the photon's third word of STREAM_PROJ_A).  SPEC = 0 reads
// every option from the parameter block; SPEC = 1 / 2 are the common case compiled with its
// options fixed (cells, TAN sky, no sigma_pos; on-axis / off-axis normal).
template <int SPEC>
__device__ __forceinline__ void scatter_external(const pxb_project_params& P, const PhiloxKeys& RK,
                                                 const CellGeom& g, const SkyConst& K, uint32_t w0,
                                                 uint32_t w1, uint32_t w2, uint64_t gid, uint64_t d,
                                                 double& xs_out, double& ys_out, double& los_out) {
  const double w = g.w, cx = g.cx, cy = g.cy;
  los_out = g.cz;  // from the unscattered centre (sky_functions.pyx:124,143)
  double ox, oy;   // offsets in the sky plane in units of w
  if (SPEC != 0 || P.data_type == 0) {
    // cells: uniform inside the cube, rotated into the sky plane (sky_functions.pyx:74-76,
    // 110-121; on-axis x_hat/y_hat are axis vectors so only two of the draws contribute)
    const double u0 = u32_centered(w0), u1 = u32_centered(w1);
    if (SPEC == 1 || (SPEC == 0 && P.vn_axis >= 0)) {
      ox = u0;
      oy = u1;
    } else {
      const double u2 = u32_centered(w2);
      ox = u0 * P.x_hat[0] + u1 * P.x_hat[1] + u2 * P.x_hat[2];
      oy = u0 * P.y_hat[0] + u1 * P.y_hat[1] + u2 * P.y_hat[2];
    }
  } else if (P.kernel == 1) {
    // two standard normals (Box-Muller), sky_functions.pyx:78-80.  The deviates are evaluated
    // in single precision (2e-7 relative, like cuRAND's float normals; the 64-bit uniform keeps
    // the tail out to 9 sigma): they are a random offset inside the smoothing kernel, compared
    // with the reference in distribution only -- the fp64 logarithm, square root and sincospi
    // were 160 of the 790 instructions per photon of the SPH projection.  Centre + offset stays fp64.
    const float rad = sqrtf(-2.0f * logf((float)u64_open0(w0, w2)));
    float sn, co;
    sincospif(2.0f * (float)u32_unit(w1), &sn, &co);
    ox = (double)(rad * co);
    oy = (double)(rad * sn);
  } else {
    // top_hat: r ~ U(0,1) (NOT sqrt(U)), theta = 2 pi U  (sky_functions.pyx:81-85)
    const double r = fma((double)w2, 5.421010862427522e-20, u32_unit(w0));
    double sn, co;
    sincospi(2.0 * u32_unit(w1), &sn, &co);
    ox = r * co;
    oy = r * sn;
  }
  double xs = cx + ox * w, ys = cy + oy * w;
  if (SPEC == 0 && P.has_sigma_pos && P.data_type == 0) {
    // photon_list.py:753-756: sigma = sigma_pos * dx
    const Philox4 rc = pxb_rand4(RK, STREAM_PROJ_C, gid, d);
    double n0, n1;
    normal2(rc, n0, n1);
    const double sg = P.sigma_pos * w;
    xs += sg * n0;
    ys += sg * n1;
  }
  if (SPEC != 0 || P.sky_mode != PXB_SKY_PHYS) {   // (cx, cy, w are already in radians)
    if (SPEC == 0 && P.sky_mode == PXB_SKY_FLAT) {
      xs = P.sky_center[0] - xs * PXB_RAD2DEG;   // photon_list.py:762-764
      ys = P.sky_center[1] + ys * PXB_RAD2DEG;
    } else {
      double ra, dec;
      tan_to_cel_fast(xs, ys, K.ra0_rad, K.d0_rad, K.sin_d0, K.cos_d0, ra, dec);
      xs = ra;
      ys = dec;
    }
  }
  xs_out = xs;
  ys_out = ys;
}

// internal observer: sky_functions.pyx:205-248
__device__ __forceinline__ void scatter_allsky(const pxb_project_params& P, const PhiloxKeys& RK,
                                               const CellGeom& g, uint32_t w0, uint32_t w1,
                                               uint32_t w2, uint64_t gid, uint64_t d,
                                               double& lon_out, double& lat_out, double& los_out) {
  const double w = g.w, cx = g.cx, cy = g.cy, cz = g.cz;
  double o0, o1, o2;
  if (P.data_type == 0) {
    o0 = u32_centered(w0);
    o1 = u32_centered(w1);
    o2 = u32_centered(w2);
  } else if (P.kernel == 1) {
    const Philox4 rc = pxb_rand4(RK, STREAM_PROJ_C, gid, d);
    const double rad = sqrt(-2.0 * log(u64_open0(w0, w2)));
    double sn, co, dump;
    sincospi(2.0 * u32_unit(w1), &sn, &co);
    o0 = rad * co;
    o1 = rad * sn;
    normal2(rc, o2, dump);
  } else {
    const Philox4 rc = pxb_rand4(RK, STREAM_PROJ_C, gid, d);
    const double r = u53(rc.x, rc.y);
    double sp, cp;
    sincospi(2.0 * u32_unit(w0), &sp, &cp);
    const double ct = 2.0 * fma((double)w2, 5.421010862427522e-20, u32_unit(w1)) - 1.0;  // cos(theta), theta = arccos(U(-1,1))
    const double stheta = sqrt(fmax(0.0, 1.0 - ct * ct));
    o0 = r * cp * stheta;
    o1 = r * sp * stheta;
    o2 = r * ct;
  }
  // the reference scatters in the dataset frame and then rotates; rotation is linear
  const double px = w * o0, py = w * o1, pz = w * o2;
  const double xx = cx + px * P.x_hat[0] + py * P.x_hat[1] + pz * P.x_hat[2];
  const double yy = cy + px * P.y_hat[0] + py * P.y_hat[1] + pz * P.y_hat[2];
  const double zz = cz + px * P.z_hat[0] + py * P.z_hat[1] + pz * P.z_hat[2];
  const double rr = sqrt(xx * xx + yy * yy + zz * zz);
  double lon = atan2(yy, xx) - 0.5 * PXB_PI;
  if (lon < 0.0) lon += 2.0 * PXB_PI;
  if (lon > 2.0 * PXB_PI) lon -= 2.0 * PXB_PI;
  const double lat = 0.5 * PXB_PI - acos(zz / rr);
  lon_out = lon * PXB_RAD2DEG;
  lat_out = lat * PXB_RAD2DEG;
  los_out = rr;
}

// absorption decision of one photon from its words.  wa: top word of the foreground uniform,
// wd: top word of the internal-absorption uniform (used only with per-cell columns).
// SPEC != 0: the common case compiled with its options fixed (wabs, a foreground column only).
template <int SPEC>
__device__ __forceinline__ bool absorb_decision(const pxb_project_params& P, const PhiloxKeys& RK,
                                                const WabsFast& F, int64_t c, uint64_t gid, uint64_t d,
                                                uint32_t wa, uint32_t wd, double en) {
  if (SPEC != 0) return wabs_survives_w(F, wa, en, P.nH, RK, gid, d);
  bool det = true;
  if (P.nH_cell) {
    // internal absorption on the source-frame energy, photon_list.py:715-723.  (Its refinement
    // word, if ever needed, comes from draw index d + 2^62 of the refinement stream so the two
    // tests of a photon never share it.)
    const double es = en * P.oneplusz, col = __ldg(P.nH_cell + c);
    const uint64_t dr = d | (1ull << 62);
    det = P.absorb_kind == PXB_ABS_WABS ? wabs_survives_w(F, wd, es, col, RK, gid, dr)
                                        : survives_w(wd, table_tau_per_nh(P, es) * col, RK, gid, dr);
  }
  if (det && P.nH > 0.0)
    det = P.absorb_kind == PXB_ABS_WABS ? wabs_survives_w(F, wa, en, P.nH, RK, gid, d)
                                        : survives_w(wa, table_tau_per_nh(P, en) * P.nH, RK, gid, d);
  return det;
}

// decision for caller-supplied double uniforms (the diagnostic pxb_absorb_decide)
__device__ __forceinline__ bool absorb_one(const pxb_project_params& P, const WabsFast& F, double u,
                                           double e, double nH) {
  if (P.absorb_kind == PXB_ABS_WABS) return wabs_survives(F, u, e, nH);
  return survives(u, table_tau_per_nh(P, e) * nH);
}

// host: argument check of the absorption part of a parameter block
static inline int pxb_check_project_params(const pxb_project_params* P) {
  PXB_REQUIRE(P, "NULL params");
  PXB_REQUIRE(P->absorb_kind >= 0 && P->absorb_kind <= 2, "bad absorb_kind");
  if (P->absorb_kind == PXB_ABS_TABLE)
    PXB_REQUIRE(P->abs_energy && P->abs_sigma && P->abs_n >= 1, "absorption table missing");
  return PXB_OK;
}
