// project.cu -- photon projection: Doppler shift, foreground/internal absorption rejection,
// position scatter, sky transform and stable stream compaction in ONE pass over the photons.
//
// Reference path: _project_photons tile loop, pyxsim/photon_list.py:661-808, calling
//   doppler_shift            pyxsim/lib/sky_functions.pyx:20-34
//   absorb_photons           pyxsim/spectral_models.py:556-583
//   scatter_events           sky_functions.pyx:40-147
//   scatter_events_allsky    sky_functions.pyx:185-250
//   pixel_to_cel             sky_functions.pyx:153-179
//
// Layout: a small per-cell pre-pass (k_project_cells) turns the 7 per-cell arrays into the 5
// derived values every photon of that cell needs (Doppler factor, rotated centre, scatter
// width).  The photon pass (k_project_photons) owns tiles of TILE photons; tiles take a
// ticket, compact their survivors with ballots + a decoupled look-back over per-tile counts,
// so the event list keeps the photon-list order like the reference's cursor `k`
// (sky_functions.pyx:87-105) and is bit-reproducible.
#include "tiles.cuh"
#include <stdlib.h>
#include <string.h>

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
    const double xx = (P.abs_grid == 2) ? log10(e) : e;
    lo = (int64_t)((xx - P.abs_x0) * P.abs_inv_dx);
    lo = lo < 0 ? 0 : (lo > n - 2 ? n - 2 : lo);
    // the analytic index may be off by one ulp-wise: fix up against the stored nodes
    if (lo > 0 && __ldg(xs + lo) > e) --lo;
    else if (lo < n - 2 && __ldg(xs + lo + 1) <= e) ++lo;
  }
  if (n == 1) return __ldg(ys);
  const double x0 = __ldg(xs + lo), x1 = __ldg(xs + lo + 1);
  const double y0 = __ldg(ys + lo), y1 = __ldg(ys + lo + 1);
  return y0 + (y1 - y0) / (x1 - x0) * (e - x0);
}

__device__ __forceinline__ double tau_per_nh(const pxb_project_params& P, double e) {
  return P.absorb_kind == PXB_ABS_WABS ? wabs_tau_per_nh(e) : table_tau_per_nh(P, e);
}

__global__ void __launch_bounds__(256)
k_absorb_transmission(const __grid_constant__ pxb_project_params P, int64_t n,
                      const double* __restrict__ e, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (P.absorb_kind == PXB_ABS_NONE) ? 1.0 : exp(-tau_per_nh(P, e[i]) * P.nH);
}

// ---------------------------------------------------------------------------------------
// stand-alone natives (deterministic parity against the reference Cython)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double doppler_factor(double beta_n, double beta2) {
  return (1.0 - beta_n) / sqrt(1.0 - beta2);  // sky_functions.pyx:31
}

__global__ void __launch_bounds__(256)
k_doppler_shift(int64_t n_cells, const double* __restrict__ beta_n,
                const double* __restrict__ beta2, const int64_t* __restrict__ poff,
                int64_t n_photons, double* __restrict__ eobs) {
  constexpr int TILE = 1024;
  __shared__ int64_t s_off[TILE + 1];
  __shared__ int64_t s_c[2];
  const int64_t p0 = (int64_t)blockIdx.x * TILE;
  const int64_t p1 = min(p0 + (int64_t)TILE, n_photons);
  TileMap<TILE> tm;
  tm.build(poff, n_cells, p0, p1, s_off, s_c);
  for (int64_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
    const int64_t c = tm.cell_of(p, s_off);
    eobs[p] = eobs[p] * doppler_factor(__ldg(beta_n + c), __ldg(beta2 + c));
  }
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

__global__ void __launch_bounds__(256)
k_pixel_to_cel(int64_t n, double* __restrict__ xs, double* __restrict__ ys, double ra0_rad,
               double sin_d0, double cos_d0) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double ra, dec;
  tan_to_cel(xs[i], ys[i], ra0_rad, sin_d0, cos_d0, ra, dec);
  xs[i] = ra;
  ys[i] = dec;
}

// ---------------------------------------------------------------------------------------
// per-cell pre-pass
// ---------------------------------------------------------------------------------------
constexpr int PROJ_WTILE_C = 256;   // photons per warp tile (== PROJ_WTILE below)
constexpr int PROJ_STAGE_CELLS = 32;  // a tile's cells are staged in shared memory up to this many

// what the scatter step needs of a cell, one 32-byte sector per cell
struct __align__(32) CellGeom {
  double cx;  // centre . x_hat
  double cy;  // centre . y_hat
  double cz;  // centre . z_hat (the reference's `los`; also the 3rd all-sky coordinate)
  double w;   // scatter width: dx (cells), dx/2 (particles, external observer)
};

struct CellDerived {
  double* shift;       // Doppler factor (1 when no_shifting)
  CellGeom* geom;
  int64_t* tile_cell;  // [n_wtiles + 1]: cell holding the first photon of every warp tile
  int64_t* gid;        // global cell id: the RNG counter key (photon_list.cell_id, or cell_id0 + c)
};

__device__ __forceinline__ CellGeom ld_geom(const CellGeom* p) {
  const double2* q = reinterpret_cast<const double2*>(p);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  return CellGeom{a.x, a.y, b.x, b.y};
}

__global__ void __launch_bounds__(256)
k_project_cells(const __grid_constant__ pxb_photon_list L,
                const __grid_constant__ pxb_project_params P, CellDerived D) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= L.n_cells) return;
  const double x = L.x[c], y = L.y[c], z = L.z[c];
  double shift = 1.0;
  if (!P.no_shifting) {
    const double vx = L.vx[c], vy = L.vy[c], vz = L.vz[c];
    double vn;
    if (P.observer == 1) {
      vn = -(vx * x + vy * y + vz * z) / sqrt(x * x + y * y + z * z);  // photon_list.py:688-697
    } else if (P.vn_axis >= 0) {
      vn = P.vn_axis == 0 ? vx : (P.vn_axis == 1 ? vy : vz);
    } else {
      vn = vx * P.z_hat[0] + vy * P.z_hat[1] + vz * P.z_hat[2];
    }
    const double v2 = vx * vx + vy * vy + vz * vz;
    const double ss = -1.0 / PXB_C_KMS;  // scale_shift
    shift = doppler_factor(vn * ss, v2 * (ss * ss));
  }
  D.shift[c] = shift;
  D.gid[c] = L.cell_id ? L.cell_id[c] : P.cell_id0 + c;
  double w = L.dx ? L.dx[c] : 0.0;
  if (P.observer == 0 && P.data_type == 1) w *= 0.5;  // photon_list.py:733-734
  CellGeom g;
  g.cx = x * P.x_hat[0] + y * P.x_hat[1] + z * P.x_hat[2];
  g.cy = x * P.y_hat[0] + y * P.y_hat[1] + z * P.y_hat[2];
  g.cz = x * P.z_hat[0] + y * P.z_hat[1] + z * P.z_hat[2];
  g.w = w;
  if (P.observer == 0 && P.sky_mode != PXB_SKY_PHYS) {
    // angular sky modes: the plane coordinates are wanted in radians (/D_A, photon_list.py:758-760);
    // scaling the centre and the scatter width here saves two multiplications per event
    const double inv_DA = P.D_A_kpc != 0.0 ? 1.0 / P.D_A_kpc : 0.0;
    g.cx *= inv_DA;
    g.cy *= inv_DA;
    g.w *= inv_DA;
  }
  D.geom[c] = g;
  // every warp tile that starts inside this cell's photon range starts in this cell
  const int64_t a = L.poff[c], b = L.poff[c + 1];
  for (int64_t t = (a + PROJ_WTILE_C - 1) / PROJ_WTILE_C; t * PROJ_WTILE_C < b; ++t) D.tile_cell[t] = c;
}

// ---------------------------------------------------------------------------------------
// photon pass
// ---------------------------------------------------------------------------------------
// Each WARP owns a tile of 256 consecutive photons (8 per lane, striped so loads/stores
// coalesce).  A CTA draws one ticket for its 8 warp tiles; after that the warps never meet
// at a CTA barrier again:
//   phase 1  find the photon's cell (32-ary warp search of the offsets, then a shared-memory
//            search), Doppler-shift, decide absorption -> survivors, ballot ranks, tile count
//   phase 2  publish the tile count and resolve the exclusive prefix with a warp-wide
//            decoupled look-back over the per-tile status words (flag | count in 64 bits)
//   phase 3  scatter + sky transform of the survivors, written straight to their final,
//            order-preserving slots.
constexpr int PROJ_THREADS = 256;
constexpr int PROJ_WARPS = PROJ_THREADS / 32;
constexpr int PROJ_ITEMS = 8;
constexpr int PROJ_WTILE = 32 * PROJ_ITEMS;            // photons per warp tile
static_assert(PROJ_WTILE == PROJ_WTILE_C, "tile size used by the cell pre-pass");
constexpr int PROJ_CTILE = PROJ_WARPS * PROJ_WTILE;    // photons per ticket

constexpr unsigned long long ST_AGG = 1ull << 62;   // tile aggregate available
constexpr unsigned long long ST_INC = 2ull << 62;   // inclusive prefix available
constexpr unsigned long long ST_MASK = 3ull << 62;

struct ProjWork {
  unsigned long long* status;  // [n_wtiles]
  unsigned int* ticket;        // [1]
  double* part_sum;            // [n_wtiles]
  double* part_min;
  double* part_max;
};

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v));
}

// exclusive prefix of the counts of all earlier tiles; whole warp participates
__device__ __forceinline__ int64_t warp_lookback(unsigned long long* status, int64_t tile,
                                                 int64_t total, int lane) {
  if (tile == 0) {
    if (lane == 0) st_status(status, ST_INC | (unsigned long long)total);
    return 0;
  }
  if (lane == 0) st_status(status + tile, ST_AGG | (unsigned long long)total);
  int64_t prefix = 0;
  int64_t j = tile - 1;
  for (;;) {
    const int64_t idx = j - lane;
    unsigned long long sv = (idx >= 0) ? ld_status(status + idx) : ST_INC;
    const unsigned inc = __ballot_sync(0xffffffffu, (sv & ST_MASK) == ST_INC);
    const unsigned emp = __ballot_sync(0xffffffffu, (sv & ST_MASK) == 0ull);
    const int first = inc ? (__ffs(inc) - 1) : 32;           // nearest tile with an inclusive prefix
    const unsigned need = first >= 31 ? 0xffffffffu : ((2u << first) - 1u);
    if (emp & need) {                                        // a needed predecessor is not ready
      __nanosleep(1000);
      continue;
    }
    int64_t v = (lane <= first) ? (int64_t)(sv & ~ST_MASK) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    prefix += v;
    if (first < 32) break;
    j -= 32;
  }
  if (lane == 0) st_status(status + tile, ST_INC | (unsigned long long)(prefix + total));
  return prefix;
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
  const double h = sqrt(x * x + den * den);
  // three quotients x^2/(h+den), -x/den, num2/den2 share one reciprocal
  const double hd = h + den;
  const double den2 = hd + x * x * cos_d0;          // (1 + s cos d0) * (h + den)
  const double num2 = y * hd - x * x * sin_d0;      // (y - s sin d0) * (h + den)
  const double inv = 1.0 / (den * den2);
  const double t1 = -x * den2 * inv;
  const double t2 = num2 * den * inv;
  if (den > 0.0 && fabs(t1) <= 0.015625 && fabs(t2) <= 0.015625) {
    // within ~0.9 degrees of the centre the series can stop at t^9 (next term < 8e-20 relative)
    ra_deg = (ra0_rad + atan_tiny(t1)) * PXB_RAD2DEG;
    dec_deg = (d0_rad + atan_tiny(t2)) * PXB_RAD2DEG;
  } else if (den > 0.0 && fabs(t1) <= 0.0625 && fabs(t2) <= 0.0625) {
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

struct SkyConst {
  double sin_d0, cos_d0, ra0_rad, d0_rad, inv_DA;
};

static SkyConst make_sky_const(const pxb_project_params* P) {
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

// scatter + sky transform of one surviving photon (external observer).  SPEC = 0 reads every
// option from the parameter block; SPEC = 1 / 2 are the common case compiled with its options
// fixed (cells, TAN sky, no sigma_pos; on-axis / off-axis normal), which removes the per-photon
// constant loads and uniform branches on those options.
template <int SPEC = 0>
__device__ __forceinline__ void scatter_external(const pxb_project_params& P, const PhiloxKeys& RK,
                                                 const CellGeom& g, const SkyConst& K, uint64_t cell_key,
                                                 uint64_t draw, double& xs_out, double& ys_out,
                                                 double& los_out) {
  const double w = g.w, cx = g.cx, cy = g.cy;
  los_out = g.cz;  // from the unscattered centre (sky_functions.pyx:124,143)
  const Philox4 rb = pxb_rand4(RK, STREAM_PROJ_B, cell_key, draw);
  double ox, oy;  // offsets in the sky plane in units of w
  if (SPEC != 0 || P.data_type == 0) {
    // cells: uniform inside the cube, rotated into the sky plane (sky_functions.pyx:74-76,
    // 110-121; on-axis x_hat/y_hat are axis vectors so only two of the draws contribute)
    const double u0 = u32_centered(rb.x), u1 = u32_centered(rb.y);
    if (SPEC == 1 || (SPEC == 0 && P.vn_axis >= 0)) {   // on-axis: x_hat, y_hat are coordinate axes, two draws suffice
      ox = u0;
      oy = u1;
    } else {
      const double u2 = u32_centered(rb.z);
      ox = u0 * P.x_hat[0] + u1 * P.x_hat[1] + u2 * P.x_hat[2];
      oy = u0 * P.y_hat[0] + u1 * P.y_hat[1] + u2 * P.y_hat[2];
    }
  } else if (P.kernel == 1) {
    normal2(rb, ox, oy);  // sky_functions.pyx:78-80
  } else {
    // top_hat: r ~ U(0,1) (NOT sqrt(U)), theta = 2 pi U  (sky_functions.pyx:81-85)
    const double r = u53(rb.x, rb.y);
    double sn, co;
    sincospi(2.0 * u53(rb.z, rb.w), &sn, &co);
    ox = r * co;
    oy = r * sn;
  }
  double xs = cx + ox * w, ys = cy + oy * w;
  if (SPEC == 0 && P.has_sigma_pos && P.data_type == 0) {
    // photon_list.py:753-756: sigma = sigma_pos * dx
    const Philox4 rc = pxb_rand4(RK, STREAM_PROJ_C, cell_key, draw);
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
                                               const CellGeom& g, uint64_t cell_key, uint64_t draw,
                                               double& lon_out, double& lat_out, double& los_out) {
  const double w = g.w, cx = g.cx, cy = g.cy, cz = g.cz;
  double o0, o1, o2;
  const Philox4 rb = pxb_rand4(RK, STREAM_PROJ_B, cell_key, draw);
  if (P.data_type == 0) {
    o0 = u32_centered(rb.x);
    o1 = u32_centered(rb.y);
    o2 = u32_centered(rb.z);
  } else if (P.kernel == 1) {
    const Philox4 rc = pxb_rand4(RK, STREAM_PROJ_C, cell_key, draw);
    double dump;
    normal2(rb, o0, o1);
    normal2(rc, o2, dump);
  } else {
    const Philox4 rc = pxb_rand4(RK, STREAM_PROJ_C, cell_key, draw);
    const double r = u53(rc.x, rc.y);
    double sp, cp;
    sincospi(2.0 * u53(rb.x, rb.y), &sp, &cp);
    const double ct = 2.0 * u53(rb.z, rb.w) - 1.0;  // cos(theta), theta = arccos(U(-1,1))
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

// absorption decision of one photon (phase 1); kept out of line so the rolled item loop stays
// small
__device__ __forceinline__ bool absorb_one(const pxb_project_params& P, const WabsFast& F, double u,
                                           double e, double nH) {
  if (P.absorb_kind == PXB_ABS_WABS) return wabs_survives(F, u, e, nH);
  return survives(u, table_tau_per_nh(P, e) * nH);
}

// SPEC = 1: the common case compiled with its options fixed (wabs, a foreground column only)
template <int SPEC = 0>
__device__ __forceinline__ bool absorb_decision(const pxb_project_params& P, const PhiloxKeys& RK,
                                                const WabsFast& F, int64_t c, uint64_t cell_key,
                                                int64_t draw, double en) {
  if (SPEC == 1) {
    const Philox4 ra = pxb_rand4(RK, STREAM_PROJ_A, cell_key, (uint64_t)draw);
    return wabs_survives(F, u53(ra.x, ra.y), en, P.nH);
  }
  bool det = true;
  if (P.nH_cell) {
    // internal absorption on the source-frame energy, photon_list.py:715-723
    const Philox4 rd = pxb_rand4(RK, STREAM_PROJ_D, cell_key, (uint64_t)draw);
    det = absorb_one(P, F, u53(rd.x, rd.y), en * P.oneplusz, __ldg(P.nH_cell + c));
  }
  if (det && P.nH > 0.0) {
    const Philox4 ra = pxb_rand4(RK, STREAM_PROJ_A, cell_key, (uint64_t)draw);
    det = absorb_one(P, F, u53(ra.x, ra.y), en, P.nH);
  }
  return det;
}

template <int OBS>
__global__ void __launch_bounds__(PROJ_THREADS, 4)
k_project_photons(const __grid_constant__ pxb_photon_list L,
                  const __grid_constant__ pxb_project_params P, const __grid_constant__ PhiloxKeys RK, const CellDerived D,
                  const ProjWork W, double* __restrict__ xsky, double* __restrict__ ysky,
                  double* __restrict__ eobs_out, double* __restrict__ los_out,
                  int64_t* __restrict__ n_events_out, int64_t n_wtiles, int64_t n_ctiles,
                  const SkyConst K) {
  // per-warp scratch: cell offsets of the tile and the per-photon state carried from phase 1
  // to phase 3 (energy, cell, output slot) -- kept in shared memory so the item loops stay
  // rolled (small code, few registers, more resident warps)
  __shared__ int64_t s_off_all[PROJ_WARPS][PROJ_WTILE + 2];
  __shared__ double s_e_all[PROJ_WARPS][PROJ_WTILE];
  __shared__ int s_rel_all[PROJ_WARPS][PROJ_WTILE];
  __shared__ short s_slot_all[PROJ_WARPS][PROJ_WTILE];
  __shared__ unsigned int s_ticket;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int64_t* s_off = s_off_all[wid];
  double* s_e = s_e_all[wid];
  int* s_rel = s_rel_all[wid];
  short* s_slot = s_slot_all[wid];
  const bool absorb = P.absorb_kind != PXB_ABS_NONE;
  __shared__ WabsFast s_wabs;
  if (P.absorb_kind == PXB_ABS_WABS) wabs_fast_init(s_wabs);   // published by the loop's barrier

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(W.ticket, 1u);
    __syncthreads();
    const int64_t ctile = s_ticket;
    if (ctile >= n_ctiles) return;
    // the scheduler favours high warp ids, so those take the EARLIER tiles of the ticket: the
    // warps that finish phase 1 last then own the tiles whose predecessors are already done
    const int64_t wtile = ctile * PROJ_WARPS + (PROJ_WARPS - 1 - wid);
    if (wtile >= n_wtiles) continue;   // past the end: nobody waits on this warp
    const int64_t p0 = wtile * PROJ_WTILE;
    const int64_t p1 = min(p0 + (int64_t)PROJ_WTILE, L.n_photons);

    // ---- cells intersecting this warp tile -------------------------------------------
    // (the per-cell pre-pass recorded the cell every tile starts in; the tile ends in the cell
    // the next tile starts in, or one before it)
    const int64_t c0 = __ldg(D.tile_cell + wtile);
    const int64_t c1 = (wtile + 1 < n_wtiles) ? __ldg(D.tile_cell + wtile + 1) : L.n_cells - 1;
    const int nc = (c1 - c0 + 1 <= PROJ_WTILE + 1) ? (int)(c1 - c0 + 1) : 0;
    const int step0 = nc > 1 ? (1 << (31 - __clz(nc - 1))) : 0;
    __syncwarp();
    for (int q = lane; q <= nc && nc > 0; q += 32) s_off[q] = __ldg(L.poff + c0 + q);
    __syncwarp();

    // all eight energy loads of the lane go out together (the item loop below is rolled)
#pragma unroll
    for (int r = 0; r < PROJ_ITEMS; ++r) {
      const int it = r * 32 + lane;
      if (p0 + it < p1) s_e[it] = ld_stream(L.energy + p0 + it);
    }

    // ---- phase 1: Doppler + absorption, survivors ----------------------------------------
    int run = 0;
    double tsum = 0.0, tmin = 1.0e99, tmax = -1.0e99;
#pragma unroll 1
    for (int r = 0; r < PROJ_ITEMS; ++r) {
      const int it = r * 32 + lane;
      const int64_t p = p0 + it;
      bool det = false;
      if (p < p1) {
        int rel = 0;
        int64_t first;
        if (nc == 1) {
          first = s_off[0];
        } else if (nc > 0) {
          // branch-free bisection: largest q in [0, nc) with s_off[q] <= p
          for (int step = step0; step > 0; step >>= 1) {   // step0: largest power of two < nc
            const int q = rel + step;
            if (q < nc && s_off[q] <= p) rel = q;
          }
          first = s_off[rel];
        } else {  // more cells than photons in the tile (many empty cells): global search
          int64_t lo = c0, hi = c1 + 1;
          while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(L.poff + mid) <= p) lo = mid; else hi = mid;
          }
          rel = (int)(lo - c0);
          first = __ldg(L.poff + lo);
        }
        const int64_t c = c0 + rel;
        const double en = s_e[it] * __ldg(D.shift + c);
        det = absorb ? absorb_decision(P, RK, s_wabs, c, (uint64_t)__ldg(D.gid + c), p - first, en) : true;
        s_e[it] = en;
        s_rel[it] = rel;
        if (det) {
          tsum += en;
          tmin = fmin(tmin, en);
          tmax = fmax(tmax, en);
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, det);
      s_slot[it] = det ? (short)(run + __popc(bal & ((1u << lane) - 1u))) : (short)-1;
      run += __popc(bal);
    }
    tsum = warp_sum(tsum);
    tmin = warp_min(tmin);
    tmax = warp_max(tmax);
    if (lane == 0) {
      W.part_sum[wtile] = tsum;
      W.part_min[wtile] = tmin;
      W.part_max[wtile] = tmax;
    }

    // ---- phase 2: exclusive prefix over tiles ----------------------------------------------
    const int64_t base = warp_lookback(W.status, wtile, (int64_t)run, lane);
    if (wtile == n_wtiles - 1 && lane == 0) *n_events_out = base + run;

    // ---- phase 3: scatter survivors into their final slots ---------------------------------
#pragma unroll 1
    for (int r = 0; r < PROJ_ITEMS; ++r) {
      const int it = r * 32 + lane;
      const int sl = s_slot[it];       // own entry: written by this lane in phase 1
      if (sl < 0 || p0 + it >= p1) continue;
      const int rel = s_rel[it];
      const int64_t c = c0 + rel;
      const int64_t first = (nc > 0) ? s_off[rel] : __ldg(L.poff + c);
      const uint64_t cell_key = (uint64_t)__ldg(D.gid + c);
      const uint64_t draw = (uint64_t)(p0 + it - first);
      double xs, ys, los;
      const CellGeom g = ld_geom(D.geom + c);
      if (OBS == 0) scatter_external(P, RK, g, K, cell_key, draw, xs, ys, los);
      else scatter_allsky(P, RK, g, cell_key, draw, xs, ys, los);
      const int64_t o = base + sl;
      st_stream(xsky + o, xs);
      st_stream(ysky + o, ys);
      st_stream(eobs_out + o, s_e[it]);
      if (los_out) st_stream(los_out + o, los);
    }
  }
}

// ---------------------------------------------------------------------------------------
// two-pass variant (default): detect (survivor records + per-tile counts) -> scan -> emit
// ---------------------------------------------------------------------------------------
// Same arithmetic and the same event list as the fused kernel, organised without any
// inter-tile dependency: pass 1 decides absorption for every photon and leaves, for every
// SURVIVOR in photon order, a 16-bit record (position inside the tile | cell relative to the
// tile's first) plus the tile's count; an exclusive scan of the counts gives every tile its
// base; pass 2 walks the dense records, so all 32 lanes of a round scatter a survivor and the
// event stores are contiguous.  Costs one more read of the energies (+8 B/photon) and the
// records (+4 B/event) and buys: no tickets, no look-back spinning, no idle lanes for absorbed
// photons.
template <int SPEC>
__global__ void __launch_bounds__(PROJ_THREADS, SPEC == 1 ? 5 : 4)
k_project_detect(const __grid_constant__ pxb_photon_list L,
                 const __grid_constant__ pxb_project_params P, const __grid_constant__ PhiloxKeys RK, const CellDerived D,
                 const ProjWork W, uint16_t* __restrict__ slots, int64_t* __restrict__ counts,
                 int64_t n_wtiles) {
  __shared__ int64_t s_off_all[PROJ_WARPS][PROJ_WTILE + 2];
  __shared__ WabsFast s_wabs;
  if (P.absorb_kind == PXB_ABS_WABS) {
    wabs_fast_init(s_wabs);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int64_t* s_off = s_off_all[wid];
  const int64_t wtile = (int64_t)blockIdx.x * PROJ_WARPS + wid;
  if (wtile >= n_wtiles) return;
  const int64_t p0 = wtile * PROJ_WTILE;
  const int64_t p1 = min(p0 + (int64_t)PROJ_WTILE, L.n_photons);
  // all eight energy loads of the lane go out first, together with the tile's cell range; the
  // per-cell loads that depend on it follow while the energies are still in flight
  __shared__ double s_e_all[PROJ_WARPS][PROJ_WTILE];
  __shared__ double s_shift_all[PROJ_WARPS][PROJ_STAGE_CELLS];
  double* s_e = s_e_all[wid];
  double* s_shift = s_shift_all[wid];
  double ev[PROJ_ITEMS];
#pragma unroll
  for (int r = 0; r < PROJ_ITEMS; ++r) {
    const int64_t p = p0 + r * 32 + lane;
    ev[r] = (p < p1) ? ld_stream(L.energy + p) : 0.0;
  }
  const int64_t c0 = __ldg(D.tile_cell + wtile);
  const int64_t c1 = (wtile + 1 < n_wtiles) ? __ldg(D.tile_cell + wtile + 1) : L.n_cells - 1;
  const int nc = (c1 - c0 + 1 <= PROJ_WTILE + 1) ? (int)(c1 - c0 + 1) : 0;
  const int step0 = nc > 1 ? (1 << (31 - __clz(nc - 1))) : 0;
  const bool staged = nc >= 1 && nc <= PROJ_STAGE_CELLS;
  for (int q = lane; q <= nc && nc > 0; q += 32) s_off[q] = __ldg(L.poff + c0 + q);
  if (staged && lane < nc) s_shift[lane] = __ldg(D.shift + c0 + lane);
#pragma unroll
  for (int r = 0; r < PROJ_ITEMS; ++r) s_e[r * 32 + lane] = ev[r];
  __syncwarp();
  const bool absorb = P.absorb_kind != PXB_ABS_NONE;
  int run = 0;
  double tsum = 0.0, tmin = 1.0e99, tmax = -1.0e99;
#pragma unroll 1
  for (int r = 0; r < PROJ_ITEMS; ++r) {
    const int64_t p = p0 + r * 32 + lane;
    bool det = false;
    int rel = 0;
    if (p < p1) {
      int64_t first;
      if (nc == 1) {
        first = s_off[0];
      } else if (nc > 0) {
        for (int step = step0; step > 0; step >>= 1) {   // step0: largest power of two < nc
          const int q = rel + step;
          if (q < nc && s_off[q] <= p) rel = q;
        }
        first = s_off[rel];
      } else {
        int64_t lo = c0, hi = c1 + 1;
        while (hi - lo > 1) {
          const int64_t mid = (lo + hi) >> 1;
          if (__ldg(L.poff + mid) <= p) lo = mid; else hi = mid;
        }
        rel = (int)(lo - c0);
        first = __ldg(L.poff + lo);
      }
      const int64_t c = c0 + rel;
      const double en = s_e[r * 32 + lane] * (staged ? s_shift[rel] : __ldg(D.shift + c));
      const uint64_t cell_key = (uint64_t)__ldg(D.gid + c);
      det = SPEC == 1 ? absorb_decision<1>(P, RK, s_wabs, c, cell_key, p - first, en)
                      : (absorb ? absorb_decision(P, RK, s_wabs, c, cell_key, p - first, en) : true);
      if (det) {
        tsum += en;
        tmin = en < tmin ? en : tmin;   // (energies are never NaN: plain compares, no fmin/fmax)
        tmax = en > tmax ? en : tmax;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, det);
    // survivor record: position inside the tile (8 bits) | cell relative to the tile's first (8)
    if (det)
      slots[wtile * PROJ_WTILE + run + __popc(bal & ((1u << lane) - 1u))] =
          (uint16_t)((r * 32 + lane) | ((rel & 255) << 8));
    run += __popc(bal);
  }
  tsum = warp_sum(tsum);
  tmin = warp_min(tmin);
  tmax = warp_max(tmax);
  if (lane == 0) {
    counts[wtile] = run;
    W.part_sum[wtile] = tsum;
    W.part_min[wtile] = tmin;
    W.part_max[wtile] = tmax;
  }
}

template <int OBS, int SPEC>
__global__ void __launch_bounds__(PROJ_THREADS, 4)
k_project_emit(const __grid_constant__ pxb_photon_list L,
               const __grid_constant__ pxb_project_params P, const __grid_constant__ PhiloxKeys RK, const CellDerived D,
               const uint16_t* __restrict__ slots, const int64_t* __restrict__ offsets,
               double* __restrict__ xsky, double* __restrict__ ysky,
               double* __restrict__ eobs_out, double* __restrict__ los_out,
               int64_t* __restrict__ n_events_out, int64_t n_wtiles, const SkyConst K) {
  __shared__ int64_t s_off_all[PROJ_WARPS][PROJ_WTILE + 2];
  __shared__ double s_e_all[PROJ_WARPS][PROJ_WTILE];
  __shared__ uint16_t s_rec_all[PROJ_WARPS][PROJ_WTILE];
  __shared__ CellGeom s_geom_all[PROJ_WARPS][PROJ_STAGE_CELLS];
  __shared__ double s_shift_all[PROJ_WARPS][PROJ_STAGE_CELLS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int64_t* s_off = s_off_all[wid];
  double* s_e = s_e_all[wid];
  uint16_t* s_rec = s_rec_all[wid];
  CellGeom* s_geom = s_geom_all[wid];
  double* s_shift = s_shift_all[wid];
  const int64_t wtile = (int64_t)blockIdx.x * PROJ_WARPS + wid;
  if (wtile >= n_wtiles) return;
  if (wtile == 0 && lane == 0) *n_events_out = __ldg(offsets + n_wtiles);
  const int64_t p0 = wtile * PROJ_WTILE;
  const int64_t p1 = min(p0 + (int64_t)PROJ_WTILE, L.n_photons);
  // ---- phase A: every load that does not depend on another goes out at once --------------
  // (tile's energies and survivor records, its base/count, the cells it starts and ends in)
  double ev[PROJ_ITEMS];
  unsigned rv[PROJ_ITEMS];
#pragma unroll
  for (int r = 0; r < PROJ_ITEMS; ++r) {
    const int64_t p = p0 + r * 32 + lane;
    ev[r] = (p < p1) ? ld_stream(L.energy + p) : 0.0;
    rv[r] = __ldg(slots + p);   // slot r*32+lane of this tile (garbage past the count: unused)
  }
  const int64_t base = __ldg(offsets + wtile);
  const int count = (int)(__ldg(offsets + wtile + 1) - base);   // survivors of this tile
  const int64_t c0 = __ldg(D.tile_cell + wtile);
  const int64_t c1 = (wtile + 1 < n_wtiles) ? __ldg(D.tile_cell + wtile + 1) : L.n_cells - 1;
  if (count == 0) return;                                       // everything absorbed
  // ---- phase B: per-cell data of the tile (depends on c0) ---------------------------------
  const int nc = (c1 - c0 + 1 <= PROJ_WTILE + 1) ? (int)(c1 - c0 + 1) : 0;
  const bool packed_rel = nc >= 1 && nc <= 256;   // the record's 8 bits hold the cell
  const bool staged = nc >= 1 && nc <= PROJ_STAGE_CELLS;
  for (int q = lane; q <= nc && nc > 0; q += 32) s_off[q] = __ldg(L.poff + c0 + q);
  if (staged && lane < nc) {
    s_geom[lane] = ld_geom(D.geom + c0 + lane);
    s_shift[lane] = __ldg(D.shift + c0 + lane);
  }
#pragma unroll
  for (int r = 0; r < PROJ_ITEMS; ++r) {
    s_e[r * 32 + lane] = ev[r];
    s_rec[r * 32 + lane] = (uint16_t)rv[r];
  }
  __syncwarp();
  // survivors were numbered by the detect pass: slot s of the tile is photon p0 + (record & 255),
  // so every round works on 32 live lanes and the event stores are dense
#pragma unroll 1
  for (int sl = lane; sl < count; sl += 32) {
    const unsigned rec = s_rec[sl];
    const int it = rec & 255u;
    const int64_t p = p0 + it;
    int rel;
    int64_t first;
    if (packed_rel) {
      rel = (int)(rec >> 8);
      first = s_off[rel];
    } else if (nc > 0) {
      rel = 0;
      for (int step = 256; step > 0; step >>= 1) {
        const int q = rel + step;
        if (q < nc && s_off[q] <= p) rel = q;
      }
      first = s_off[rel];
    } else {
      int64_t lo = c0, hi = c1 + 1;
      while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(L.poff + mid) <= p) lo = mid; else hi = mid;
      }
      rel = (int)(lo - c0);
      first = __ldg(L.poff + lo);
    }
    const int64_t c = c0 + rel;
    const CellGeom g = staged ? s_geom[rel] : ld_geom(D.geom + c);
    const double en = s_e[it] * (staged ? s_shift[rel] : __ldg(D.shift + c));
    const uint64_t cell_key = (uint64_t)__ldg(D.gid + c);
    const uint64_t draw = (uint64_t)(p - first);
    double xs, ys, los;
    if (OBS == 0) scatter_external<SPEC>(P, RK, g, K, cell_key, draw, xs, ys, los);
    else scatter_allsky(P, RK, g, cell_key, draw, xs, ys, los);
    const int64_t o = base + sl;
    st_stream(xsky + o, xs);
    st_stream(ysky + o, ys);
    st_stream(eobs_out + o, en);
    if (los_out) st_stream(los_out + o, los);
  }
}

// deterministic two-level reduction of the per-tile statistics
__global__ void __launch_bounds__(256)
k_reduce_stats(const double* __restrict__ ps, const double* __restrict__ pmin,
               const double* __restrict__ pmax, int64_t n, double* __restrict__ out,
               int64_t per_block) {
  __shared__ double s[3][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t b0 = (int64_t)blockIdx.x * per_block;
  const int64_t b1 = min(b0 + per_block, n);
  double a = 0.0, mn = 1.0e99, mx = -1.0e99;
  for (int64_t i = b0 + threadIdx.x; i < b1; i += blockDim.x) {
    a += ps[i]; mn = fmin(mn, pmin[i]); mx = fmax(mx, pmax[i]);
  }
  a = warp_sum(a); mn = warp_min(mn); mx = warp_max(mx);
  if (lane == 0) { s[0][wid] = a; s[1][wid] = mn; s[2][wid] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { a += s[0][w]; mn = fmin(mn, s[1][w]); mx = fmax(mx, s[2][w]); }
    out[3 * blockIdx.x + 0] = a; out[3 * blockIdx.x + 1] = mn; out[3 * blockIdx.x + 2] = mx;
  }
}
__global__ void k_reduce_stats_final(const double* __restrict__ part, int nb,
                                     double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a = 0.0, mn = 1.0e99, mx = -1.0e99;
  for (int i = 0; i < nb; ++i) {
    a += part[3 * i]; mn = fmin(mn, part[3 * i + 1]); mx = fmax(mx, part[3 * i + 2]);
  }
  out[0] = a; out[1] = mn; out[2] = mx;
}

// ---------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------
constexpr int STATS_BLOCKS = 256;

static inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

struct ProjLayout {
  int64_t ntiles;
  int64_t off_cells, off_tilecell, off_status, off_ticket, off_part, off_red, total;
  int64_t off_masks, off_counts, off_offsets, off_arank, off_totals, off_scanws, scanws_bytes;
};

static ProjLayout proj_layout(int64_t n_cells, int64_t n_photons) {
  ProjLayout l;
  l.ntiles = (n_photons + PROJ_WTILE - 1) / PROJ_WTILE;   // warp tiles
  int64_t o = 0;
  l.off_cells = o;  o += align256(6 * n_cells * (int64_t)sizeof(double));
  l.off_tilecell = o; o += align256((l.ntiles + 1) * 8);
  l.off_status = o; o += align256(l.ntiles * 8);
  l.off_ticket = o; o += 256;
  l.off_part = o;   o += align256(3 * l.ntiles * 8);
  l.off_red = o;    o += align256(3 * STATS_BLOCKS * 8);
  // two-pass variant
  l.off_masks = o;   o += align256(l.ntiles * PROJ_WTILE * 2);
  l.off_counts = o;  o += align256(l.ntiles * 8);
  l.off_offsets = o; o += align256((l.ntiles + 1) * 8);
  l.off_arank = o;   o += align256((l.ntiles + 1) * 8);
  l.off_totals = o;  o += 256;
  l.scanws_bytes = 0;
  pxb_scan_workspace_bytes(l.ntiles, &l.scanws_bytes);
  l.off_scanws = o;  o += align256(l.scanws_bytes);
  l.total = o;
  return l;
}

extern "C" int pxb_project_workspace_bytes(int64_t n_cells, int64_t n_photons, int64_t* bytes) {
  PXB_REQUIRE(bytes && n_photons >= 0 && n_cells >= 0, "bad argument");
  *bytes = proj_layout(n_cells, n_photons).total;
  return PXB_OK;
}

static int check_params(const pxb_project_params* P) {
  PXB_REQUIRE(P, "NULL params");
  PXB_REQUIRE(P->absorb_kind >= 0 && P->absorb_kind <= 2, "bad absorb_kind");
  if (P->absorb_kind == PXB_ABS_TABLE)
    PXB_REQUIRE(P->abs_energy && P->abs_sigma && P->abs_n >= 1, "absorption table missing");
  return PXB_OK;
}

extern "C" int pxb_project(const pxb_photon_list* L, const pxb_project_params* P, double* xsky,
                           double* ysky, double* eobs, double* los, int64_t* n_events_out,
                           double* stats_out, void* workspace, int64_t workspace_bytes,
                           void* stream) {
  PXB_REQUIRE(L && n_events_out && stats_out, "NULL argument");
  int rc = check_params(P);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (L->n_photons == 0 || L->n_cells == 0) {
    PXB_CUDA(cudaMemsetAsync(n_events_out, 0, sizeof(int64_t), st));
    const double init[3] = {0.0, 1.0e99, -1.0e99};
    PXB_CUDA(cudaMemcpyAsync(stats_out, init, sizeof(init), cudaMemcpyHostToDevice, st));
    PXB_CUDA(cudaStreamSynchronize(st));
    return PXB_OK;
  }
  PXB_REQUIRE(xsky && ysky && eobs, "output arrays missing");
  PXB_REQUIRE(!P->save_los || los, "save_los needs the los array");
  PXB_REQUIRE(L->nph && L->poff && L->x && L->y && L->z && L->energy, "photon list incomplete");
  PXB_REQUIRE(P->no_shifting || (L->vx && L->vy && L->vz), "velocities missing");
  PXB_REQUIRE(P->observer == 1 || P->sky_mode == PXB_SKY_PHYS || P->D_A_kpc > 0.0,
              "D_A must be positive for sky coordinates");
  const ProjLayout lay = proj_layout(L->n_cells, L->n_photons);
  PXB_REQUIRE(workspace && workspace_bytes >= lay.total,
              "projection workspace too small (%lld < %lld)", (long long)workspace_bytes,
              (long long)lay.total);
  const int64_t n_ctiles = (lay.ntiles + PROJ_WARPS - 1) / PROJ_WARPS;
  PXB_REQUIRE(n_ctiles < (1LL << 32) - 65536, "too many photons for one projection call");
  char* ws = (char*)workspace;
  CellDerived D;
  D.geom = (CellGeom*)(ws + lay.off_cells);                 // 32-byte records first (alignment)
  D.shift = (double*)(ws + lay.off_cells + 4 * L->n_cells * (int64_t)sizeof(double));
  D.gid = (int64_t*)(ws + lay.off_cells + 5 * L->n_cells * (int64_t)sizeof(double));
  D.tile_cell = (int64_t*)(ws + lay.off_tilecell);
  ProjWork W;
  W.status = (unsigned long long*)(ws + lay.off_status);
  W.ticket = (unsigned int*)(ws + lay.off_ticket);
  W.part_sum = (double*)(ws + lay.off_part);
  W.part_min = W.part_sum + lay.ntiles;
  W.part_max = W.part_min + lay.ntiles;
  double* red = (double*)(ws + lay.off_red);
  // status + ticket are adjacent: one memset
  PXB_CUDA(cudaMemsetAsync(ws + lay.off_status, 0, (size_t)(lay.off_part - lay.off_status), st));
  k_project_cells<<<(unsigned)((L->n_cells + 255) / 256), 256, 0, st>>>(*L, *P, D);
  PXB_LAUNCH_CHECK();
  // default: two passes (detect -> scan -> emit): no inter-tile waiting, small rolled kernels,
  // at the price of reading the energies twice (+8 B/photon).  params.single_pass selects
  // the single-pass kernel (DRAM traffic == algorithmic bytes, decoupled look-back), which
  // produces the identical event list; on B200 it is ~10 % slower while the path is ALU-bound.
  const SkyConst K = make_sky_const(P);
  const PhiloxKeys RK = pxb_make_keys(P->seed);
  const bool fused = P->single_pass != 0;
  if (!fused) {
    uint16_t* masks = (uint16_t*)(ws + lay.off_masks);   // survivor records
    int64_t* counts = (int64_t*)(ws + lay.off_counts);
    int64_t* offsets = (int64_t*)(ws + lay.off_offsets);
    int64_t* arank = (int64_t*)(ws + lay.off_arank);
    int64_t* totals = (int64_t*)(ws + lay.off_totals);
    const unsigned g = (unsigned)n_ctiles;
    const bool generic = P->generic_kernels != 0;   // diagnostic: no specialisation
    if (!generic && P->absorb_kind == PXB_ABS_WABS && !P->nH_cell && P->nH > 0.0)
      k_project_detect<1><<<g, PROJ_THREADS, 0, st>>>(*L, *P, RK, D, W, masks, counts, lay.ntiles);
    else
      k_project_detect<0><<<g, PROJ_THREADS, 0, st>>>(*L, *P, RK, D, W, masks, counts, lay.ntiles);
    PXB_LAUNCH_CHECK();
    rc = pxb_scan_counts(counts, lay.ntiles, offsets, arank, totals, ws + lay.off_scanws,
                         lay.scanws_bytes, stream);
    if (rc) return rc;
    // the common external case (cells, TAN sky, no sigma_pos) runs a kernel compiled for it
    const bool common = !generic && P->observer == 0 && P->data_type == 0 && !P->has_sigma_pos &&
                        P->sky_mode == PXB_SKY_TAN;
    double* losp = P->save_los ? los : nullptr;
#define PXB_EMIT(OBS_, SPEC_)                                                                   \
  k_project_emit<OBS_, SPEC_><<<g, PROJ_THREADS, 0, st>>>(*L, *P, RK, D, masks, offsets, xsky, ysky, \
                                                         eobs, losp, n_events_out, lay.ntiles, K)
    if (P->observer != 0) PXB_EMIT(1, 0);
    else if (common && P->vn_axis >= 0) PXB_EMIT(0, 1);
    else if (common) PXB_EMIT(0, 2);
    else PXB_EMIT(0, 0);
#undef PXB_EMIT
    PXB_LAUNCH_CHECK();
  } else {
  int per_sm = 0;
  if (P->observer == 0) {
    PXB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_project_photons<0>, PROJ_THREADS, 0));
  } else {
    PXB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_project_photons<1>, PROJ_THREADS, 0));
  }
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)pxb_sm_count() * per_sm;
  if (grid > n_ctiles) grid = n_ctiles;
  if (P->observer == 0) {
    k_project_photons<0><<<(unsigned)grid, PROJ_THREADS, 0, st>>>(
        *L, *P, RK, D, W, xsky, ysky, eobs, P->save_los ? los : nullptr, n_events_out, lay.ntiles, n_ctiles, K);
  } else {
    k_project_photons<1><<<(unsigned)grid, PROJ_THREADS, 0, st>>>(
        *L, *P, RK, D, W, xsky, ysky, eobs, P->save_los ? los : nullptr, n_events_out, lay.ntiles, n_ctiles, K);
  }
  PXB_LAUNCH_CHECK();
  }
  const int64_t per_block = (lay.ntiles + STATS_BLOCKS - 1) / STATS_BLOCKS;
  const int nb = (int)((lay.ntiles + per_block - 1) / per_block);
  k_reduce_stats<<<nb, 256, 0, st>>>(W.part_sum, W.part_min, W.part_max, lay.ntiles, red, per_block);
  PXB_LAUNCH_CHECK();
  k_reduce_stats_final<<<1, 32, 0, st>>>(red, nb, stats_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_doppler_shift(int64_t n_cells, const double* beta_n, const double* beta2,
                                 const int64_t* poff, int64_t n_photons, double* eobs,
                                 void* stream) {
  PXB_REQUIRE(n_cells >= 0 && n_photons >= 0, "bad sizes");
  if (n_cells == 0 || n_photons == 0) return PXB_OK;
  PXB_REQUIRE(beta_n && beta2 && poff && eobs, "NULL argument");
  const int64_t ntiles = (n_photons + 1023) / 1024;
  PXB_REQUIRE(ntiles < (1LL << 31), "too many photons");
  k_doppler_shift<<<(unsigned)ntiles, 256, 0, (cudaStream_t)stream>>>(n_cells, beta_n, beta2, poff,
                                                                     n_photons, eobs);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_pixel_to_cel(int64_t n, double* xsky, double* ysky, double ra0_deg,
                                double dec0_deg, void* stream) {
  PXB_REQUIRE(n >= 0, "bad n");
  if (n == 0) return PXB_OK;
  PXB_REQUIRE(xsky && ysky, "NULL argument");
  const double d0 = dec0_deg * PXB_PI / 180.0;
  k_pixel_to_cel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      n, xsky, ysky, ra0_deg * PXB_PI / 180.0, sin(d0), cos(d0));
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

// ---------------------------------------------------------------------------------------
// internal absorption: column-density cube lookup (photon_list.py:672-680)
// ---------------------------------------------------------------------------------------
// scipy RegularGridInterpolator, method="linear": per axis i = clip(searchsorted(g, x) - 1, 0,
// n-2), t = (x - g[i]) / (g[i+1] - g[i]); outside [g[0], g[n-1]] on any axis -> fill value 0.
__device__ __forceinline__ bool grid_locate(const double* __restrict__ g, int64_t n, double x,
                                            int64_t& i, double& t) {
  if (!(x >= __ldg(g)) || !(x <= __ldg(g + n - 1))) return false;   // NaN is outside too
  int64_t lo = 0, hi = n - 1;   // g[lo] <= x <= g[hi]
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(g + mid) <= x) lo = mid; else hi = mid;
  }
  // searchsorted(side="left") - 1 puts x == g[k] (k > 0) in interval k-1 with t = 1; the
  // interpolant is continuous there, the choice only matters for bit patterns
  if (lo > 0 && __ldg(g + lo) == x) --lo;
  i = lo;
  const double g0 = __ldg(g + lo), g1 = __ldg(g + lo + 1);
  t = (x - g0) / (g1 - g0);
  return true;
}

__global__ void __launch_bounds__(256)
k_column_lookup(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                const double* __restrict__ z, const double* __restrict__ uv, int64_t nw,
                int64_t nd, const double* __restrict__ wmid, const double* __restrict__ dmid,
                const double* __restrict__ grid, double* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const double px = x[c], py = y[c], pz = z[c];
  const double a = __ldg(uv + 0) * px + __ldg(uv + 1) * py + __ldg(uv + 2) * pz;
  const double b = __ldg(uv + 3) * px + __ldg(uv + 4) * py + __ldg(uv + 5) * pz;
  const double d = __ldg(uv + 6) * px + __ldg(uv + 7) * py + __ldg(uv + 8) * pz;
  int64_t ia, ib, id;
  double ta, tb, td;
  double v = 0.0;
  if (grid_locate(wmid, nw, a, ia, ta) && grid_locate(wmid, nw, b, ib, tb) &&
      grid_locate(dmid, nd, d, id, td)) {
#pragma unroll
    for (int ca = 0; ca < 2; ++ca)
#pragma unroll
      for (int cb = 0; cb < 2; ++cb)
#pragma unroll
        for (int cd = 0; cd < 2; ++cd) {
          // same association and no fused multiply-add: bit-identical to scipy's hypercube sum
          const double w = __dmul_rn(__dmul_rn(ca ? ta : 1.0 - ta, cb ? tb : 1.0 - tb),
                                     cd ? td : 1.0 - td);
          v = __dadd_rn(v, __dmul_rn(__ldg(grid + ((ia + ca) * nw + (ib + cb)) * nd + (id + cd)), w));
        }
  }
  out[c] = v;
}

extern "C" int pxb_column_lookup(int64_t n_cells, const double* x, const double* y, const double* z,
                                 const double* unit_vectors, int64_t nw, int64_t nd,
                                 const double* wmid, const double* dmid, const double* nH_grid,
                                 double* nH_cell_out, void* stream) {
  PXB_REQUIRE(n_cells >= 0, "bad n_cells");
  PXB_REQUIRE(nw >= 2 && nd >= 2, "the column cube needs at least two nodes per axis");
  if (n_cells == 0) return PXB_OK;
  PXB_REQUIRE(x && y && z && unit_vectors && wmid && dmid && nH_grid && nH_cell_out, "NULL argument");
  k_column_lookup<<<(unsigned)((n_cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      n_cells, x, y, z, unit_vectors, nw, nd, wmid, dmid, nH_grid, nH_cell_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

__global__ void __launch_bounds__(256)
k_absorb_decide(const __grid_constant__ pxb_project_params P, int64_t n,
                const double* __restrict__ e, const double* __restrict__ u,
                uint8_t* __restrict__ out) {
  __shared__ WabsFast s_wabs;
  if (P.absorb_kind == PXB_ABS_WABS) {
    wabs_fast_init(s_wabs);
    __syncthreads();
  }
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (P.absorb_kind == PXB_ABS_NONE) ? 1 : (absorb_one(P, s_wabs, u[i], e[i], P.nH) ? 1 : 0);
}

extern "C" int pxb_absorb_decide(const pxb_project_params* params, int64_t n, const double* energy,
                                 const double* u, uint8_t* survive_out, void* stream) {
  int rc = check_params(params);
  if (rc) return rc;
  PXB_REQUIRE(n >= 0, "bad n");
  if (n == 0) return PXB_OK;
  PXB_REQUIRE(energy && u && survive_out, "NULL argument");
  k_absorb_decide<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      *params, n, energy, u, survive_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_absorb_transmission(const pxb_project_params* params, int64_t n,
                                       const double* energy, double* trans_out, void* stream) {
  int rc = check_params(params);
  if (rc) return rc;
  PXB_REQUIRE(n >= 0, "bad n");
  if (n == 0) return PXB_OK;
  PXB_REQUIRE(energy && trans_out, "NULL argument");
  k_absorb_transmission<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      *params, n, energy, trans_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}
