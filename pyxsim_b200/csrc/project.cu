// project.cu -- stand-alone natives of the projection step (deterministic parity against the
// reference Cython / NumPy): in-place Doppler shift, inverse gnomonic projection, absorption
// transmission and decisions, column-density cube lookup.  The projection itself is the photon
// pipeline in pipeline.cu; the per-photon device code both share is projection.cuh.
//
// Reference:
//   doppler_shift            pyxsim/lib/sky_functions.pyx:20-34
//   pixel_to_cel             sky_functions.pyx:153-179
//   absorb_photons           pyxsim/spectral_models.py:556-583
//   internal absorption      pyxsim/photon_list.py:672-680
#include "projection.cuh"
#include "tiles.cuh"

__global__ void __launch_bounds__(256)
k_absorb_transmission(const __grid_constant__ pxb_project_params P, int64_t n,
                      const double* __restrict__ e, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (P.absorb_kind == PXB_ABS_NONE) ? 1.0 : exp(-tau_per_nh(P, e[i]) * P.nH);
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
  int rc = pxb_check_project_params(params);
  if (rc) return rc;
  PXB_REQUIRE(n >= 0, "bad n");
  if (n == 0) return PXB_OK;
  PXB_REQUIRE(energy && u && survive_out, "NULL argument");
  k_absorb_decide<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      *params, n, energy, u, survive_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

// the pipeline's own decision from random WORDS (projection.cuh): photon i of cell `cell_key` has
// draw index draw0 + i, top word w[i]; the low word comes from its refinement stream when needed
__global__ void __launch_bounds__(256)
k_absorb_decide_words(const __grid_constant__ pxb_project_params P, const __grid_constant__ PhiloxKeys RK,
                      int64_t n, const double* __restrict__ e, const uint32_t* __restrict__ w,
                      uint64_t cell_key, uint64_t draw0, uint8_t* __restrict__ out) {
  __shared__ WabsFast s_wabs;
  if (P.absorb_kind == PXB_ABS_WABS) {
    wabs_fast_init(s_wabs);
    __syncthreads();
  }
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool det = true;
  if (P.absorb_kind == PXB_ABS_WABS) det = wabs_survives_w(s_wabs, w[i], e[i], P.nH, RK, cell_key, draw0 + i);
  else if (P.absorb_kind == PXB_ABS_TABLE)
    det = survives_w(w[i], table_tau_per_nh(P, e[i]) * P.nH, RK, cell_key, draw0 + i);
  out[i] = det ? 1 : 0;
}

extern "C" int pxb_absorb_decide_words(const pxb_project_params* params, int64_t n,
                                       const double* energy, const uint32_t* w, uint64_t cell_key,
                                       uint64_t draw0, uint8_t* survive_out, void* stream) {
  int rc = pxb_check_project_params(params);
  if (rc) return rc;
  PXB_REQUIRE(n >= 0, "bad n");
  if (n == 0) return PXB_OK;
  PXB_REQUIRE(energy && w && survive_out, "NULL argument");
  k_absorb_decide_words<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      *params, pxb_make_keys(params->seed), n, energy, w, cell_key, draw0, survive_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_absorb_transmission(const pxb_project_params* params, int64_t n,
                                       const double* energy, double* trans_out, void* stream) {
  int rc = pxb_check_project_params(params);
  if (rc) return rc;
  PXB_REQUIRE(n >= 0, "bad n");
  if (n == 0) return PXB_OK;
  PXB_REQUIRE(energy && trans_out, "NULL argument");
  k_absorb_transmission<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      *params, n, energy, trans_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}
