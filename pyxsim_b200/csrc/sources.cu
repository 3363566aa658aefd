// sources.cu -- power-law and line source models (closed-form samplers).
//
// Reference: PowerLawSourceModel.process_data photons branch,
// pyxsim/source_models/power_law_sources.py:186-250, and LineSourceModel.process_data,
// pyxsim/source_models/line_sources.py:193-234.
#include "tiles.cuh"

// ---------------------------------------------------------------------------------------
// power law
// ---------------------------------------------------------------------------------------
struct PlCoef {
  double a;      // ei^(1-alpha)
  double nf;     // norm_fac = e0^alpha * (ef^(1-alpha) - ei^(1-alpha))   (power_law_sources.py:214)
  double inv;    // 1/(1-alpha); 0 flags alpha == 1
  double lognf;  // alpha == 1: log(ef/ei)
};

__device__ __forceinline__ PlCoef pl_coef(const pxb_powerlaw_cells& C, double alpha) {
  PlCoef k;
  if (alpha == 1.0) {
    k.a = C.emin; k.nf = C.emax / C.emin; k.inv = 0.0; k.lognf = log(k.nf);
  } else {
    const double om = 1.0 - alpha;
    k.a = pow(C.emin, om);
    k.nf = pow(C.e0, alpha) * (pow(C.emax, om) - k.a);
    k.inv = 1.0 / om;
    k.lognf = 0.0;
  }
  return k;
}

__device__ __forceinline__ double pl_lambda(const pxb_powerlaw_cells& C, int64_t i) {
  const double alpha = C.alpha ? C.alpha[i] : C.alpha_const;
  const double etoalpha = pow(C.e0, alpha);
  // K_fac, power_law_sources.py:198-203
  double kfac;
  if (alpha == 2.0) kfac = log(C.emax / C.emin) * etoalpha;
  else kfac = (pow(C.emax, 2.0 - alpha) - pow(C.emin, 2.0 - alpha)) * etoalpha / (2.0 - alpha);
  const double K = C.lum[i] / kfac;
  // Nph, :211-219
  double nph;
  if (alpha == 1.0) nph = log(C.emax / C.emin) * (K * etoalpha);
  else nph = (pow(C.emax, 1.0 - alpha) - pow(C.emin, 1.0 - alpha)) * (K * etoalpha) / (1.0 - alpha);
  nph *= C.spectral_norm * C.scale_factor;
  if (C.r2) nph /= C.r2[i];
  return nph;
}

__global__ void __launch_bounds__(256)
k_powerlaw_counts(const __grid_constant__ pxb_powerlaw_cells C, double* __restrict__ lambda_out,
                  int64_t* __restrict__ counts_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C.ncell) return;
  const double lam = pl_lambda(C, i);
  if (lambda_out) lambda_out[i] = lam;
  CellStream rs(C.seed, STREAM_POISSON, cell_gid(C, i));
  counts_out[i] = poisson_draw(lam, rs);
}

constexpr int SRC_TILE = 1024;

__global__ void __launch_bounds__(256)
k_powerlaw_sample(const __grid_constant__ pxb_powerlaw_cells C,
                  const __grid_constant__ PhiloxKeys RK, int64_t n_active,
                  const int64_t* __restrict__ idx, const int64_t* __restrict__ poff,
                  int64_t n_photons, double* __restrict__ energies) {
  __shared__ int64_t s_off[SRC_TILE + 1];
  __shared__ int64_t s_c[2];
  __shared__ PlCoef s_k[SRC_TILE];
  const int64_t p0 = (int64_t)blockIdx.x * SRC_TILE;
  const int64_t p1 = min(p0 + (int64_t)SRC_TILE, n_photons);
  TileMap<SRC_TILE> tm;
  tm.build(poff, n_active, p0, p1, s_off, s_c);
  // sampler coefficients (three pow() each): once per cell of the tile, or once per CTA for a
  // constant photon index -- never per thread
  __shared__ PlCoef s_kc;
  if (!C.alpha) {
    if (threadIdx.x == 0) s_kc = pl_coef(C, C.alpha_const);
    __syncthreads();
  } else if (tm.nc > 0) {
    for (int q = threadIdx.x; q < tm.nc; q += blockDim.x)
      s_k[q] = pl_coef(C, C.alpha[idx[tm.c0 + q]]);
    __syncthreads();
  }
  for (int64_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
    const int64_t c = tm.cell_of(p, s_off);
    PlCoef k;
    if (!C.alpha) k = s_kc;
    else k = (tm.nc > 0) ? s_k[c - tm.c0] : pl_coef(C, C.alpha[idx[c]]);
    const uint64_t gid = cell_gid(C, idx[c]);
    const uint64_t draw = (uint64_t)(p - __ldg(poff + c));
    // photons 2m and 2m+1 of a cell use the two halves of one Philox block
    const Philox4 r = pxb_rand4(RK, STREAM_POWERLAW, gid, draw >> 1);
    const double u = (draw & 1) ? u53(r.z, r.w) : u53(r.x, r.y);
    // x**y as exp(y log x): both bases are positive and |y log x| stays of order one (it is the
    // log of an energy ratio), so this is within a few ulp of pow() at a third of its cost
    double e;
    if (k.inv == 0.0) e = k.a * exp(u * k.lognf);      // ei * (ef/ei)**u, :233-234
    else e = exp(k.inv * log(k.a + u * k.nf));         // :236-237
    st_stream(energies + p, e * C.scale_factor);
  }
}

extern "C" int pxb_powerlaw_counts(const pxb_powerlaw_cells* cells, double* lambda_out,
                                   int64_t* counts_out, void* stream) {
  PXB_REQUIRE(cells && counts_out, "NULL argument");
  PXB_REQUIRE(cells->ncell >= 0, "negative ncell");
  if (cells->ncell == 0) return PXB_OK;
  PXB_REQUIRE(cells->lum, "luminosity array missing");
  k_powerlaw_counts<<<(unsigned)((cells->ncell + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      *cells, lambda_out, counts_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_powerlaw_sample(const pxb_powerlaw_cells* cells, int64_t n_active,
                                   const int64_t* idx, const int64_t* nph, const int64_t* poff,
                                   int64_t n_photons, double* energies, void* stream) {
  (void)nph;
  PXB_REQUIRE(cells, "NULL argument");
  if (n_active == 0 || n_photons == 0) return PXB_OK;
  PXB_REQUIRE(idx && poff && energies, "NULL argument");
  const int64_t ntiles = (n_photons + SRC_TILE - 1) / SRC_TILE;
  PXB_REQUIRE(ntiles < (1LL << 31), "too many photons for one call");
  k_powerlaw_sample<<<(unsigned)ntiles, 256, 0, (cudaStream_t)stream>>>(*cells, pxb_make_keys(cells->seed),
                                                                       n_active, idx, poff,
                                                                       n_photons, energies);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

// ---------------------------------------------------------------------------------------
// line
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_line_counts(const __grid_constant__ pxb_line_cells C, double* __restrict__ lambda_out,
              int64_t* __restrict__ counts_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C.ncell) return;
  double lam = C.rate[i] * C.spectral_norm * C.scale_factor;   // line_sources.py:205
  if (C.r2) lam /= C.r2[i];
  if (lambda_out) lambda_out[i] = lam;
  CellStream rs(C.seed, STREAM_POISSON, cell_gid(C, i));
  counts_out[i] = poisson_draw(lam, rs);
}

__global__ void __launch_bounds__(256)
k_line_sample(const __grid_constant__ pxb_line_cells C, const __grid_constant__ PhiloxKeys RK,
              int64_t n_active,
              const int64_t* __restrict__ idx, const int64_t* __restrict__ poff,
              int64_t n_photons, double* __restrict__ energies) {
  __shared__ int64_t s_off[SRC_TILE + 1];
  __shared__ int64_t s_c[2];
  const int64_t p0 = (int64_t)blockIdx.x * SRC_TILE;
  const int64_t p1 = min(p0 + (int64_t)SRC_TILE, n_photons);
  TileMap<SRC_TILE> tm;
  tm.build(poff, n_active, p0, p1, s_off, s_c);
  for (int64_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
    const int64_t c = tm.cell_of(p, s_off);
    const int64_t i = idx[c];
    const double sigma = C.sigma ? __ldg(C.sigma + i) : C.sigma_const;
    const uint64_t draw = (uint64_t)(p - __ldg(poff + c));
    // one Philox block -> two normals: photon pairs (2m, 2m+1) share block m
    const Philox4 r = pxb_rand4(RK, STREAM_LINE, cell_gid(C, i), draw >> 1);
    double n0, n1;
    normal2(r, n0, n1);
    const double g = (draw & 1) ? n1 : n0;
    st_stream(energies + p, (C.e0 + sigma * g) * C.scale_factor);   // :215-229
  }
}

extern "C" int pxb_line_counts(const pxb_line_cells* cells, double* lambda_out,
                               int64_t* counts_out, void* stream) {
  PXB_REQUIRE(cells && counts_out, "NULL argument");
  PXB_REQUIRE(cells->ncell >= 0, "negative ncell");
  if (cells->ncell == 0) return PXB_OK;
  PXB_REQUIRE(cells->rate, "emission array missing");
  k_line_counts<<<(unsigned)((cells->ncell + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      *cells, lambda_out, counts_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_line_sample(const pxb_line_cells* cells, int64_t n_active, const int64_t* idx,
                               const int64_t* nph, const int64_t* poff, int64_t n_photons,
                               double* energies, void* stream) {
  (void)nph;
  PXB_REQUIRE(cells, "NULL argument");
  if (n_active == 0 || n_photons == 0) return PXB_OK;
  PXB_REQUIRE(idx && poff && energies, "NULL argument");
  const int64_t ntiles = (n_photons + SRC_TILE - 1) / SRC_TILE;
  PXB_REQUIRE(ntiles < (1LL << 31), "too many photons for one call");
  k_line_sample<<<(unsigned)ntiles, 256, 0, (cudaStream_t)stream>>>(*cells, pxb_make_keys(cells->seed),
                                                                   n_active, idx, poff,
                                                                   n_photons, energies);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

// ---------------------------------------------------------------------------------------
// diagnostics: the discrete samplers on their own (statistical tests against scipy)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_rng_poisson(double lam, uint64_t seed, int64_t count, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  CellStream rs(seed, STREAM_POISSON, (uint64_t)i);
  out[i] = poisson_draw(lam, rs);
}

__global__ void __launch_bounds__(256)
k_rng_binomial(int64_t n, double p, uint64_t seed, int64_t count, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  CellStream rs(seed, STREAM_SPLIT, (uint64_t)i);
  out[i] = binomial_draw(n, p, rs);
}

extern "C" int pxb_rng_poisson(double lam, uint64_t seed, int64_t count, int64_t* out, void* stream) {
  PXB_REQUIRE(count >= 0, "bad count");
  if (count == 0) return PXB_OK;
  PXB_REQUIRE(out, "NULL argument");
  k_rng_poisson<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(lam, seed, count, out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_rng_binomial(int64_t n, double p, uint64_t seed, int64_t count, int64_t* out,
                                void* stream) {
  PXB_REQUIRE(count >= 0 && n >= 0, "bad count");
  if (count == 0) return PXB_OK;
  PXB_REQUIRE(out, "NULL argument");
  k_rng_binomial<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, p, seed, count, out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}
