// thermal.cu -- thermal source model: table upload, expected counts + Poisson draws,
// per-cell spectra, inverse-CDF energy sampling.
//
// Reference path: ThermalSourceModel.process_data photons branch,
// pyxsim/source_models/thermal_sources/base.py:335-524, with the table lookup of
// pyxsim/spectral_models.py:34-47,63-82,429-462 and pyxsim/lib/interpolate.pyx.
#include <cstdlib>
#include "thermal.cuh"
#include <new>
#include <vector>

// =======================================================================================
// host: model create / destroy
// =======================================================================================
static int upload(pxb_model* m, const void* host, size_t bytes, const void** dev_out) {
  void* d = nullptr;
  if (bytes == 0) { *dev_out = nullptr; return PXB_OK; }
  PXB_CUDA(cudaMalloc(&d, bytes));
  if (m->nalloc >= 32) { cudaFree(d); pxb_set_error("too many model allocations"); return PXB_ERR_INVALID; }
  m->allocs[m->nalloc++] = d;
  PXB_CUDA(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
  *dev_out = d;
  return PXB_OK;
}

static void row_sums(const double* tab, int64_t nrows, int nbins, std::vector<double>& out,
                     int* nonneg) {
  out.resize((size_t)nrows);
  for (int64_t r = 0; r < nrows; ++r) {
    long double s = 0.0L;
    const double* p = tab + r * nbins;
    for (int j = 0; j < nbins; ++j) {
      if (!(p[j] >= 0.0)) *nonneg = 0;
      s += p[j];
    }
    out[(size_t)r] = (double)s;
  }
}

static int build_table(pxb_model* m, const pxb_table_desc& d, DevTable& t) {
  t = DevTable{};
  if (d.nx == 0) return PXB_OK;
  PXB_REQUIRE(d.nx >= 2, "table needs at least 2 temperature nodes (got %d)", d.nx);
  PXB_REQUIRE(d.ny == 0 || d.ny >= 2, "2-D table needs at least 2 density nodes");
  PXB_REQUIRE(d.nbins >= 1, "nbins must be positive");
  PXB_REQUIRE(d.nvar >= 0 && d.nvar <= PXB_MAX_VAR, "nvar=%d exceeds PXB_MAX_VAR", d.nvar);
  PXB_REQUIRE(d.xnodes && d.cosmic && d.metal, "table pointers must not be NULL");
  PXB_REQUIRE(d.nvar == 0 || d.var, "var table missing");
  t.nx = d.nx; t.ny = d.ny; t.nbins = d.nbins; t.nvar = d.nvar;
  t.nrows = d.nx * (d.ny > 0 ? d.ny : 1);
  t.nonneg = 1;
  const size_t tb = (size_t)t.nrows * d.nbins * sizeof(double);
  int rc;
  if ((rc = upload(m, d.xnodes, d.nx * sizeof(double), (const void**)&t.xnodes))) return rc;
  if (d.ny > 0) {
    PXB_REQUIRE(d.ynodes, "ynodes missing");
    if ((rc = upload(m, d.ynodes, d.ny * sizeof(double), (const void**)&t.ynodes))) return rc;
  }
  if ((rc = upload(m, d.cosmic, tb, (const void**)&t.cosmic))) return rc;
  if ((rc = upload(m, d.metal, tb, (const void**)&t.metal))) return rc;
  if (d.nvar) if ((rc = upload(m, d.var, tb * d.nvar, (const void**)&t.var))) return rc;
  std::vector<double> rs;
  row_sums(d.cosmic, t.nrows, d.nbins, rs, &t.nonneg);
  if ((rc = upload(m, rs.data(), rs.size() * sizeof(double), (const void**)&t.rs_c))) return rc;
  row_sums(d.metal, t.nrows, d.nbins, rs, &t.nonneg);
  if ((rc = upload(m, rs.data(), rs.size() * sizeof(double), (const void**)&t.rs_m))) return rc;
  if (d.nvar) {
    row_sums(d.var, (int64_t)t.nrows * d.nvar, d.nbins, rs, &t.nonneg);
    if ((rc = upload(m, rs.data(), rs.size() * sizeof(double), (const void**)&t.rs_v))) return rc;
  }
  if (t.nonneg && d.nbins < 65536) {
    // per-row normalised CDFs + guide tables for the mixture sampler
    const int ntab = 2 + d.nvar;
    const size_t nb = (size_t)d.nbins;
    const size_t cs = (nb + 2) & ~(size_t)1;     // doubles per CDF row (nb+1 used)
    const size_t gs = (nb + 8) & ~(size_t)7;     // entries per guide row (>= nb + 1)
    t.cdf_stride = (int)cs;
    t.guide_stride = (int)gs;
    std::vector<double> cdf((size_t)ntab * t.nrows * cs, 1.0);
    std::vector<uint16_t> guide((size_t)ntab * t.nrows * gs, 0);
    for (int tb = 0; tb < ntab; ++tb) {
      const double* tab = tb == 0 ? d.cosmic : (tb == 1 ? d.metal : d.var + (size_t)(tb - 2) * t.nrows * nb);
      for (int r = 0; r < t.nrows; ++r) {
        const double* p = tab + (size_t)r * nb;
        double* c = cdf.data() + ((size_t)tb * t.nrows + r) * cs;
        uint16_t* g = guide.data() + ((size_t)tb * t.nrows + r) * gs;
        long double tot = 0.0L;
        for (size_t j = 0; j < nb; ++j) tot += p[j];
        long double run = 0.0L;
        c[0] = 0.0;
        for (size_t j = 0; j < nb; ++j) {
          run += p[j];
          c[j + 1] = tot > 0.0L ? (double)(run / tot) : 0.0;
        }
        if (tot > 0.0L) {
          // everything from the last non-empty bin on is exactly 1
          size_t last = nb;
          while (last > 0 && p[last - 1] == 0.0) --last;
          for (size_t j = last; j <= nb; ++j) c[j] = 1.0;
        }
        size_t j = 0;
        for (size_t k = 0; k < nb; ++k) {
          const double q = (double)k / (double)nb;
          while (j + 1 < nb && c[j + 1] <= q) ++j;
          g[k] = (uint16_t)j;
        }
        for (size_t k = nb; k < gs; ++k) g[k] = (uint16_t)(nb - 1);   // sentinel: g[k+1] is always valid
      }
    }
    if ((rc = upload(m, cdf.data(), cdf.size() * sizeof(double), (const void**)&t.rowcdf))) return rc;
    if ((rc = upload(m, guide.data(), guide.size() * sizeof(uint16_t), (const void**)&t.guide))) return rc;
  }
  return PXB_OK;
}

// unnormalised per-row prefix sums (photon and energy weighted) for the O(1) band sums
static int build_prefix(pxb_model* m, const pxb_table_desc& d, DevTable& t, const double* emid) {
  if (d.nx == 0 || !t.nonneg) return PXB_OK;
  const int ntab = 2 + d.nvar;
  const size_t nb = (size_t)d.nbins;
  std::vector<double> pn((size_t)ntab * t.nrows * (nb + 1)), pe(pn.size());
  for (int tb = 0; tb < ntab; ++tb) {
    const double* tab = tb == 0 ? d.cosmic : (tb == 1 ? d.metal : d.var + (size_t)(tb - 2) * t.nrows * nb);
    for (int r = 0; r < t.nrows; ++r) {
      const double* p = tab + (size_t)r * nb;
      double* a = pn.data() + ((size_t)tb * t.nrows + r) * (nb + 1);
      double* b = pe.data() + ((size_t)tb * t.nrows + r) * (nb + 1);
      long double sa = 0.0L, sb = 0.0L;
      a[0] = 0.0;
      b[0] = 0.0;
      for (size_t j = 0; j < nb; ++j) {
        sa += p[j];
        sb += (long double)emid[j] * p[j];
        a[j + 1] = (double)sa;
        b[j + 1] = (double)sb;
      }
    }
  }
  int rc;
  if ((rc = upload(m, pn.data(), pn.size() * sizeof(double), (const void**)&t.pre_n))) return rc;
  return upload(m, pe.data(), pe.size() * sizeof(double), (const void**)&t.pre_e);
}

extern "C" int pxb_model_destroy(pxb_model* m) {
  if (!m) return PXB_OK;
  for (int i = 0; i < m->nalloc; ++i) cudaFree(m->allocs[i]);
  delete m;
  return PXB_OK;
}

extern "C" int pxb_model_create(const pxb_model_desc* desc, pxb_model** out) {
  PXB_REQUIRE(desc && out, "NULL argument");
  *out = nullptr;
  pxb_model* m = new (std::nothrow) pxb_model();
  PXB_REQUIRE(m, "out of host memory");
  m->nalloc = 0;
  int rc = build_table(m, desc->primary, m->dev.prim);
  if (!rc) rc = build_table(m, desc->fallback, m->dev.fb);
  DevModel& D = m->dev;
  if (!rc) {
    if (D.prim.nx == 0) { pxb_set_error("primary table missing"); rc = PXB_ERR_INVALID; }
  }
  if (!rc) {
    D.has_fb = D.fb.nx > 0;
    if (D.has_fb && (D.fb.nbins != D.prim.nbins || D.fb.nvar != D.prim.nvar || D.fb.ny != 0)) {
      pxb_set_error("fallback table must be 1-D with the primary's nbins/nvar");
      rc = PXB_ERR_INVALID;
    }
    if (desc->density_dependent && D.prim.ny == 0) {
      pxb_set_error("density_dependent model needs a 2-D primary table");
      rc = PXB_ERR_INVALID;
    }
    if (!desc->density_dependent && D.prim.ny != 0) {
      pxb_set_error("2-D table given to a model that is not density dependent");
      rc = PXB_ERR_INVALID;
    }
    if (!desc->bin_edges || !desc->emid) { pxb_set_error("bin_edges/emid missing"); rc = PXB_ERR_INVALID; }
  }
  if (!rc) {
    D.log_t = desc->log_t; D.fb_log_t = desc->fallback_log_t;
    D.density_dependent = desc->density_dependent;
    D.binscale_log = desc->binscale_log; D.method = desc->method;
    D.nbins = D.prim.nbins; D.nvar = D.prim.nvar;
    D.K_per_keV = desc->K_per_keV;
    D.kT_lo = desc->kT_lo; D.kT_hi = desc->kT_hi; D.nH_lo = desc->nH_lo; D.nH_hi = desc->nH_hi;
    rc = upload(m, desc->bin_edges, (D.nbins + 1) * sizeof(double), (const void**)&D.edges);
    if (!rc) rc = upload(m, desc->emid, D.nbins * sizeof(double), (const void**)&D.emid);
    if (!rc) rc = upload(m, desc->ebins ? desc->ebins : desc->bin_edges, (D.nbins + 1) * sizeof(double),
                         (const void**)&D.ebins);
    D.fast_ok = D.prim.rowcdf != nullptr && (!D.has_fb || D.fb.rowcdf != nullptr);
    if (!rc) rc = build_prefix(m, desc->primary, D.prim, desc->emid);
    if (!rc && D.has_fb) rc = build_prefix(m, desc->fallback, D.fb, desc->emid);
    {
      const double* eb = desc->bin_edges;
      const int nb = D.nbins;
      const double de = (eb[nb] - eb[0]) / nb;
      bool uni = de > 0.0;
      for (int j = 0; j <= nb && uni; ++j)
        if (fabs(eb[j] - (eb[0] + j * de)) > 4e-16 * fmax(fabs(eb[j]), fabs(eb[nb]))) uni = false;
      D.edges_uniform = uni ? 1 : 0;
      D.edge0 = eb[0];
      D.dedge = de;
    }
  }
  if (rc) { pxb_model_destroy(m); return rc; }
  *out = m;
  return PXB_OK;
}

extern "C" int pxb_model_nonnegative(const pxb_model* m) {
  if (!m) return 0;
  return m->dev.prim.nonneg && (!m->dev.has_fb || m->dev.fb.nonneg);
}

// =======================================================================================
// kernels
// =======================================================================================
// total spectrum value of bin j for a cell (clipped at 0): base.py:456-460
__device__ __forceinline__ double tot_bin(const CellMix& mx, double Z, const double* eZ,
                                          int nvar, int j) {
  const DevTable* t = mx.t;
  const int nb = t->nbins;
  double c = 0.0, m = 0.0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (r < mx.nr) {
      const size_t o = (size_t)mx.row[r] * nb + j;
      c += __ldg(t->cosmic + o) * mx.w[r];
      m += __ldg(t->metal + o) * mx.w[r];
    }
  }
  double tot = c + Z * m;
  for (int k = 0; k < nvar; ++k) {
    double v = 0.0;
    const double* vt = t->var + (size_t)k * t->nrows * nb;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r < mx.nr) v += __ldg(vt + (size_t)mx.row[r] * nb + j) * mx.w[r];
    tot += eZ[k] * v;
  }
  return fmax(tot, 0.0);
}

constexpr int64_t COUNT_NEEDS_GENERAL = -1;

// ---- expected counts, fast path: one thread per cell, O(1) via row sums ----------------
__global__ void __launch_bounds__(256)
k_thermal_counts_fast(const DevModel M, const __grid_constant__ pxb_thermal_cells C,
                      double* __restrict__ lambda_out, int64_t* __restrict__ counts_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C.ncell) return;
  const double kT = ld_stream(C.kT + i);
  if (!cell_cut(C, i, kT)) {
    if (lambda_out) lambda_out[i] = 0.0;
    counts_out[i] = 0;
    return;
  }
  const double nH = C.nH ? ld_stream(C.nH + i) : 0.0;
  CellMix mx;
  cell_mix(M, kT, nH, mx);
  const double xh = cell_xh(C, i);
  const double Z = cell_metal(C, i, xh);
  const DevTable* t = mx.t;
  bool fast = t->nonneg && mx.interior && (Z >= 0.0);
  double sc = 0.0, sm = 0.0;
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (r < mx.nr) {
      sc += mx.w[r] * __ldg(t->rs_c + mx.row[r]);
      sm += mx.w[r] * __ldg(t->rs_m + mx.row[r]);
    }
  double s = sc + Z * sm;
  for (int k = 0; k < C.nvar; ++k) {
    double e = cell_elem(C, k, i, xh);
    fast = fast && (e >= 0.0);
    double sv = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r < mx.nr) sv += mx.w[r] * __ldg(t->rs_v + (size_t)k * t->nrows + mx.row[r]);
    s += e * sv;
  }
  if (!fast) {
    counts_out[i] = COUNT_NEEDS_GENERAL;
    return;
  }
  const double lam = s * cell_norm(C, i);
  if (lambda_out) lambda_out[i] = lam;
  CellStream rs(C.seed, STREAM_POISSON, cell_gid(C, i));
  counts_out[i] = poisson_draw(lam, rs);
}

// ---- expected counts, general path: a warp sums the clipped spectrum of each flagged cell
__global__ void __launch_bounds__(256)
k_thermal_counts_general(const DevModel M, const __grid_constant__ pxb_thermal_cells C,
                         double* __restrict__ lambda_out, int64_t* __restrict__ counts_out) {
  __shared__ double s_eZ[8][PXB_MAX_VAR];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t base = warp * 32;
  if (base >= C.ncell) return;
  const int64_t mine = base + lane;
  bool flagged = (mine < C.ncell) && (counts_out[mine] == COUNT_NEEDS_GENERAL);
  unsigned mask = __ballot_sync(0xffffffffu, flagged);
  while (mask) {
    const int b = __ffs(mask) - 1;
    mask &= mask - 1;
    const int64_t i = base + b;
    const double kT = C.kT[i];
    const double nH = C.nH ? C.nH[i] : 0.0;
    CellMix mx;
    cell_mix(M, kT, nH, mx);
    const double xh = cell_xh(C, i);
    const double Z = cell_metal(C, i, xh);
    __syncwarp();
    for (int k = lane; k < C.nvar; k += 32) s_eZ[wib][k] = cell_elem(C, k, i, xh);
    __syncwarp();
    double acc = 0.0;
    for (int j = lane; j < M.nbins; j += 32) acc += tot_bin(mx, Z, s_eZ[wib], C.nvar, j);
    acc = warp_sum(acc);
    if (lane == 0) {
      const double lam = acc * cell_norm(C, i);
      if (lambda_out) lambda_out[i] = lam;
      CellStream rs(C.seed, STREAM_POISSON, cell_gid(C, i));
      counts_out[i] = poisson_draw(lam, rs);
    }
  }
}

// ---- diagnostic: full clipped spectra ---------------------------------------------------
__global__ void __launch_bounds__(256)
k_thermal_spectra(const DevModel M, const __grid_constant__ pxb_thermal_cells C,
                  double* __restrict__ spec_out) {
  __shared__ double s_eZ[PXB_MAX_VAR];
  const int64_t i = blockIdx.x;
  const double kT = C.kT[i];
  const double nH = C.nH ? C.nH[i] : 0.0;
  CellMix mx;
  cell_mix(M, kT, nH, mx);
  const double xh = cell_xh(C, i);
  const double Z = cell_metal(C, i, xh);
  for (int k = threadIdx.x; k < C.nvar; k += blockDim.x) s_eZ[k] = cell_elem(C, k, i, xh);
  __syncthreads();
  const bool keep = cell_cut(C, i, kT);
  for (int j = threadIdx.x; j < M.nbins; j += blockDim.x)
    spec_out[(size_t)i * M.nbins + j] = keep ? tot_bin(mx, Z, s_eZ, C.nvar, j) : 0.0;
}

// ---- bit-exact interp1d_spec / interp2d_spec -------------------------------------------
// Operation order follows interpolate.pyx:36-53 / :100-138 with contraction disabled so the
// result equals the C code compiled without FMA.
__global__ void __launch_bounds__(256)
k_interp_spec(const DevTable T, int64_t n, const double* __restrict__ xv,
              const int32_t* __restrict__ xis, const double* __restrict__ yv,
              const int32_t* __restrict__ yis, double* __restrict__ c_out,
              double* __restrict__ m_out, double* __restrict__ v_out) {
  const int64_t i = blockIdx.x;
  const int xi = xis[i];
  const double x = xv[i];
  const double dxi = __ddiv_rn(1.0, __dsub_rn(T.xnodes[xi + 1], T.xnodes[xi]));
  const double xp = __dmul_rn(__dsub_rn(x, T.xnodes[xi]), dxi);
  const double xm = __dmul_rn(__dsub_rn(T.xnodes[xi + 1], x), dxi);
  const int nb = T.nbins;
  if (T.ny == 0) {
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
      const size_t o0 = (size_t)xi * nb + j, o1 = o0 + nb;
      c_out[(size_t)i * nb + j] =
          __dadd_rn(__dmul_rn(T.cosmic[o0], xm), __dmul_rn(T.cosmic[o1], xp));
      m_out[(size_t)i * nb + j] =
          __dadd_rn(__dmul_rn(T.metal[o0], xm), __dmul_rn(T.metal[o1], xp));
      for (int k = 0; k < T.nvar; ++k) {
        const double* vt = T.var + (size_t)k * T.nrows * nb;
        v_out[((size_t)k * n + i) * nb + j] =
            __dadd_rn(__dmul_rn(vt[o0], xm), __dmul_rn(vt[o1], xp));
      }
    }
  } else {
    const int yi = yis[i];
    const double y = yv[i];
    const double dyi = __ddiv_rn(1.0, __dsub_rn(T.ynodes[yi + 1], T.ynodes[yi]));
    const double yp = __dmul_rn(__dsub_rn(y, T.ynodes[yi]), dyi);
    const double ym = __dmul_rn(__dsub_rn(T.ynodes[yi + 1], y), dyi);
    const size_t z1 = (size_t)T.nx * yi + xi, z2 = z1 + T.nx, z3 = z1 + 1, z4 = z2 + 1;
    auto blend = [&](const double* tab, int j) {
      double a = __dmul_rn(__dmul_rn(tab[z1 * nb + j], xm), ym);
      double b = __dmul_rn(__dmul_rn(tab[z2 * nb + j], xm), yp);
      double c = __dmul_rn(__dmul_rn(tab[z3 * nb + j], xp), ym);
      double d = __dmul_rn(__dmul_rn(tab[z4 * nb + j], xp), yp);
      return __dadd_rn(__dadd_rn(__dadd_rn(a, b), c), d);
    };
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
      c_out[(size_t)i * nb + j] = blend(T.cosmic, j);
      m_out[(size_t)i * nb + j] = blend(T.metal, j);
      for (int k = 0; k < T.nvar; ++k)
        v_out[((size_t)k * n + i) * nb + j] = blend(T.var + (size_t)k * T.nrows * nb, j);
    }
  }
}

// ---- energy sampling, general path: one CTA per active cell -----------------------------
// The CTA blends + clips the cell's spectrum, warp-scans it into an (unnormalised) CDF in
// shared memory, then draws the cell's photons by binary search (np.interp semantics of
// base.py:480-482: uniform inside a bin, empty bins skipped).
constexpr int SAMPLE_THREADS = 256;

__device__ __forceinline__ double invert_cdf(const double* __restrict__ cdf, int nbins,
                                             double total, double u, const DevModel& M) {
  const double t = u * total;
  if (!(t < total)) return __ldg(M.edges + nbins);
  int lo = 0, hi = nbins;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] <= t) lo = mid; else hi = mid;
  }
  if (M.method == 1) return __ldg(M.emid + lo);
  const double c0 = cdf[lo], c1 = cdf[lo + 1];
  const double e0 = __ldg(M.edges + lo), e1 = __ldg(M.edges + lo + 1);
  return e0 + (t - c0) / (c1 - c0) * (e1 - e0);
}

__global__ void __launch_bounds__(SAMPLE_THREADS)
k_thermal_sample_cta(const DevModel M, const __grid_constant__ pxb_thermal_cells C,
                     int64_t n_active, const int64_t* __restrict__ idx,
                     const int64_t* __restrict__ nph, const int64_t* __restrict__ poff,
                     double* __restrict__ energies, const int64_t* __restrict__ list,
                     const int64_t* __restrict__ list_count) {
  extern __shared__ double smem[];
  double* cdf = smem;                       // nbins + 1
  double* s_eZ = smem + (M.nbins + 1);      // PXB_MAX_VAR
  double* s_wtot = s_eZ + PXB_MAX_VAR;      // nwarps
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int NW = SAMPLE_THREADS / 32;
  const int nbins = M.nbins;
  const int seg = ((nbins + NW - 1) / NW + 31) & ~31;  // bins per warp, multiple of 32

  // with a list only the listed active cells (those the mixture kernel declined) are done
  const int64_t n_work = list ? *list_count : n_active;
  for (int64_t kk = blockIdx.x; kk < n_work; kk += gridDim.x) {
    const int64_t k = list ? list[kk] : kk;
    const int64_t i = idx[k];
    const int64_t N = nph[k];
    const int64_t off = poff[k];
    const double kT = C.kT[i];
    const double nH = C.nH ? C.nH[i] : 0.0;
    CellMix mx;
    cell_mix(M, kT, nH, mx);
    const double xh = cell_xh(C, i);
    const double Z = cell_metal(C, i, xh);
    __syncthreads();  // previous cell's sampling is done with cdf/s_eZ
    for (int ke = threadIdx.x; ke < C.nvar; ke += blockDim.x) s_eZ[ke] = cell_elem(C, ke, i, xh);
    if (threadIdx.x == 0) cdf[0] = 0.0;
    __syncthreads();
    // pass 1: per-warp inclusive scan of its contiguous segment
    const int j0 = wid * seg, j1 = min(j0 + seg, nbins);
    double carry = 0.0;
    for (int jb = j0; jb < j1; jb += 32) {
      const int j = jb + lane;
      double v = (j < j1) ? tot_bin(mx, Z, s_eZ, C.nvar, j) : 0.0;
      v = warp_incl_scan(v, lane) + carry;
      if (j < j1) cdf[j + 1] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
    if (lane == 0) s_wtot[wid] = carry;
    __syncthreads();
    // pass 2: add the totals of the preceding warps
    double basev = 0.0;
    for (int w = 0; w < wid; ++w) basev += s_wtot[w];
    if (wid > 0)
      for (int j = j0 + lane; j < j1; j += 32) cdf[j + 1] += basev;
    __syncthreads();
    const double total = cdf[nbins];
    const uint64_t gid = cell_gid(C, i);
    // sampling: one Philox block -> two photons
    const int64_t npair = (N + 1) >> 1;
    for (int64_t p = threadIdx.x; p < npair; p += blockDim.x) {
      const Philox4 r = pxb_rand4(C.seed, STREAM_ENERGY, gid, (uint64_t)p);
      double e0 = invert_cdf(cdf, nbins, total, u53(r.x, r.y), M);
      if (M.binscale_log) e0 = exp10(e0);
      st_stream(energies + off + 2 * p, e0);
      if (2 * p + 1 < N) {
        double e1 = invert_cdf(cdf, nbins, total, u53(r.z, r.w), M);
        if (M.binscale_log) e1 = exp10(e1);
        st_stream(energies + off + 2 * p + 1, e1);
      }
    }
  }
}


// ---- energy sampling, mixture fast path ------------------------------------------------
// When every table entry, abundance and interpolation weight is non-negative the clip of
// base.py:460 is a no-op and the cell's normalised spectrum is a MIXTURE of the normalised
// table rows it blends: p_j = sum_c q_c * row_c[j]/rowsum_c with q_c = coef_c*rowsum_c/S.
// Sampling "pick component c with probability q_c, then invert row c's own CDF" has exactly
// the distribution of inverting the blended CDF (base.py:471-482), needs no per-cell CDF
// build, and only touches per-row CDFs + guide tables that are shared by all cells.
// A per-cell prep kernel computes the mixture weights once; the sampler is thread-per-photon.
#include "tiles.cuh"


struct FastCell {
  int z1;        // first row
  int meta;      // bits 0-2: number of rows, bit 3: fallback table, bit 4: eligible,
                 // bit 5: split mode (cnt valid) instead of probability mode (cum valid),
                 // bit 6: general split (cnt = cumulative counts per ROW, the split over the
                 //        row's tables is in the per-cell side array)
  union {
    double cum[3];     // cumulative row probabilities (last is implicitly 1)
    long long cnt[3];  // split mode: cumulative photon counts of the components
                       // (row0,cosmic), (row0,metal), (row1,cosmic) [, (row1,metal) = rest]
  };
};

__device__ __forceinline__ int fast_row(const DevTable* t, int z1, int nr, int r) {
  if (nr == 2) return z1 + r;
  return z1 + ((r & 1) ? t->nx : 0) + ((r >> 1) ? 1 : 0);   // order of cell_mix: z1, z2, z3, z4
}

// N = photons of the cell.  Plain 1-D tables without variable elements have only four mixture
// components; for those the cell's N photons are split over the components ONCE here by
// sequential binomial draws (a multinomial draw -- the same joint distribution as letting every
// photon pick its component independently), so the per-photon work is just the CDF inversion
// and one Philox block serves two photons.
// GENERAL = false: everything except the general split, which is only REQUESTED (meta bit 7) so
// that the few cells that want it can be handled by a compacted second pass instead of stalling
// the 31 other lanes of their warp; GENERAL = true: that second pass.
template <bool GENERAL>
__device__ __forceinline__ FastCell fast_setup(const DevModel& M, const pxb_thermal_cells& C,
                                            int64_t i, int64_t N, int64_t* tabcum,
                                            int split_factor = 6) {
  FastCell fc;
  const double kT = C.kT[i];
  const double nH = C.nH ? C.nH[i] : 0.0;
  CellMix mx;
  cell_mix(M, kT, nH, mx);
  const double xh = cell_xh(C, i);
  const double Z = cell_metal(C, i, xh);
  const DevTable* t = mx.t;
  bool ok = mx.interior && (Z >= 0.0);
  double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (r < mx.nr) s[r] = __ldg(t->rs_c + mx.row[r]) + Z * __ldg(t->rs_m + mx.row[r]);
  for (int k = 0; k < C.nvar; ++k) {
    const double e = cell_elem(C, k, i, xh);
    ok = ok && (e >= 0.0);
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r < mx.nr) s[r] += e * __ldg(t->rs_v + (size_t)k * t->nrows + mx.row[r]);
  }
  double tot = 0.0;
  int last = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (r < mx.nr) {
      s[r] *= mx.w[r];
      tot += s[r];
      if (s[r] > 0.0) last = r;
    }
  ok = ok && (tot > 0.0);
  double run = 0.0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    run += s[r];
    fc.cum[r] = (r >= last) ? 1.0 : run / tot;
  }
  fc.z1 = mx.row[0];
  fc.meta = mx.nr | ((t == &M.fb) ? 8 : 0) | (ok ? 16 : 0);
  if (ok && C.nvar == 0 && mx.nr == 2 && N > 0) {
    double q[4];
    q[0] = mx.w[0] * __ldg(t->rs_c + mx.row[0]);
    q[1] = mx.w[0] * Z * __ldg(t->rs_m + mx.row[0]);
    q[2] = mx.w[1] * __ldg(t->rs_c + mx.row[1]);
    q[3] = mx.w[1] * Z * __ldg(t->rs_m + mx.row[1]);
    CellStream rs(C.seed, STREAM_SPLIT, cell_gid(C, i));
    double rest = q[0] + q[1] + q[2] + q[3];
    int64_t left = N, acc = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double pc = rest > 0.0 ? fmin(fmax(q[c] / rest, 0.0), 1.0) : 0.0;
      const int64_t nk = binomial_draw(left, pc, rs);
      left -= nk;
      acc += nk;
      rest -= q[c];
      fc.cnt[c] = acc;
    }
    if (!(q[3] > 0.0) && left > 0) {
      // rounding left photons for an empty last component: give them to the last non-empty one
      int last_ne = q[2] > 0.0 ? 2 : (q[1] > 0.0 ? 1 : 0);
      for (int c = last_ne; c < 3; ++c) fc.cnt[c] += left;
    }
    fc.meta |= 32;
  } else if (ok && tabcum && C.nvar > 0 && N >= split_factor * (int64_t)(mx.nr * (2 + C.nvar))) {
    if (!GENERAL) {
      fc.meta |= 128;
      return fc;
    }
    // Any other table shape (2-D tables blend four rows; variable elements add tables): the
    // same multinomial split done hierarchically -- N over the rows, then every row's share
    // over its tables -- with the within-row cumulative counts in the side array.  Only with
    // variable elements (whose per-photon component choice loops over the element tables; a
    // binomial draw costs about as much as ten such choices, hence the threshold); without them
    // the per-photon choice is cheap and the sampler is bound by the CDF gathers anyway
    // (measured on the 2-D table configuration: the split made it slower).
    const int ntab = 2 + C.nvar;
    CellStream rs(C.seed, STREAM_SPLIT, cell_gid(C, i));
    double rest = tot;
    int64_t left = N, acc = 0;
    int64_t nrow[4] = {0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r < mx.nr) {
        int64_t nk;
        if (r == last) nk = left;
        else if (r > last) nk = 0;
        else nk = binomial_draw(left, rest > 0.0 ? fmin(fmax(s[r] / rest, 0.0), 1.0) : 0.0, rs);
        left -= nk;
        rest -= s[r];
        nrow[r] = nk;
        acc += nk;
      }
      if (r < 3) fc.cnt[r] = (r < mx.nr) ? acc : N;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r >= mx.nr) continue;
      const int row = mx.row[r];
      int64_t* tc = tabcum + r * ntab;
      // the row's component weights: cosmic, Z * metal, e_k * var_k; last non-zero one takes
      // whatever the conditional draws leave
      double A = 0.0;
      int last_t = 0;
      for (int tb = 0; tb < ntab; ++tb) {
        const double a = tb == 0 ? __ldg(t->rs_c + row)
                       : tb == 1 ? Z * __ldg(t->rs_m + row)
                                 : cell_elem(C, tb - 2, i, xh) * __ldg(t->rs_v + (size_t)(tb - 2) * t->nrows + row);
        A += a;
        if (a > 0.0) last_t = tb;
      }
      int64_t lt = nrow[r], at = 0;
      for (int tb = 0; tb < ntab; ++tb) {
        const double a = tb == 0 ? __ldg(t->rs_c + row)
                       : tb == 1 ? Z * __ldg(t->rs_m + row)
                                 : cell_elem(C, tb - 2, i, xh) * __ldg(t->rs_v + (size_t)(tb - 2) * t->nrows + row);
        int64_t nk;
        if (tb == last_t) nk = lt;
        else if (tb > last_t || lt == 0) nk = 0;
        else nk = binomial_draw(lt, A > 0.0 ? fmin(fmax(a / A, 0.0), 1.0) : 0.0, rs);
        lt -= nk;
        A -= a;
        at += nk;
        tc[tb] = at;
      }
    }
    fc.meta |= 64;
  }
  return fc;
}

// n / d for 0 <= n < d <= 1 (a position inside a CDF step): single-precision reciprocal seed,
// two Newton steps and one residual correction in fp64 -- within one ulp of the IEEE quotient at
// about a third of its instructions.  Steps too small for the fp32 seed take the IEEE division.
__device__ __forceinline__ double cdf_fraction(double n, double d) {
  if (d < 1.0e-30) return n / d;
  float rf;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)d));
  double r = (double)rf;
  r = fma(r, fma(-d, r, 1.0), r);
  r = fma(r, fma(-d, r, 1.0), r);
  const double q = n * r;
  return fma(fma(-q, d, n), r, q);
}

// invert one table row's CDF at u2 (staged rows gather from shared memory)

// Rows staged in shared memory by the persistent sampler: CDFs and guides of rows (tag, tag+1)
// of the cosmic and metal tables of the primary table (slot = 2*table + (row - tag)).
struct StagedRows {
  int tag;                 // first staged row, -1: nothing staged
  const double* cdf;       // [4][cdf_stride]
  const uint16_t* guide;   // [4][guide_stride]
};

template <int SPEC = 0>
__device__ __forceinline__ double invert_row(const DevModel& M, const DevTable* t, int tab, int row,
                                             double u2, const StagedRows& S, bool prim);

__device__ __forceinline__ double fast_draw(const DevModel& M, const pxb_thermal_cells& C,
                                            const FastCell& fc, int64_t i, double u1, double u2,
                                            const StagedRows& S) {
  const DevTable* t = (fc.meta & 8) ? &M.fb : &M.prim;
  const int nr = fc.meta & 7;
  // row
  int r = (u1 >= fc.cum[0]) + (u1 >= fc.cum[1]) + (u1 >= fc.cum[2]);
  r = min(r, nr - 1);
  const double lo = r > 0 ? fc.cum[r - 1] : 0.0;
  const double hi = r < 3 ? fc.cum[r] : 1.0;
  const int row = fast_row(t, fc.z1, nr, r);
  // table within the row: cosmic, metal, var_k with weights rs_c, Z rs_m, eZ_k rs_v
  const double xh = cell_xh(C, i);
  const double a0 = __ldg(t->rs_c + row), a1 = cell_metal(C, i, xh) * __ldg(t->rs_m + row);
  int tab = 0;
  if (C.nvar == 0) {
    // v = (u1 - lo)/(hi - lo) is the within-row uniform; metal iff v*(a0+a1) >= a0
    tab = (a1 > 0.0 && ((u1 - lo) * (a0 + a1) >= a0 * (hi - lo) || !(a0 > 0.0))) ? 1 : 0;
  } else {
    const double v = (u1 - lo) / (hi - lo);
    double A = a0 + a1;
    for (int k = 0; k < C.nvar; ++k)
      A += cell_elem(C, k, i, xh) * __ldg(t->rs_v + (size_t)k * t->nrows + row);
    double tgt = v * A;
    int chosen = -1;
    bool done = false;
    if (a0 > 0.0) { chosen = 0; if (tgt < a0) done = true; else tgt -= a0; }
    if (!done && a1 > 0.0) { chosen = 1; if (tgt < a1) done = true; else tgt -= a1; }
    for (int k = 0; k < C.nvar && !done; ++k) {
      const double a = cell_elem(C, k, i, xh) * __ldg(t->rs_v + (size_t)k * t->nrows + row);
      if (a > 0.0) { chosen = 2 + k; if (tgt < a) done = true; else tgt -= a; }
    }
    tab = max(chosen, 0);
  }
  return invert_row(M, t, tab, row, u2, S, !(fc.meta & 8));
}

// SPEC = 1: compiled for inverse-CDF sampling on uniformly spaced linear bins
template <int SPEC>
__device__ __forceinline__ double invert_row(const DevModel& M, const DevTable* t, int tab, int row,
                                             double u2, const StagedRows& S, bool prim) {
  const int nb = t->nbins;
  const size_t ro = (size_t)tab * t->nrows + row;
  const double* cdf = t->rowcdf + ro * (size_t)t->cdf_stride;
  const uint16_t* g = t->guide + ro * (size_t)t->guide_stride;
  const unsigned dr = (unsigned)(row - S.tag);
  if (S.tag >= 0 && tab < 2 && dr < 2u && prim) {   // staged: shared-memory gathers
    const int slot = tab * 2 + (int)dr;
    cdf = S.cdf + (size_t)slot * t->cdf_stride;
    g = S.guide + (size_t)slot * t->guide_stride;
  }
  // the bin lies between the guide entries of quantiles k/nb and (k+1)/nb: bisect that range
  // (a linear walk would make the whole warp wait for its unluckiest lane)
  const int k = min((int)(u2 * nb), nb - 1);
  int j = g[k];
  int jhi = g[k + 1];
  while (jhi > j) {                      // largest j in [j, jhi] with cdf[j] <= u2
    const int mid = (j + jhi + 1) >> 1;
    if (cdf[mid] <= u2) j = mid; else jhi = mid - 1;
  }
  double2 c;
  c.x = cdf[j];
  c.y = cdf[j + 1];
  if (SPEC == 0 && M.method == 1) return __ldg(M.emid + j);
  double e0, e1;
  if (SPEC == 1 || M.edges_uniform) {
    e0 = M.edge0 + j * M.dedge;
    e1 = M.edge0 + (j + 1) * M.dedge;
  } else {
    e0 = __ldg(M.edges + j);
    e1 = __ldg(M.edges + j + 1);
  }
  return e0 + cdf_fraction(u2 - c.x, c.y - c.x) * (e1 - e0);
}

// invert_row for a row whose CDF and guide table sit in shared memory (32-bit addressing, no
// per-photon pointer selection): s_cdf/s_guide point at the staged slot
template <int SPEC = 0>
__device__ __forceinline__ double invert_row_staged(const DevModel& M, int nb, const double* s_cdf,
                                                    const uint16_t* s_guide, double u2) {
  const int k = min((int)(u2 * nb), nb - 1);
  int j = s_guide[k];
  int jhi = s_guide[k + 1];
  while (jhi > j) {
    const int mid = (j + jhi + 1) >> 1;
    if (s_cdf[mid] <= u2) j = mid; else jhi = mid - 1;
  }
  if (SPEC == 0 && M.method == 1) return __ldg(M.emid + j);
  const double ca = s_cdf[j], cb = s_cdf[j + 1];
  double e0, de;
  if (SPEC == 1 || M.edges_uniform) {
    e0 = M.edge0 + j * M.dedge;
    de = (M.edge0 + (j + 1) * M.dedge) - e0;
  } else {
    e0 = __ldg(M.edges + j);
    de = __ldg(M.edges + j + 1) - e0;
  }
  return e0 + cdf_fraction(u2 - ca, cb - ca) * de;
}

// ---- mixture sampler with TMA-staged table rows -----------------------------------------
// Persistent CTAs (one per SM).  A CTA takes chunks of 32768 consecutive photons; the rows the
// chunk's first cell blends (t, t+1 of the cosmic and metal tables: 4 x (32 KB CDF + 8 KB
// guide) at nbins = 4000) are brought into shared memory with four+four cp.async.bulk copies
// completing on one mbarrier, and stay there across chunks while the temperature bin does not
// change.  Inside a chunk every warp owns 128-photon tiles (no CTA barriers); photons whose
// component lives in a staged row gather from shared memory, all others from L1/L2.
constexpr int STG_THREADS = 1024;
constexpr int STG_WARPS = STG_THREADS / 32;
constexpr int STG_WTILE = 128;
constexpr int STG_CHUNK = 32768;
constexpr int STG_TILES_PER_CHUNK = STG_CHUNK / STG_WTILE;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// Per-cell preparation for the mixture sampler: one thread per ACTIVE cell computes the cell's
// mixture weights once (FastCell, 32 B) and records, for every 128-photon tile that starts
// inside the cell's photon range, that the tile starts in this cell -- so the sampler neither
// repeats the table-node search per tile nor searches the offsets.  Cells the mixture path
// cannot take (extrapolating, negative abundance) are appended to the general kernel's list.
__global__ void __launch_bounds__(256)
k_thermal_cellprep(const DevModel M, const __grid_constant__ pxb_thermal_cells C,
                   int64_t n_active, const int64_t* __restrict__ idx,
                   const int64_t* __restrict__ poff, FastCell* __restrict__ fcs,
                   int64_t* __restrict__ tile_cell, int64_t* __restrict__ list,
                   unsigned long long* __restrict__ list_count, int64_t* __restrict__ tabcum,
                   int64_t* __restrict__ glist, unsigned long long* __restrict__ glist_count,
                   int split_factor) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool want_general = false;
  if (k < n_active) {
    const int64_t a = poff[k], b = poff[k + 1];
    FastCell fc = fast_setup<false>(M, C, idx[k], b - a, tabcum, split_factor);
    want_general = (fc.meta & 128) != 0;
    fc.meta &= ~128;
    fcs[k] = fc;
    if (!(fc.meta & 16) && b > a) {
      const unsigned long long slot = atomicAdd(list_count, 1ull);
      list[slot] = k;
    }
    for (int64_t t = (a + STG_WTILE - 1) / STG_WTILE; t * STG_WTILE < b; ++t) tile_cell[t] = k;
  }
  // cells that want the general split: one warp-aggregated append
  const unsigned bal = __ballot_sync(0xffffffffu, want_general);
  if (bal) {
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(glist_count, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (want_general) glist[base + __popc(bal & ((1u << lane) - 1u))] = k;
  }
}

// second pass of the cell preparation: the general multinomial split for the cells that asked
// for it (all lanes busy with binomial draws instead of one in a warp)
__global__ void __launch_bounds__(128)
k_thermal_cellsplit(const DevModel M, const __grid_constant__ pxb_thermal_cells C,
                    const int64_t* __restrict__ idx, const int64_t* __restrict__ poff,
                    FastCell* __restrict__ fcs, int64_t* __restrict__ tabcum, int tab_stride,
                    const int64_t* __restrict__ glist,
                    const unsigned long long* __restrict__ glist_count) {
  const unsigned long long n = *glist_count;
  for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (unsigned long long)gridDim.x * blockDim.x) {
    const int64_t k = glist[j];
    FastCell fc = fast_setup<true>(M, C, idx[k], poff[k + 1] - poff[k], tabcum + k * tab_stride, 0);
    fcs[k] = fc;
  }
}

// SPEC = 1: the common case compiled with its options fixed -- a plain 1-D table without
// variable elements (every eligible cell is in four-component split mode), inverse-CDF sampling,
// uniformly spaced linear bins.
template <int SPEC>
__global__ void __launch_bounds__(STG_THREADS, 1)
k_thermal_sample_staged(const DevModel M, const __grid_constant__ pxb_thermal_cells C,
                        const __grid_constant__ PhiloxKeys RK, int64_t n_active, const int64_t* __restrict__ idx,
                        const int64_t* __restrict__ poff, int64_t n_photons,
                        double* __restrict__ energies, const FastCell* __restrict__ fcs,
                        const int64_t* __restrict__ tile_cell,
                        unsigned int* __restrict__ chunk_counter, int use_staging,
                        const int64_t* __restrict__ tabcum, int tab_stride, int guide_off,
                        int soff_off) {
  // shared memory: [4 staged CDF rows | 4 staged guide rows | per-warp cell offsets]; the two
  // byte offsets come precomputed from the host (kernel parameters are constant-bank operands)
  extern __shared__ __align__(128) unsigned char dyn[];
  const DevTable& T = M.prim;
  const size_t cdf_bytes = use_staging ? (size_t)T.cdf_stride * sizeof(double) : 0;
  const size_t gd_bytes = use_staging ? (size_t)T.guide_stride * sizeof(uint16_t) : 0;
  double* s_cdf = reinterpret_cast<double*>(dyn);
  uint16_t* s_guide = reinterpret_cast<uint16_t*>(dyn + guide_off);
  int64_t* s_off_all = reinterpret_cast<int64_t*>(dyn + soff_off);
  __shared__ uint64_t s_bar;
  __shared__ unsigned int s_chunk;

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int64_t* s_off = s_off_all + (size_t)wid * (STG_WTILE + 2);
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  int tag = -1;
  uint32_t parity = 0;
  const int64_t nchunks = (n_photons + STG_CHUNK - 1) / STG_CHUNK;
  const int64_t ntiles_all = (n_photons + STG_WTILE - 1) / STG_WTILE;

  for (;;) {
    if (threadIdx.x == 0) s_chunk = atomicAdd(chunk_counter, 1u);
    __syncthreads();     // also: every warp is done with the previous chunk's staged rows
    const int64_t chunk = s_chunk;
    if (chunk >= nchunks) break;
    const int64_t tile0 = chunk * STG_TILES_PER_CHUNK;
    const int ntile = (int)min((int64_t)STG_TILES_PER_CHUNK, ntiles_all - tile0);
    if (use_staging) {
      // temperature bin of the chunk's first cell decides what is staged
      const FastCell f0 = fcs[__ldg(tile_cell + tile0)];
      const int want = ((f0.meta & 7) == 2 && !(f0.meta & 8)) ? f0.z1 : -1;
      if (want >= 0 && want != tag) {
        if (threadIdx.x == 0) {
          mbar_expect_tx(&s_bar, (uint32_t)(4 * (cdf_bytes + gd_bytes)));
#pragma unroll
          for (int slot = 0; slot < 4; ++slot) {
            const size_t ro = (size_t)(slot >> 1) * T.nrows + want + (slot & 1);
            tma_bulk_g2s(s_cdf + (size_t)slot * T.cdf_stride, T.rowcdf + ro * T.cdf_stride,
                         (uint32_t)cdf_bytes, &s_bar);
            tma_bulk_g2s(s_guide + (size_t)slot * T.guide_stride, T.guide + ro * T.guide_stride,
                         (uint32_t)gd_bytes, &s_bar);
          }
        }
        mbar_wait(&s_bar, parity);
        parity ^= 1u;
        tag = want;
      }
    }
    const StagedRows S{tag, s_cdf, s_guide};

    for (int tl = wid; tl < ntile; tl += STG_WARPS) {
      const int64_t tile = tile0 + tl;
      const int64_t p0 = tile * STG_WTILE;
      const int64_t p1 = min(p0 + (int64_t)STG_WTILE, n_photons);
      const int64_t c0 = __ldg(tile_cell + tile);
      const int64_t c1 = (tile + 1 < ntiles_all) ? __ldg(tile_cell + tile + 1) : n_active - 1;
      const int nc = (c1 - c0 + 1 <= STG_WTILE + 1) ? (int)(c1 - c0 + 1) : 0;
      const int step0 = nc > 1 ? (1 << (31 - __clz(nc - 1))) : 0;
      __syncwarp();
      for (int q = lane; q <= nc && nc > 0; q += 32) s_off[q] = __ldg(poff + c0 + q);
      __syncwarp();
      // every lane handles PAIRS of adjacent photons: in split mode photons 2m and 2m+1 of a
      // cell share one Philox block, and the pair leaves as one 16-byte store
#pragma unroll 1
      for (int r = 0; r < STG_WTILE / 64; ++r) {
        const int64_t pa = p0 + r * 64 + 2 * lane;
        if (pa >= p1) continue;
        Philox4 blk{0u, 0u, 0u, 0u};
        uint64_t blk_cell = ~0ull, blk_idx = ~0ull;
        double e2[2];
        bool have[2] = {false, false};
        // the pair's cell: searched for the first photon, kept for the second unless it starts
        // the next cell
        int64_t first = 0, next = -1, i = 0, kc = 0;
        uint64_t gid = 0;
        int fz1 = 0, fmeta = 0;
        long long fcnt0 = 0, fcnt1 = 0, fcnt2 = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int64_t p = pa + h;
          if (p >= p1) continue;
          if (p >= next) {
            int64_t k;
            if (nc > 0) {
              int rel = 0;
              if (nc > 1) {
                for (int step = step0; step > 0; step >>= 1) {   // largest power of two < nc
                  const int q = rel + step;
                  if (q < nc && s_off[q] <= p) rel = q;
                }
              }
              k = c0 + rel;
              first = s_off[rel];
              next = s_off[rel + 1];
            } else {  // more cells than photons in the tile (never for active lists)
              int64_t lo = c0, hi = c1 + 1;
              while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(poff + mid) <= p) lo = mid; else hi = mid;
              }
              k = lo;
              first = __ldg(poff + k);
              next = __ldg(poff + k + 1);
            }
            const FastCell fc = fcs[k];
            fz1 = fc.z1;
            fmeta = fc.meta;
            fcnt0 = fc.cnt[0];
            fcnt1 = fc.cnt[1];
            fcnt2 = fc.cnt[2];
            kc = k;
            i = __ldg(idx + k);
            gid = cell_gid(C, i);
          }
          if (!(fmeta & 16)) continue;
          const uint64_t d = (uint64_t)(p - first);
          double e;
          if (SPEC == 1 || (fmeta & 32)) {
            if (!(blk_cell == gid && blk_idx == (d >> 1))) {
              blk = pxb_rand4(RK, STREAM_ENERGY, gid, d >> 1);
              blk_cell = gid;
              blk_idx = d >> 1;
            }
            const double u2 = (d & 1) ? u53(blk.z, blk.w) : u53(blk.x, blk.y);
            const long long dd = (long long)d;
            const int comp = (dd >= fcnt0) + (dd >= fcnt1) + (dd >= fcnt2);
            const bool prim = SPEC == 1 || !(fmeta & 8);
            const unsigned dr = (unsigned)(fz1 + (comp >> 1) - S.tag);
            if (prim && S.tag >= 0 && dr < 2u) {
              const int slot = (comp & 1) * 2 + (int)dr;
              e = invert_row_staged<SPEC>(M, T.nbins, s_cdf + slot * T.cdf_stride,
                                          s_guide + slot * T.guide_stride, u2);
            } else {
              e = invert_row<SPEC>(M, prim ? &M.prim : &M.fb, comp & 1, fz1 + (comp >> 1), u2, S, prim);
            }
          } else if (fmeta & 64) {
            // general split: row from the cumulative row counts, table from the side array
            if (!(blk_cell == gid && blk_idx == (d >> 1))) {
              blk = pxb_rand4(RK, STREAM_ENERGY, gid, d >> 1);
              blk_cell = gid;
              blk_idx = d >> 1;
            }
            const double u2 = (d & 1) ? u53(blk.z, blk.w) : u53(blk.x, blk.y);
            const long long dd = (long long)d;
            const int r = (dd >= fcnt0) + (dd >= fcnt1) + (dd >= fcnt2);
            const long long din = dd - (r == 0 ? 0 : (r == 1 ? fcnt0 : (r == 2 ? fcnt1 : fcnt2)));
            const int ntab = 2 + C.nvar;
            const int64_t* tc = tabcum + kc * tab_stride + r * ntab;
            int tab = 0;
            if (ntab <= 8) {
              for (int tb = 0; tb < ntab - 1; ++tb) tab += (din >= __ldg(tc + tb)) ? 1 : 0;
            } else {
              int lo = 0, hi = ntab - 1;     // first tb with din < tc[tb]
              while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (din >= __ldg(tc + mid)) lo = mid + 1; else hi = mid;
              }
              tab = lo;
            }
            const bool prim = !(fmeta & 8);
            const DevTable* tbl = prim ? &M.prim : &M.fb;
            e = invert_row(M, tbl, tab, fast_row(tbl, fz1, fmeta & 7, r), u2, S, prim);
          } else {
            const Philox4 rn = pxb_rand4(RK, STREAM_ENERGY, gid, d);
            const FastCell fc = fcs[kc];
            e = fast_draw(M, C, fc, i, u53(rn.x, rn.y), u53(rn.z, rn.w), S);
          }
          if (SPEC == 0 && M.binscale_log) e = exp10(e);
          e2[h] = e;
          have[h] = true;
        }
        if (have[0] && have[1]) {
          asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(energies + pa),
                       "d"(e2[0]), "d"(e2[1]));
        } else {
          if (have[0]) st_stream(energies + pa, e2[0]);
          if (have[1]) st_stream(energies + pa + 1, e2[1]);
        }
      }
    }
  }
}

// =======================================================================================
// host entry points
// =======================================================================================
static int check_cells(const pxb_model* model, const pxb_thermal_cells* c) {
  PXB_REQUIRE(model && c, "NULL model/cells");
  PXB_REQUIRE(c->ncell >= 0, "negative ncell");
  PXB_REQUIRE(c->nvar == model->dev.nvar, "cells.nvar=%d does not match the table's nvar=%d",
              c->nvar, model->dev.nvar);
  if (c->ncell > 0) {
    PXB_REQUIRE(c->kT && c->em, "kT/em arrays are required");
    PXB_REQUIRE(!model->dev.density_dependent || c->nH, "density dependent model needs nH");
  }
  return PXB_OK;
}

extern "C" int pxb_thermal_counts(const pxb_model* model, const pxb_thermal_cells* cells,
                                  double* lambda_out, int64_t* counts_out, void* stream) {
  int rc = check_cells(model, cells);
  if (rc) return rc;
  PXB_REQUIRE(counts_out, "counts_out is NULL");
  if (cells->ncell == 0) return PXB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = cells->ncell;
  const unsigned grid = (unsigned)((n + 255) / 256);
  k_thermal_counts_fast<<<grid, 256, 0, st>>>(model->dev, *cells, lambda_out, counts_out);
  PXB_LAUNCH_CHECK();
  const int64_t nwarp = (n + 31) / 32;
  const unsigned grid2 = (unsigned)((nwarp + 7) / 8);
  k_thermal_counts_general<<<grid2, 256, 0, st>>>(model->dev, *cells, lambda_out, counts_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_thermal_spectra(const pxb_model* model, const pxb_thermal_cells* cells,
                                   double* spec_out, void* stream) {
  int rc = check_cells(model, cells);
  if (rc) return rc;
  PXB_REQUIRE(spec_out, "spec_out is NULL");
  if (cells->ncell == 0) return PXB_OK;
  PXB_REQUIRE(cells->ncell < (1LL << 31), "too many cells for the diagnostic spectra kernel");
  k_thermal_spectra<<<(unsigned)cells->ncell, 256, 0, (cudaStream_t)stream>>>(model->dev, *cells,
                                                                            spec_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_interp_spec(const pxb_model* model, int64_t n, const double* x_vals,
                               const int32_t* x_is, const double* y_vals, const int32_t* y_is,
                               double* c_out, double* m_out, double* v_out, void* stream) {
  PXB_REQUIRE(model, "NULL model");
  if (n == 0) return PXB_OK;
  PXB_REQUIRE(n > 0 && n < (1LL << 31), "bad n");
  const DevTable& T = model->dev.prim;
  PXB_REQUIRE(x_vals && x_is && c_out && m_out, "NULL argument");
  PXB_REQUIRE(T.ny == 0 || (y_vals && y_is), "2-D table needs y_vals/y_is");
  PXB_REQUIRE(T.nvar == 0 || v_out, "v_out required when the table has variable elements");
  k_interp_spec<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(T, n, x_vals, x_is, y_vals, y_is,
                                                              c_out, m_out, v_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

// side array of the general split: cumulative counts per (row, table) of every active cell
static int split_stride(const DevModel& D) {
  if (D.nvar == 0) return 0;        // no variable elements: no general split (see fast_setup)
  return (D.density_dependent ? 4 : 2) * (2 + D.nvar);
}

extern "C" int pxb_thermal_sample_workspace_bytes(const pxb_model* model, int64_t n_active,
                                                  int64_t n_photons, int64_t* bytes) {
  PXB_REQUIRE(model && bytes && n_active >= 0 && n_photons >= 0, "bad argument");
  const int64_t ntiles = (n_photons + STG_WTILE - 1) / STG_WTILE;
  const int64_t stride = split_stride(model->dev);
  *bytes = (4 + n_active + ntiles + 1) * (int64_t)sizeof(int64_t) + n_active * (int64_t)sizeof(FastCell) +
           (stride ? n_active * (stride + 1) * (int64_t)sizeof(int64_t) : 0);
  return PXB_OK;
}

extern "C" int pxb_thermal_sample(const pxb_model* model, const pxb_thermal_cells* cells,
                                  int64_t n_active, const int64_t* idx, const int64_t* nph,
                                  const int64_t* poff, int64_t n_photons, double* energies,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = check_cells(model, cells);
  if (rc) return rc;
  if (n_active == 0 || n_photons == 0) return PXB_OK;
  PXB_REQUIRE(idx && nph && poff && energies, "NULL argument");
  PXB_REQUIRE(((uintptr_t)energies & 15) == 0, "energies must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const DevModel& D = model->dev;
  const size_t smem = ((size_t)D.nbins + 1 + PXB_MAX_VAR + SAMPLE_THREADS / 32) * sizeof(double);
  PXB_REQUIRE(smem <= 227 * 1024, "nbins=%d needs %zu B of shared memory for the per-cell CDF (max 227 KB)",
              D.nbins, smem);
  // attribute + occupancy query once per (process, shared-memory size), not per call
  static thread_local size_t cta_smem_set = 0;
  static thread_local int cta_per_sm = 1;
  if (cta_smem_set != smem) {
    PXB_CUDA(cudaFuncSetAttribute(k_thermal_sample_cta, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    int q = 0;
    PXB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, k_thermal_sample_cta,
                                                           SAMPLE_THREADS, smem));
    cta_per_sm = q < 1 ? 1 : q;
    cta_smem_set = smem;
  }
  const int per_sm = cta_per_sm;
  int64_t grid = (int64_t)pxb_sm_count() * per_sm;
  if (grid > n_active) grid = n_active;
  int64_t* list = nullptr;
  unsigned long long* count = nullptr;
  if (D.fast_ok) {
    int64_t need = 0;
    pxb_thermal_sample_workspace_bytes(model, n_active, n_photons, &need);
    PXB_REQUIRE(workspace && workspace_bytes >= need, "sample workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)need);
    // workspace: [count, chunk counter, pad, pad | list[n_active] | tile_cell[ntiles+1] | FastCell[n_active]
    //             | general-split side array [n_active][rows * tables] | general-split cell list [n_active]]
    const int64_t ntiles = (n_photons + STG_WTILE - 1) / STG_WTILE;
    count = (unsigned long long*)workspace;
    unsigned int* chunk_counter = (unsigned int*)((int64_t*)workspace + 2);
    list = (int64_t*)workspace + 4;
    int64_t* tile_cell = list + n_active;
    FastCell* fcs = (FastCell*)(tile_cell + ntiles + 1);
    const int tab_stride = split_stride(D);
    int64_t* tabcum = tab_stride ? (int64_t*)(fcs + n_active) : nullptr;
    int64_t* glist = tab_stride ? tabcum + n_active * (int64_t)tab_stride : nullptr;
    unsigned long long* glist_count = (unsigned long long*)workspace + 3;
    if (cells->split_factor < 0) tabcum = nullptr;   // per-photon component choice
    const int split_factor = cells->split_factor > 0 ? cells->split_factor : 6;
    PXB_CUDA(cudaMemsetAsync(workspace, 0, 4 * sizeof(int64_t), st));
    k_thermal_cellprep<<<(unsigned)((n_active + 255) / 256), 256, 0, st>>>(
        D, *cells, n_active, idx, poff, fcs, tile_cell, list, count, tabcum, glist, glist_count,
        split_factor);
    PXB_LAUNCH_CHECK();
    if (tabcum) {
      int64_t gs = (n_active + 127) / 128;
      if (gs > (int64_t)pxb_sm_count() * 16) gs = (int64_t)pxb_sm_count() * 16;
      k_thermal_cellsplit<<<(unsigned)gs, 128, 0, st>>>(D, *cells, idx, poff, fcs, tabcum, tab_stride,
                                                        glist, glist_count);
      PXB_LAUNCH_CHECK();
    }
    // rows are staged in shared memory when the model is a plain 1-D table whose four blended
    // rows fit beside the per-warp offsets; otherwise the same kernel gathers from L1/L2
    const size_t off_bytes = (size_t)STG_WARPS * (STG_WTILE + 2) * sizeof(int64_t);
    const size_t row_bytes = 4 * ((size_t)D.prim.cdf_stride * 8 + (size_t)D.prim.guide_stride * 2);
    const int use_staging = (!D.density_dependent && !D.has_fb && row_bytes + off_bytes <= 220 * 1024) ? 1 : 0;
    const size_t stg_smem = off_bytes + (use_staging ? row_bytes : 0);
    PXB_REQUIRE(n_photons < (int64_t)STG_CHUNK * ((1LL << 32) - 4096), "too many photons for one call");
    const bool common = D.method == 0 && !D.binscale_log && D.edges_uniform && D.nvar == 0 &&
                        !D.density_dependent && !D.has_fb && !cells->generic_kernels;
    auto kern = common ? k_thermal_sample_staged<1> : k_thermal_sample_staged<0>;
    static thread_local size_t stg_set[2] = {0, 0};
    if (stg_set[common ? 1 : 0] != stg_smem + 1) {
      PXB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stg_smem));
      stg_set[common ? 1 : 0] = stg_smem + 1;
    }
    const int64_t nchunks = (n_photons + STG_CHUNK - 1) / STG_CHUNK;
    int64_t g2 = pxb_sm_count();
    if (g2 > nchunks) g2 = nchunks;
    kern<<<(unsigned)g2, STG_THREADS, stg_smem, st>>>(
        D, *cells, pxb_make_keys(cells->seed), n_active, idx, poff, n_photons, energies, fcs, tile_cell,
        chunk_counter, use_staging,
        tabcum, tab_stride, (int)(use_staging ? row_bytes - 4 * (size_t)D.prim.guide_stride * 2 : 0),
        (int)(use_staging ? row_bytes : 0));
    PXB_LAUNCH_CHECK();
  }
  k_thermal_sample_cta<<<(unsigned)grid, SAMPLE_THREADS, smem, st>>>(
      D, *cells, n_active, idx, nph, poff, energies, list, (const int64_t*)count);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}
