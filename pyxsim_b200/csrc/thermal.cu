// thermal.cu -- thermal source model: table upload, expected counts + Poisson draws,
// per-cell spectra, inverse-CDF energy sampling.
//
// Reference path: ThermalSourceModel.process_data photons branch,
// pyxsim/source_models/thermal_sources/base.py:335-524, with the table lookup of
// pyxsim/spectral_models.py:34-47,63-82,429-462 and pyxsim/lib/interpolate.pyx.
#include <cstdlib>
#include "thermal.cuh"
#include <algorithm>
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

// Walker / Vose alias table of one non-negative table row (see DevTable::alias).  The sampled bin
// distribution is P(j) = (thr_j + sum_{k: other_k = j} (2^32 - thr_k)) / (2^32 nb): the row's
// normalised probabilities rounded to multiples of 2^-32 / nb (6e-14 at nb = 4000).  Bins of zero
// probability get thr = 0 and are nobody's `other`: they are never produced.
void pxb_build_alias(const double* p, int nb, uint2* out) {
  long double tot = 0.0L;
  for (int j = 0; j < nb; ++j) tot += p[j];
  for (int j = 0; j < nb; ++j) out[j] = make_uint2(0xFFFFFFFFu, (unsigned)j);   // "always j"
  if (!(tot > 0.0L)) return;                       // an empty row is never selected (weight 0)
  std::vector<long double> q((size_t)nb);
  std::vector<int> small, large;
  int jmax = 0;
  for (int j = 0; j < nb; ++j) {
    q[(size_t)j] = (long double)p[j] * nb / tot;
    if (p[j] > p[jmax]) jmax = j;
    (q[(size_t)j] < 1.0L ? small : large).push_back(j);
  }
  // smallest first (the zero bins while donors are plentiful): back() of a descending sort
  std::sort(small.begin(), small.end(), [&](int a, int b) { return q[(size_t)a] > q[(size_t)b]; });
  auto put = [&](int j, long double qq, int other) {
    long double t = floorl(qq * 4294967296.0L + 0.5L);
    if (t >= 4294967296.0L) out[j] = make_uint2(0xFFFFFFFFu, (unsigned)j);   // probability one: both branches give j
    else out[j] = make_uint2((unsigned)(t < 0.0L ? 0.0L : t), (unsigned)other);
  };
  while (!small.empty() && !large.empty()) {
    const int s = small.back();
    small.pop_back();
    const int l = large.back();
    put(s, q[(size_t)s], l);
    q[(size_t)l] = (q[(size_t)l] + q[(size_t)s]) - 1.0L;
    if (q[(size_t)l] < 1.0L) {
      large.pop_back();
      small.push_back(l);
    }
  }
  // leftovers are 1 up to rounding; a zero bin can only be left without donor through rounding
  // and is then sent to the row's strongest bin
  for (int s : small) {
    if (p[s] > 0.0) put(s, 1.0L, s);
    else out[s] = make_uint2(0u, (unsigned)jmax);
  }
}

// Diagnostic (host only, no GPU): the alias table libpxb builds for one table row, so a CPU test
// can check the implied bin probabilities against the row.
extern "C" int pxb_alias_table_host(const double* row, int nbins, uint32_t* thr_out, uint32_t* other_out) {
  PXB_REQUIRE(row && thr_out && other_out && nbins >= 1 && nbins < 65536, "bad argument");
  std::vector<uint2> t((size_t)nbins);
  pxb_build_alias(row, nbins, t.data());
  for (int j = 0; j < nbins; ++j) { thr_out[j] = t[(size_t)j].x; other_out[j] = t[(size_t)j].y; }
  return PXB_OK;
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
  // one alias table per (table, row) for the mixture sampler (entries are addressed by 32-bit
  // element indices: sampling.cuh, alias_row_index)
  if (t.nonneg && d.nbins < 65536 &&
      (size_t)(2 + d.nvar) * t.nrows * (((size_t)d.nbins + 1) & ~(size_t)1) < ((size_t)1 << 31)) {
    const int ntab = 2 + d.nvar;
    const size_t nb = (size_t)d.nbins;
    const size_t as = (nb + 1) & ~(size_t)1;
    t.alias_stride = (int)as;
    std::vector<uint2> al((size_t)ntab * t.nrows * as, make_uint2(0xFFFFFFFFu, 0u));
    for (int tb = 0; tb < ntab; ++tb) {
      const double* tab = tb == 0 ? d.cosmic : (tb == 1 ? d.metal : d.var + (size_t)(tb - 2) * t.nrows * nb);
      for (int r = 0; r < t.nrows; ++r)
        pxb_build_alias(tab + (size_t)r * nb, (int)nb, al.data() + ((size_t)tb * t.nrows + r) * as);
    }
    if ((rc = upload(m, al.data(), al.size() * sizeof(uint2), (const void**)&t.alias))) return rc;
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
    D.fast_ok = D.prim.alias != nullptr && (!D.has_fb || D.fb.alias != nullptr);
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

// =======================================================================================
// host entry points
// =======================================================================================
int pxb_check_cells(const pxb_model* model, const pxb_thermal_cells* c) {
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
  int rc = pxb_check_cells(model, cells);
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
  int rc = pxb_check_cells(model, cells);
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
