// thermal.cuh -- device-side view of a thermal spectral table model and the per-cell
// "row mix" that every thermal kernel shares.
//
// Reference semantics being reproduced:
//   node index + weights ... pyxsim/spectral_models.py:35-37,64-69 and
//                            pyxsim/lib/interpolate.pyx:36-41,100-107
//   2-D row index .......... interpolate.pyx:112-115 (row = ntbins*y_i + x_i)
//   Pion IGM/CIE switch .... spectral_models.py:432-456 (divide by nH inside the 2-D table)
//   abundances ............. pyxsim/source_models/thermal_sources/base.py:395-421
//   combine + clip ......... base.py:456-460
#pragma once
#include "common.cuh"

struct DevTable {
  int nx, ny, nbins, nvar, nrows;
  int nonneg;             // every entry of cosmic/metal/var >= 0
  const double* xnodes;   // [nx]
  const double* ynodes;   // [ny]
  const double* cosmic;   // [nrows, nbins]
  const double* metal;
  const double* var;      // [nvar, nrows, nbins]
  const double* rs_c;     // row sums [nrows]
  const double* rs_m;
  const double* rs_v;     // [nvar, nrows]
  // mixture fast path (only when nonneg): one Walker alias table per table row,
  // [(2+nvar), nrows, alias_stride] (table order cosmic, metal, var_0..).  Entry j = {thr, other}:
  // a 64-bit uniform u picks bin j = floor(u nbins / 2^64); the top 32 bits of the remaining
  // fraction are compared with thr -- below: bin j, else bin `other`; the low 32 bits are the
  // position inside the chosen bin.  No search, no data-dependent loop.
  const uint2* alias;      // row stride alias_stride entries (even: rows are 16-byte multiples for TMA)
  int alias_stride;
  // band sums in O(1) (only when nonneg): unnormalised prefix sums over the bins of every row,
  // pre_n[tab][row][j] = sum_{q<j} T[q], pre_e[...] = sum_{q<j} emid_q T[q]; [(2+nvar), nrows, nbins+1]
  const double* pre_n;
  const double* pre_e;
};

struct DevModel {
  DevTable prim, fb;
  int has_fb;
  int log_t, fb_log_t, density_dependent, binscale_log, method;
  int nbins, nvar;
  double K_per_keV;
  double kT_lo, kT_hi, nH_lo, nH_hi;
  const double* edges;    // [nbins+1] (log10 when binscale_log)
  const double* ebins;    // [nbins+1] linear keV
  const double* emid;     // [nbins]
  int fast_ok;            // mixture sampler usable (nonneg tables, nbins < 65536)
  int edges_uniform;      // edges[j] == edge0 + j*dedge (to rounding): computed, not loaded
  double edge0, dedge;
};

struct pxb_model {
  DevModel dev;
  void* allocs[32];
  int nalloc;
};

// Up to four table rows with weights: spectrum_j = sum_r w[r] * T[row[r]][j]
struct CellMix {
  const DevTable* t;
  int nr;
  int row[4];
  double w[4];
  bool interior;  // all weights within [0,1]: no extrapolation beyond the table
};

__device__ __forceinline__ void cell_mix(const DevModel& M, double kT, double nH, CellMix& mx) {
  bool use_prim = true;
  if (M.density_dependent && M.has_fb) {
    use_prim = (kT >= M.kT_lo) && (kT <= M.kT_hi) && (nH >= M.nH_lo) && (nH <= M.nH_hi);
  }
  const DevTable* t = use_prim ? &M.prim : &M.fb;
  const int lt = use_prim ? M.log_t : M.fb_log_t;
  mx.t = t;
  double x = lt ? log10(kT * M.K_per_keV) : kT;
  int xi = node_index(t->xnodes, t->nx, x);
  double x0 = __ldg(t->xnodes + xi), x1 = __ldg(t->xnodes + xi + 1);
  double dxi = 1.0 / (x1 - x0);
  double xp = (x - x0) * dxi, xm = (x1 - x) * dxi;
  bool inter = (xp >= 0.0) && (xm >= 0.0);
  if (use_prim && M.density_dependent && t->ny > 0) {
    double y = log10(nH);
    int yi = node_index(t->ynodes, t->ny, y);
    double y0 = __ldg(t->ynodes + yi), y1 = __ldg(t->ynodes + yi + 1);
    double dyi = 1.0 / (y1 - y0);
    double yp = (y - y0) * dyi, ym = (y1 - y) * dyi;
    inter = inter && (yp >= 0.0) && (ym >= 0.0);
    double inh = 1.0 / nH;
    int z1 = t->nx * yi + xi;
    mx.nr = 4;
    mx.row[0] = z1;          mx.w[0] = xm * ym * inh;
    mx.row[1] = z1 + t->nx;  mx.w[1] = xm * yp * inh;
    mx.row[2] = z1 + 1;      mx.w[2] = xp * ym * inh;
    mx.row[3] = z1 + t->nx + 1; mx.w[3] = xp * yp * inh;
    inter = inter && (nH > 0.0);
  } else {
    mx.nr = 2;
    mx.row[0] = xi;      mx.w[0] = xm;
    mx.row[1] = xi + 1;  mx.w[1] = xp;
    mx.row[2] = 0; mx.w[2] = 0.0;
    mx.row[3] = 0; mx.w[3] = 0.0;
  }
  mx.interior = inter;
}

__device__ __forceinline__ bool cell_cut(const pxb_thermal_cells& C, int64_t i, double kT) {
  bool keep = (kT >= C.kT_min) && (kT <= C.kT_max);
  if (C.density) keep = keep && (ld_stream(C.density + i) < C.max_density);
  if (C.entropy) keep = keep && (ld_stream(C.entropy + i) > C.min_entropy);
  return keep;
}

__device__ __forceinline__ double cell_xh(const pxb_thermal_cells& C, int64_t i) {
  return C.xh ? ld_stream(C.xh + i) : C.xh_const;
}

__device__ __forceinline__ double cell_metal(const pxb_thermal_cells& C, int64_t i, double xh) {
  if (!C.zmet) return C.zmet_const;
  double fac = C.zmet_scale;
  if (C.zmet_by_xh) fac /= xh;
  return ld_stream(C.zmet + i) * fac;
}

__device__ __forceinline__ double cell_elem(const pxb_thermal_cells& C, int k, int64_t i,
                                            double xh) {
  if (!C.elem[k]) return C.elem_const[k];
  double fac = C.elem_scale[k];
  if (C.elem_by_xh[k]) fac /= xh;
  return ld_stream(C.elem[k] + i) * fac;
}

__device__ __forceinline__ double cell_norm(const pxb_thermal_cells& C, int64_t i) {
  double cnm = ld_stream(C.em + i) * C.spectral_norm;
  if (C.r2) cnm /= ld_stream(C.r2 + i);
  return cnm;
}

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

// host: argument checks shared by every entry point that takes (model, cells)
int pxb_check_cells(const pxb_model* model, const pxb_thermal_cells* c);
