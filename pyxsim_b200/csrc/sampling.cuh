// sampling.cuh -- energy sampling of the thermal source, mixture fast path (device code).
//
// Reference: per-cell CDF + inverse-CDF sampling, pyxsim/source_models/thermal_sources/base.py:
// 471-491,522-523.  When every table entry, abundance and interpolation weight is non-negative
// the clip of base.py:460 is a no-op and the cell's normalised spectrum is a MIXTURE of the
// normalised table rows it blends: p_j = sum_c q_c * row_c[j]/rowsum_c with
// q_c = coef_c*rowsum_c/S.  "Pick component c with probability q_c, then sample row c" has exactly
// the distribution of inverting the blended CDF, needs no per-cell CDF build, and only touches
// per-row alias tables that are shared by all cells.  A per-cell prep kernel computes the mixture
// weights (and, where it pays, splits the cell's photon count over the components by a multinomial
// draw) once; the photon kernels are thread-per-photon-pair.
#pragma once
#include "thermal.cuh"

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
                                            int split_factor = 40) {
  FastCell fc;
  const double kT = C.kT[i];
  const double nH = C.nH ? C.nH[i] : 0.0;
  CellMix mx;
  cell_mix(M, kT, nH, mx);
  const double xh = cell_xh(C, i);
  const double Z = cell_metal(C, i, xh);
  const DevTable* t = mx.t;
  // (a cell with 2^31 photons or more goes to the general kernel: the photon kernels compare a
  // photon's index inside its cell with the component counts in 32 bits)
  bool ok = mx.interior && (Z >= 0.0) && N < 0x7fffffffll;
  // (explicit multiply-adds: the threshold mode below re-accumulates the same sums and relies on
  // reaching exactly these totals)
  double s[4] = {0.0, 0.0, 0.0, 0.0}, sraw[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (r < mx.nr) s[r] = fma(Z, __ldg(t->rs_m + mx.row[r]), __ldg(t->rs_c + mx.row[r]));
  for (int k = 0; k < C.nvar; ++k) {
    const double e = cell_elem(C, k, i, xh);
    ok = ok && (e >= 0.0);
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r < mx.nr) s[r] = fma(e, __ldg(t->rs_v + (size_t)k * t->nrows + mx.row[r]), s[r]);
  }
  double tot = 0.0;
  int last = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (r < mx.nr) {
      sraw[r] = s[r];
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
    // variable elements and only for cells with many photons: a split photon shares its Philox
    // block with its pair partner, but the split costs one binomial draw per component, and the
    // per-photon choice from the stored thresholds below is cheap.  Without variable elements
    // the per-photon choice is one compare and the sampler is bound by the table gathers anyway
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
  } else if (!GENERAL && ok && tabcum && C.nvar > 0 && N > 0) {
    // Variable elements, too few photons for the split to pay: every photon picks its component
    // itself, from ONE ascending array of nr x ntab cumulative probabilities written here (the
    // side array's slot of the cell, as doubles) instead of re-deriving the element weights per
    // photon.  Component c = row r, table tb is chosen for th[c-1] <= u < th[c]; zero-weight
    // components have zero width, and everything from a row's last non-empty table on carries
    // the row's upper end, so rounding never selects an empty table.
    const int ntab = 2 + C.nvar;
    double* th = reinterpret_cast<double*>(tabcum);
    double c[4] = {0.0, 0.0, 0.0, 0.0}, scale[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const double lo = r > 0 ? fc.cum[r - 1] : 0.0;
      const double hi = r < 3 ? fc.cum[r] : 1.0;
      scale[r] = (r < mx.nr && sraw[r] > 0.0) ? (hi - lo) / sraw[r] : 0.0;
    }
    for (int tb = 0; tb < ntab; ++tb) {
      const double e = tb == 0 ? 1.0 : (tb == 1 ? Z : cell_elem(C, tb - 2, i, xh));
      const double* rs = tb == 0 ? t->rs_c : (tb == 1 ? t->rs_m : t->rs_v + (size_t)(tb - 2) * t->nrows);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (r >= mx.nr) continue;
        // the running sum of the row's table weights, accumulated exactly as sraw[r] was: from the
        // row's last non-empty table on it EQUALS sraw[r]
        c[r] = tb == 0 ? __ldg(rs + mx.row[r]) : fma(e, __ldg(rs + mx.row[r]), c[r]);
        const double lo = r > 0 ? fc.cum[r - 1] : 0.0;
        const double hi = r < 3 ? fc.cum[r] : 1.0;
        th[r * ntab + tb] = (c[r] >= sraw[r]) ? hi : fmin(fma(c[r], scale[r], lo), hi);
      }
    }
    fc.meta |= 256;
  }
  return fc;
}

// ---- sampling a table row through its alias table ---------------------------------------
// (w0, w1): 64 random bits.  Returns the bin; v receives 32 uniform bits for the position inside
// it.  u nb = j + frac with u = (w0 2^32 + w1) / 2^64: two 32x32->64 multiplies and a carry.
// The row is entries [off, off + nb) of tab (one 32-bit index, one scaled address).
__device__ __forceinline__ int alias_pick(const uint2* __restrict__ tab, unsigned off, unsigned nb,
                                          uint32_t w0, uint32_t w1, uint32_t& v) {
  const unsigned long long a = (unsigned long long)w0 * nb;
  const unsigned long long b = (unsigned long long)w1 * nb;
  const uint32_t alo = (uint32_t)a;
  const uint32_t top = alo + (uint32_t)(b >> 32);        // top 32 bits of the fraction
  const uint32_t j = (uint32_t)(a >> 32) + (top < alo);  // (carry; j stays < nb)
  v = (uint32_t)b;
  PXB_DASSERT(j < nb);
  const uint2 t = tab[off + j];
  PXB_DASSERT(t.y < nb);
  return top < t.x ? (int)j : (int)t.y;
}

// energy of (bin j, position v / 2^32 inside it): np.interp between the bin edges (base.py:480-482)
// -- uniform inside the bin; accept_reject returns the bin centre (base.py:484-491).
// SPEC = 1: compiled for inverse-CDF sampling on uniformly spaced linear bins.
template <int SPEC>
__device__ __forceinline__ double alias_energy(const DevModel& M, int j, uint32_t v) {
  if (SPEC == 0 && M.method == 1) return __ldg(M.emid + j);
  if (SPEC == 1 || M.edges_uniform) {
    // j 2^32 + v as an exact double: mantissa bits of 2^52, minus 2^52
    const double t = __hiloint2double(0x43300000 | j, (int)v) - 4503599627370496.0;
    return fma(t, M.dedge * 2.3283064365386963e-10, M.edge0);
  }
  const double e0 = __ldg(M.edges + j), e1 = __ldg(M.edges + j + 1);
  return e0 + ((double)v * 2.3283064365386963e-10) * (e1 - e0);
}

// global-memory alias row of (table, row).  (32-bit element index: the whole table set has fewer
// than 2^31 entries, checked when the model is uploaded -- one multiply-add instead of 64-bit
// address arithmetic per photon)
__device__ __forceinline__ unsigned alias_row_index(const DevTable* t, int tab, int row) {
  return (unsigned)(tab * t->nrows + row) * (unsigned)t->alias_stride;
}
__device__ __forceinline__ const uint2* alias_row(const DevTable* t, int tab, int row) {
  return t->alias + alias_row_index(t, tab, row);
}

// probability mode (no split): the photon picks its component with u1, then samples that row
// with (w0, w1).  Component order and weights as in fast_setup.
// th != nullptr: the cell's cumulative component probabilities (fast_setup, meta bit 256).
template <int SPEC>
__device__ __forceinline__ double fast_draw(const DevModel& M, const pxb_thermal_cells& C,
                                            const FastCell& fc, int64_t i, const double* th,
                                            double u1, uint32_t w0, uint32_t w1) {
  const DevTable* t = (fc.meta & 8) ? &M.fb : &M.prim;
  const int nr = fc.meta & 7;
  if (th) {
    const int ntab = 2 + C.nvar;
    int lo = 0, hi = nr * ntab - 1;          // first component with u1 < th[c]
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (u1 >= __ldg(th + mid)) lo = mid + 1; else hi = mid;
    }
    const int r = lo / ntab;
    PXB_DASSERT(lo >= 0 && lo < nr * ntab && u1 < __ldg(th + lo));
    uint32_t v;
    const int j = alias_pick(t->alias, alias_row_index(t, lo - r * ntab, fast_row(t, fc.z1, nr, r)), (unsigned)t->nbins, w0, w1, v);
    return alias_energy<SPEC>(M, j, v);
  }
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
  uint32_t v;
  const int j = alias_pick(t->alias, alias_row_index(t, tab, row), (unsigned)t->nbins, w0, w1, v);
  return alias_energy<SPEC>(M, j, v);
}
