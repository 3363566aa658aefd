// spectra.cu -- spectrum / band modes ("next" row f1): the reference's remaining native
// routines (pyxsim/lib/spectra.pyx: shift_spectrum :67-93, make_band :96-126,
// power_law_spectrum :11-32, line_spectrum :35-64), the thermal ones fused with the table
// lookup + abundance combine + clip of ThermalSourceModel.process_data (base.py:453-460).
#include <cstdlib>
#include "thermal.cuh"

// total spectrum value of bin j for a cell (clipped at 0): base.py:456-460
__device__ __forceinline__ double spec_bin(const CellMix& mx, double Z, const double* eZ, int nvar,
                                           int j) {
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

// np.interp(x, xp, fp) with xp ascending in global memory and fp in shared memory.  When the
// nodes are uniformly spaced (linear energy binning: uni != 0, xp[j] ~ x0 + j*dx) the interval
// comes from one division and a fix-up against the stored nodes instead of a binary search --
// the same interval either way.
__device__ __forceinline__ double interp_cum(double x, const double* __restrict__ xp,
                                             const double* fp, int n, int uni = 0, double x0n = 0.0,
                                             double inv_dx = 0.0) {
  if (x <= __ldg(xp)) return fp[0];
  if (x >= __ldg(xp + n - 1)) return fp[n - 1];
  int lo;
  if (uni) {
    lo = (int)((x - x0n) * inv_dx);
    lo = min(max(lo, 0), n - 2);
    while (lo > 0 && __ldg(xp + lo) > x) --lo;
    while (lo < n - 2 && __ldg(xp + lo + 1) <= x) ++lo;
  } else {
    lo = 0;
    int hi = n - 1;  // xp[lo] <= x < xp[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(xp + mid) <= x) lo = mid; else hi = mid;
    }
  }
  const double x0 = __ldg(xp + lo), x1 = __ldg(xp + lo + 1);
  const double slope = (fp[lo + 1] - fp[lo]) / (x1 - x0);
  return slope * (x - x0) + fp[lo];
}

constexpr int SPEC_THREADS = 256;

// persistent CTAs; each keeps a private accumulator of the regridded spectrum in shared memory
__global__ void __launch_bounds__(SPEC_THREADS)
k_thermal_spectrum(const DevModel M, const __grid_constant__ pxb_thermal_cells C,
                   const double* __restrict__ shift, int nnew, const double* __restrict__ ebnew,
                   double* __restrict__ spec_out, const int64_t* __restrict__ list,
                   const unsigned long long* __restrict__ list_count) {
  extern __shared__ double smem[];
  const int nbins = M.nbins;
  double* cum = smem;                      // nbins + 1
  double* acc = cum + (nbins + 1);         // nnew
  double* s_eZ = acc + nnew;               // PXB_MAX_VAR
  double* s_wtot = s_eZ + PXB_MAX_VAR;     // warps
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int NW = SPEC_THREADS / 32;
  const int seg = ((nbins + NW - 1) / NW + 31) & ~31;
  for (int j = threadIdx.x; j < nnew; j += blockDim.x) acc[j] = 0.0;
  // linear binning: M.edges are the keV edges themselves and uniformly spaced
  const int uni = (M.edges_uniform && !M.binscale_log) ? 1 : 0;
  const double inv_de = uni ? 1.0 / M.dedge : 0.0;
  const int64_t nwork = list ? (int64_t)*list_count : C.ncell;   // a list: only the cells on it
  for (int64_t it = blockIdx.x; it < nwork; it += gridDim.x) {
    const int64_t i = list ? list[it] : it;
    const double kT = C.kT[i];
    if (!cell_cut(C, i, kT)) continue;                  // uniform across the CTA
    const double nH = C.nH ? C.nH[i] : 0.0;
    CellMix mx;
    cell_mix(M, kT, nH, mx);
    const double xh = cell_xh(C, i);
    const double Z = cell_metal(C, i, xh);
    __syncthreads();
    for (int k = threadIdx.x; k < C.nvar; k += blockDim.x) s_eZ[k] = cell_elem(C, k, i, xh);
    if (threadIdx.x == 0) cum[0] = 0.0;
    __syncthreads();
    const int j0 = wid * seg, j1 = min(j0 + seg, nbins);
    double carry = 0.0;
    for (int jb = j0; jb < j1; jb += 32) {
      const int j = jb + lane;
      double v = (j < j1) ? spec_bin(mx, Z, s_eZ, C.nvar, j) : 0.0;
      v = warp_incl_scan(v, lane) + carry;
      if (j < j1) cum[j + 1] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
    if (lane == 0) s_wtot[wid] = carry;
    __syncthreads();
    double basev = 0.0;
    for (int w = 0; w < wid; ++w) basev += s_wtot[w];
    if (wid > 0)
      for (int j = j0 + lane; j < j1; j += 32) cum[j + 1] += basev;
    __syncthreads();
    const double s = shift ? shift[i] : 1.0;
    const double wgt = s * s * s * cell_norm(C, i);
    const double inv_s = 1.0 / s;
    for (int j = threadIdx.x; j < nnew; j += blockDim.x) {
      // ebnew / shift as in the reference (a division, not a multiplication by 1/shift)
      const double a = interp_cum(__ldg(ebnew + j) / s, M.ebins, cum, nbins + 1, uni, M.edge0, inv_de);
      const double b = interp_cum(__ldg(ebnew + j + 1) / s, M.ebins, cum, nbins + 1, uni, M.edge0, inv_de);
      acc[j] += wgt * (b - a);
    }
    (void)inv_s;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nnew; j += blockDim.x)
    if (acc[j] != 0.0) atomicAdd(spec_out + j, acc[j]);
}

// ---- unshifted spectra by linearity ------------------------------------------------------
// Without a Doppler shift the regridding of shift_spectrum is ONE linear map applied to every
// cell's spectrum, and for a cell whose clip (base.py:460) is a no-op -- non-negative tables,
// abundances and interpolation weights -- that spectrum is a weighted sum of table rows.  So
//   sum_cells cnm * regrid(tot_spec) = regrid( sum_rows W[table][row] * table[row] ),
//   W[table][row] = sum_cells cnm * w_row * coefficient_table,
// O(1) work per cell instead of O(nbins).  Cells that do need the clip go, through a list, to
// the per-cell kernel above.
__global__ void __launch_bounds__(256)
k_spectrum_weights(const DevModel M, const __grid_constant__ pxb_thermal_cells C,
                   double* __restrict__ W, int64_t* __restrict__ list,
                   unsigned long long* __restrict__ list_count, int use_smem) {
  // W layout: primary table [(2+nvar)][nrows_prim], then fallback table [(2+nvar)][nrows_fb]
  extern __shared__ double s_w[];
  const int ntab = 2 + C.nvar;
  const int np = M.prim.nrows, nf = M.has_fb ? M.fb.nrows : 0;
  const int nw = ntab * (np + nf);
  if (use_smem) {
    for (int j = threadIdx.x; j < nw; j += blockDim.x) s_w[j] = 0.0;
    __syncthreads();
  }
  double* acc = use_smem ? s_w : W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < C.ncell; i += stride) {
    const double kT = C.kT[i];
    if (!cell_cut(C, i, kT)) continue;
    const double nH = C.nH ? C.nH[i] : 0.0;
    CellMix mx;
    cell_mix(M, kT, nH, mx);
    const double xh = cell_xh(C, i);
    const double Z = cell_metal(C, i, xh);
    bool lin = mx.t->nonneg && mx.interior && (Z >= 0.0);
    for (int k = 0; k < C.nvar; ++k) lin = lin && (cell_elem(C, k, i, xh) >= 0.0);
    if (!lin) {
      list[atomicAdd(list_count, 1ull)] = i;
      continue;
    }
    const double cnm = cell_norm(C, i);
    const bool prim = mx.t == &M.prim;
    double* base = acc + (prim ? 0 : ntab * np);
    const int nr_t = prim ? np : nf;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r >= mx.nr) continue;
      const double wr = cnm * mx.w[r];
      if (wr == 0.0) continue;
      atomicAdd(base + mx.row[r], wr);
      atomicAdd(base + nr_t + mx.row[r], wr * Z);
      for (int k = 0; k < C.nvar; ++k)
        atomicAdd(base + (2 + k) * nr_t + mx.row[r], wr * cell_elem(C, k, i, xh));
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int j = threadIdx.x; j < nw; j += blockDim.x)
      if (s_w[j] != 0.0) atomicAdd(W + j, s_w[j]);
  }
}

// one CTA: summed spectrum on the model's bins from the row weights, cumulative, np.interp at
// the new edges (shift_spectrum with shift = 1, spectra.pyx:67-93), accumulate into spec_out
__global__ void __launch_bounds__(1024)
k_spectrum_from_weights(const DevModel M, int nvar, const double* __restrict__ W, int nnew,
                        const double* __restrict__ ebnew, double* __restrict__ spec_out) {
  extern __shared__ double smem[];
  const int nbins = M.nbins;
  double* cum = smem;                 // nbins + 1
  double* s_wtot = cum + (nbins + 1);  // warps
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int NW = blockDim.x / 32;
  const int seg = ((nbins + NW - 1) / NW + 31) & ~31;
  const int ntab = 2 + nvar;
  if (threadIdx.x == 0) cum[0] = 0.0;
  const int j0 = wid * seg, j1 = min(j0 + seg, nbins);
  double carry = 0.0;
  for (int jb = j0; jb < j1; jb += 32) {
    const int j = jb + lane;
    double v = 0.0;
    if (j < j1) {
      for (int tb = 0; tb < 2; ++tb) {           // primary, fallback
        const DevTable& t = tb == 0 ? M.prim : M.fb;
        if (tb == 1 && !M.has_fb) break;
        const double* Wt = W + (tb == 0 ? 0 : (size_t)ntab * M.prim.nrows);
        for (int q = 0; q < ntab; ++q) {
          const double* tab = q == 0 ? t.cosmic : (q == 1 ? t.metal : t.var + (size_t)(q - 2) * t.nrows * nbins);
          for (int r = 0; r < t.nrows; ++r) {
            const double w = __ldg(Wt + (size_t)q * t.nrows + r);
            if (w != 0.0) v += w * __ldg(tab + (size_t)r * nbins + j);
          }
        }
      }
    }
    v = warp_incl_scan(v, lane) + carry;
    if (j < j1) cum[j + 1] = v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 0) s_wtot[wid] = carry;
  __syncthreads();
  double basev = 0.0;
  for (int w = 0; w < wid; ++w) basev += s_wtot[w];
  if (wid > 0)
    for (int j = j0 + lane; j < j1; j += 32) cum[j + 1] += basev;
  __syncthreads();
  for (int j = threadIdx.x; j < nnew; j += blockDim.x) {
    const double a = interp_cum(__ldg(ebnew + j), M.ebins, cum, nbins + 1);
    const double b = interp_cum(__ldg(ebnew + j + 1), M.ebins, cum, nbins + 1);
    const double d = b - a;
    if (d != 0.0) atomicAdd(spec_out + j, d);
  }
}

// ---- Doppler-shifted spectra of cells whose clip is a no-op: bucketed by table row ----------
// For such a cell (non-negative 1-D table, Z >= 0, interior temperature) the cumulative
// spectrum that shift_spectrum interpolates (spectra.pyx:80-90) is a blend of four per-row
// prefix sums,
//   cum = w0 (Pc[r0] + Z Pm[r0]) + w1 (Pc[r0+1] + Z Pm[r0+1]),
// and np.interp is linear in its ordinates, so the cell's contribution at a new edge e is the
// same blend of the rows' interpolants at e / shift.  Cells are counting-sorted by r0; a CTA
// stages the four rows of one bucket in shared memory once per 1024 cells, each LANE owns one
// cell (its five constants in registers) and walks the new edges; the 32 cells x 32 bins of a
// warp are summed by a transposing butterfly (lane L ends with bin L's total) and added to the
// CTA's accumulator.  Linear model bins only (the interval is one multiply-add); everything
// else goes through the list to the per-cell kernel above.
constexpr int SB_CHUNK = 1024;      // cells per work item
constexpr int SB_THREADS = 512;

struct SpecBuckets {
  int* key;                  // [ncell] bucket of the cell, -1: not on this path
  double* par;               // [5][ncell] sorted cells: a0..a3 (blend x weight), k1 = 1/(shift*de)
  unsigned* hist;            // [nb] cells per bucket
  unsigned* start;           // [nb] first sorted slot of the bucket
  unsigned* cursor;          // [nb]
  int2* items;               // work items: (bucket, first sorted slot); count in n_items
  unsigned* n_items;
  unsigned* ticket;
  int nb;
};

__device__ __forceinline__ bool spec_bucket_cell(const DevModel& M, const pxb_thermal_cells& C,
                                                 int64_t i, CellMix& mx, double& Z, bool& listed) {
  listed = false;
  const double kT = C.kT[i];
  if (!cell_cut(C, i, kT)) return false;
  const double nH = C.nH ? C.nH[i] : 0.0;
  cell_mix(M, kT, nH, mx);
  const double xh = cell_xh(C, i);
  Z = cell_metal(C, i, xh);
  const bool ok = mx.t->nonneg && mx.t->pre_n && mx.interior && (Z >= 0.0) && mx.nr == 2;
  listed = !ok;
  return ok;
}

__global__ void __launch_bounds__(256)
k_spec_bucket_count(const DevModel M, const __grid_constant__ pxb_thermal_cells C, SpecBuckets B,
                    int64_t* __restrict__ list, unsigned long long* __restrict__ list_count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < C.ncell; i0 += stride) {
    const int64_t i = i0 + threadIdx.x;
    int key = -1;
    if (i < C.ncell) {
      CellMix mx;
      double Z;
      bool listed;
      if (spec_bucket_cell(M, C, i, mx, Z, listed))
        key = (mx.t == &M.prim ? 0 : M.prim.nrows) + mx.row[0];
      else if (listed)
        list[atomicAdd(list_count, 1ull)] = i;
      B.key[i] = key;
    }
    // one atomic per distinct bucket of the warp
    const unsigned act = __ballot_sync(0xffffffffu, key >= 0);
    if (key >= 0) {
      const unsigned same = __match_any_sync(act, key);
      if ((int)(threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(B.hist + key, (unsigned)__popc(same));
    }
  }
}

// one CTA: bucket starts and the work items (nb is a few hundred at most)
__global__ void k_spec_bucket_plan(SpecBuckets B) {
  if (threadIdx.x != 0) return;
  unsigned at = 0, ni = 0;
  for (int b = 0; b < B.nb; ++b) {
    const unsigned n = B.hist[b];
    B.start[b] = at;
    B.cursor[b] = at;
    for (unsigned o = 0; o < n; o += SB_CHUNK) B.items[ni++] = make_int2(b, (int)(at + o));
    at += n;
  }
  *B.n_items = ni;
}

__global__ void __launch_bounds__(256)
k_spec_bucket_fill(const DevModel M, const __grid_constant__ pxb_thermal_cells C, SpecBuckets B,
                   const double* __restrict__ shift) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const double inv_de = 1.0 / M.dedge;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < C.ncell; i0 += stride) {
    const int64_t i = i0 + threadIdx.x;
    const int key = i < C.ncell ? B.key[i] : -1;
    const unsigned act = __ballot_sync(0xffffffffu, key >= 0);
    if (key < 0) continue;
    const unsigned same = __match_any_sync(act, key);
    const int leader = __ffs(same) - 1;
    unsigned pos = 0;
    if (lane == leader) pos = atomicAdd(B.cursor + key, (unsigned)__popc(same));
    pos = __shfl_sync(same, pos, leader) + (unsigned)__popc(same & ((1u << lane) - 1u));
    CellMix mx;
    double Z;
    bool listed;
    spec_bucket_cell(M, C, i, mx, Z, listed);
    const double s = shift[i];
    const double wgt = s * s * s * cell_norm(C, i);          // spectra.pyx:91 with the cell's norm
    const int64_t n = C.ncell;
    B.par[pos] = wgt * mx.w[0];
    B.par[n + pos] = wgt * mx.w[0] * Z;
    B.par[2 * n + pos] = wgt * mx.w[1];
    B.par[3 * n + pos] = wgt * mx.w[1] * Z;
    B.par[4 * n + pos] = inv_de / s;
  }
}

// value of the blended cumulative spectrum at model-bin coordinate u (bins from the first edge).
// ZCONST (one metallicity for the whole call): the staged rows are already Pc + Z Pm.
template <bool CLAMP, bool ZCONST>
__device__ __forceinline__ double sb_eval(const double2* __restrict__ s_c, const double2* __restrict__ s_m,
                                          double u, int nbins, double a0, double a1, double a2,
                                          double a3) {
  if (CLAMP) u = fmin(fmax(u, 0.0), (double)nbins);          // np.interp holds the end values
  int lo = __double2int_rd(u);
  if (CLAMP) lo = min(lo, nbins - 1);
  const double t = u - (double)lo;
  PXB_DASSERT(lo >= 0 && lo + 1 <= nbins);
  const double2 c0 = s_c[lo], c1 = s_c[lo + 1];
  double f0, f1;
  if (ZCONST) {
    f0 = fma(a2, c0.y, a0 * c0.x);
    f1 = fma(a2, c1.y, a0 * c1.x);
  } else {
    const double2 m0 = s_m[lo], m1 = s_m[lo + 1];
    f0 = fma(a3, m0.y, fma(a2, c0.y, fma(a1, m0.x, a0 * c0.x)));
    f1 = fma(a3, m1.y, fma(a2, c1.y, fma(a1, m1.x, a0 * c1.x)));
  }
  return fma(t, f1 - f0, f0);
}

// CPL cells per lane: the butterfly, the edge loads and the loop are shared by 32 CPL cells
template <bool ZCONST, int CPL>
__global__ void __launch_bounds__(SB_THREADS, 1)
k_spec_bucket_run(const DevModel M, SpecBuckets B, int64_t ncell, double zconst, int nnew,
                  const double* __restrict__ ebnew, double* __restrict__ spec_out) {
  extern __shared__ __align__(16) unsigned char sb_dyn[];
  const int nbins = M.nbins;
  double2* s_c = reinterpret_cast<double2*>(sb_dyn);           // [nbins+1] {Pc[r0], Pc[r0+1]}
  double2* s_m = s_c + (nbins + 1);                             // [nbins+1] {Pm[r0], Pm[r0+1]} (!ZCONST)
  double* s_acc = reinterpret_cast<double*>(ZCONST ? s_m : s_m + (nbins + 1));   // [ngrp*32]
  const int ngrp = (nnew + 31) / 32;
  double* s_eb = s_acc + ngrp * 32;                             // [ngrp*32 + 1], padded with the last edge
  __shared__ int s_item;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int j = threadIdx.x; j < ngrp * 32; j += blockDim.x) s_acc[j] = 0.0;
  for (int j = threadIdx.x; j <= ngrp * 32; j += blockDim.x) s_eb[j] = __ldg(ebnew + min(j, nnew));
  const double c0 = -M.edge0 / M.dedge;
  const double e_first = __ldg(ebnew), e_last = __ldg(ebnew + nnew);
  const unsigned n_items = *B.n_items;
  int cur = -1;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = (int)atomicAdd(B.ticket, 1u);
    __syncthreads();
    const unsigned it = (unsigned)s_item;
    if (it >= n_items) break;
    const int2 item = B.items[it];
    if (item.x != cur) {
      cur = item.x;
      const bool prim = cur < M.prim.nrows;
      const DevTable& t = prim ? M.prim : M.fb;
      const int r0 = prim ? cur : cur - M.prim.nrows;
      const size_t rs = (size_t)(nbins + 1);
      const double* pc = t.pre_n + (size_t)r0 * rs;
      const double* pm = t.pre_n + ((size_t)t.nrows + r0) * rs;
      for (int j = threadIdx.x; j <= nbins; j += blockDim.x) {
        if (ZCONST) {
          s_c[j] = make_double2(fma(zconst, __ldg(pm + j), __ldg(pc + j)),
                                fma(zconst, __ldg(pm + rs + j), __ldg(pc + rs + j)));
        } else {
          s_c[j] = make_double2(__ldg(pc + j), __ldg(pc + rs + j));
          s_m[j] = make_double2(__ldg(pm + j), __ldg(pm + rs + j));
        }
      }
      __syncthreads();
    }
    const unsigned bend = B.start[cur] + B.hist[cur];
    const unsigned iend = min((unsigned)item.y + SB_CHUNK, bend);
    for (unsigned base = (unsigned)item.y + wid * (32 * CPL); base < iend; base += SB_THREADS * CPL) {
      double a0[CPL], a1[CPL], a2[CPL], a3[CPL], k1[CPL], vprev[CPL];
      double kmin = 1.0e300, kmax = 0.0;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const unsigned p = base + c * 32 + lane;
        const bool valid = p < iend;
        a0[c] = valid ? B.par[p] : 0.0;
        a2[c] = valid ? B.par[2 * ncell + p] : 0.0;
        a1[c] = (valid && !ZCONST) ? B.par[ncell + p] : 0.0;
        a3[c] = (valid && !ZCONST) ? B.par[3 * ncell + p] : 0.0;
        k1[c] = B.par[4 * ncell + (valid ? p : base)];
        kmin = fmin(kmin, k1[c]);
        kmax = fmax(kmax, k1[c]);
      }
      // do all the warp's cells keep every new edge inside the model's range?
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        kmin = fmin(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = fmax(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
      }
      const bool inside = fma(e_first, kmin, c0) >= 0.0 && fma(e_first, kmax, c0) >= 0.0 &&
                          fma(e_last, kmax, c0) < (double)nbins && fma(e_last, kmin, c0) < (double)nbins;
#pragma unroll
      for (int c = 0; c < CPL; ++c)
        vprev[c] = sb_eval<true, ZCONST>(s_c, s_m, fma(s_eb[0], k1[c], c0), nbins, a0[c], a1[c], a2[c], a3[c]);
#pragma unroll 1
      for (int g = 0; g < ngrp; ++g) {
        double d[32];
        const double* eb = s_eb + g * 32 + 1;
        if (inside) {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const double e = eb[q];
            double sum = 0.0;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
              const double v = sb_eval<false, ZCONST>(s_c, s_m, fma(e, k1[c], c0), nbins, a0[c], a1[c], a2[c], a3[c]);
              sum += v - vprev[c];
              vprev[c] = v;
            }
            d[q] = sum;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const double e = eb[q];
            double sum = 0.0;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
              const double v = sb_eval<true, ZCONST>(s_c, s_m, fma(e, k1[c], c0), nbins, a0[c], a1[c], a2[c], a3[c]);
              sum += v - vprev[c];
              vprev[c] = v;
            }
            d[q] = sum;
          }
        }
        // transposing butterfly: after the five steps d[0] of lane L is bin (32 g + L) summed
        // over the warp's cells
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const bool up = (lane & o) != 0;
#pragma unroll
          for (int q = 0; q < o; ++q) {
            const double keep = up ? d[q + o] : d[q];
            const double send = up ? d[q] : d[q + o];
            d[q] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
        if (d[0] != 0.0) atomicAdd(s_acc + g * 32 + lane, d[0]);
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nnew; j += blockDim.x)
    if (s_acc[j] != 0.0) atomicAdd(spec_out + j, s_acc[j]);
}

// make_band in O(1) per cell.  For a cell whose clip is a no-op the band sum of its blended
// spectrum is the blend of the rows' band sums, and a band sum over whole bins is a difference of
// two prefix sums:  sum_{j in band} ener_j spec_ij = sum_rows w_r (P_r[j1] - P_r[j0]).  The band's
// bins are those with ebins[j] >= emin/shift and ebins[j+1] <= emax/shift (spectra.pyx:115-118):
// j0 = first j with ebins[j] >= lo, j1 = last q with ebins[q] <= hi.  Other cells are listed for
// the warp-per-cell kernel.
__global__ void __launch_bounds__(256)
k_thermal_band_fast(const DevModel M, const __grid_constant__ pxb_thermal_cells C, int use_energy,
                    double emin, double emax, const double* __restrict__ shift,
                    double* __restrict__ out, int64_t* __restrict__ list,
                    unsigned long long* __restrict__ list_count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C.ncell) return;
  const double kT = C.kT[i];
  if (!cell_cut(C, i, kT)) {
    out[i] = 0.0;
    return;
  }
  const double nH = C.nH ? C.nH[i] : 0.0;
  CellMix mx;
  cell_mix(M, kT, nH, mx);
  const double xh = cell_xh(C, i);
  const double Z = cell_metal(C, i, xh);
  const DevTable* t = mx.t;
  bool lin = t->nonneg && t->pre_n && mx.interior && (Z >= 0.0);
  for (int k = 0; k < C.nvar; ++k) lin = lin && (cell_elem(C, k, i, xh) >= 0.0);
  if (!lin) {
    list[atomicAdd(list_count, 1ull)] = i;
    return;
  }
  const double s = shift ? shift[i] : 1.0;
  const double lo = emin / s, hi = emax / s;
  const int nb = M.nbins;
  int j0, j1;
  {
    int a = 0, b = nb + 1;                  // first j in [0, nb] with ebins[j] >= lo
    while (a < b) { const int m = (a + b) >> 1; if (__ldg(M.ebins + m) < lo) a = m + 1; else b = m; }
    j0 = a;
    a = 0; b = nb + 1;                      // number of edges <= hi
    while (a < b) { const int m = (a + b) >> 1; if (__ldg(M.ebins + m) <= hi) a = m + 1; else b = m; }
    j1 = a - 1;
  }
  double sum = 0.0;
  if (j1 > j0) {
    const double* P = use_energy ? t->pre_e : t->pre_n;
    const size_t rs = (size_t)(nb + 1), ts = (size_t)t->nrows * rs;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r >= mx.nr) continue;
      const double* pr = P + (size_t)mx.row[r] * rs;
      double v = (__ldg(pr + j1) - __ldg(pr + j0)) + Z * (__ldg(pr + ts + j1) - __ldg(pr + ts + j0));
      for (int k = 0; k < C.nvar; ++k)
        v += cell_elem(C, k, i, xh) * (__ldg(pr + (2 + k) * ts + j1) - __ldg(pr + (2 + k) * ts + j0));
      sum += mx.w[r] * v;
    }
  }
  double sh = s * s * s;
  if (use_energy) sh *= s;
  out[i] = sh * sum * cell_norm(C, i);
}

// make_band: one warp per cell (or per listed cell)
__global__ void __launch_bounds__(256)
k_thermal_band(const DevModel M, const __grid_constant__ pxb_thermal_cells C, int use_energy,
               double emin, double emax, const double* __restrict__ shift,
               double* __restrict__ out, const int64_t* __restrict__ list,
               const unsigned long long* __restrict__ list_count) {
  __shared__ double s_eZ[8][PXB_MAX_VAR];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (it >= (list ? (int64_t)*list_count : C.ncell)) return;
  const int64_t i = list ? list[it] : it;
  const double kT = C.kT[i];
  if (!cell_cut(C, i, kT)) {
    if (lane == 0) out[i] = 0.0;
    return;
  }
  const double nH = C.nH ? C.nH[i] : 0.0;
  CellMix mx;
  cell_mix(M, kT, nH, mx);
  const double xh = cell_xh(C, i);
  const double Z = cell_metal(C, i, xh);
  for (int k = lane; k < C.nvar; k += 32) s_eZ[wib][k] = cell_elem(C, k, i, xh);
  __syncwarp();
  const double s = shift ? shift[i] : 1.0;
  const double lo = emin / s, hi = emax / s;
  double sum = 0.0;
  for (int j = lane; j < M.nbins; j += 32) {
    if (__ldg(M.ebins + j) >= lo && __ldg(M.ebins + j + 1) <= hi) {
      const double ener = use_energy ? __ldg(M.emid + j) : 1.0;
      sum += ener * spec_bin(mx, Z, s_eZ[wib], C.nvar, j);
    }
  }
  sum = warp_sum(sum);
  if (lane == 0) {
    double sh = s * s * s;
    if (use_energy) sh *= s;
    out[i] = sh * sum * cell_norm(C, i);
  }
}

// power_law_spectrum / line_spectrum: CTA = 256 bins x a slab of cells, register accumulators
constexpr int SLAB = 4096;

__global__ void __launch_bounds__(256)
k_powerlaw_spectrum(int64_t ncell, const double* __restrict__ alpha, const double* __restrict__ K,
                    const double* __restrict__ shift, int64_t nbins,
                    const double* __restrict__ emid, double* __restrict__ spec_out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i0 = (int64_t)blockIdx.y * SLAB, i1 = min(i0 + (int64_t)SLAB, ncell);
  const double em = j < nbins ? emid[j] : 1.0;
  double acc = 0.0;
  for (int64_t i = i0; i < i1; ++i) {
    const double s = shift ? __ldg(shift + i) : 1.0;
    acc += s * s * __ldg(K + i) * pow(em / s, -__ldg(alpha + i));   // spectra.pyx:30
  }
  if (j < nbins && acc != 0.0) atomicAdd(spec_out + j, acc);
}

__global__ void __launch_bounds__(256)
k_line_spectrum(int64_t ncell, double e0, const double* __restrict__ sigma,
                const double* __restrict__ N, const double* __restrict__ shift, int64_t nbins,
                const double* __restrict__ ee, int64_t ng, const double* __restrict__ gx,
                const double* __restrict__ gpdf, double* __restrict__ spec_out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i0 = (int64_t)blockIdx.y * SLAB, i1 = min(i0 + (int64_t)SLAB, ncell);
  const double e = j < nbins ? ee[j] : 0.0;
  const double g0 = __ldg(gx), g1 = __ldg(gx + ng - 1);
  const double inv_dg = (double)(ng - 1) / (g1 - g0);
  double acc = 0.0;
  for (int64_t i = i0; i < i1; ++i) {
    const double s = shift ? __ldg(shift + i) : 1.0;
    const double sg = __ldg(sigma + i);
    const double x = (e / s - e0) / sg;                            // spectra.pyx:59
    double r;                                                       // np.interp(x, gx, gpdf)
    if (x <= g0) r = __ldg(gpdf);
    else if (x >= g1) r = __ldg(gpdf + ng - 1);
    else {
      int64_t k = (int64_t)((x - g0) * inv_dg);                    // gx is np.linspace: direct index
      k = k < 0 ? 0 : (k > ng - 2 ? ng - 2 : k);
      if (k > 0 && __ldg(gx + k) > x) --k;
      else if (k < ng - 2 && __ldg(gx + k + 1) <= x) ++k;
      const double x0 = __ldg(gx + k), x1 = __ldg(gx + k + 1);
      const double slope = (__ldg(gpdf + k + 1) - __ldg(gpdf + k)) / (x1 - x0);
      r = slope * (x - x0) + __ldg(gpdf + k);
    }
    acc += s * s * __ldg(N + i) * r / sg;                           // spectra.pyx:62
  }
  if (j < nbins && acc != 0.0) atomicAdd(spec_out + j, acc);
}

// compute_shift (sources.py:89-95): sqrt(1 - beta^2) / (1 - beta_n), beta_n = v . n / c
__global__ void __launch_bounds__(256)
k_shift_factor(int64_t n, const double* __restrict__ vx, const double* __restrict__ vy,
               const double* __restrict__ vz, double nx, double ny, double nz,
               double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double c = 299792.458;
  const double bx = vx[i] / c, by = vy[i] / c, bz = vz[i] / c;
  const double bn = bx * nx + by * ny + bz * nz;
  const double b2 = bx * bx + by * by + bz * bz;
  out[i] = sqrt(1.0 - b2) / (1.0 - bn);
}

extern "C" int pxb_shift_factor(int64_t n, const double* vx, const double* vy, const double* vz,
                                const double* normal, double* out, void* stream) {
  PXB_REQUIRE(n >= 0 && normal, "bad argument");
  if (n == 0) return PXB_OK;
  PXB_REQUIRE(vx && vy && vz && out, "NULL argument");
  const double nn = sqrt(normal[0] * normal[0] + normal[1] * normal[1] + normal[2] * normal[2]);
  PXB_REQUIRE(nn > 0.0, "zero normal vector");
  k_shift_factor<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      n, vx, vy, vz, normal[0] / nn, normal[1] / nn, normal[2] / nn, out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

// K = L / K_fac of the power-law model (power_law_sources.py:196-206)
__global__ void __launch_bounds__(256)
k_powerlaw_K(const __grid_constant__ pxb_powerlaw_cells C, double* __restrict__ K_out,
             double* __restrict__ alpha_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C.ncell) return;
  const double alpha = C.alpha ? C.alpha[i] : C.alpha_const;
  const double etoalpha = pow(C.e0, alpha);
  double kfac;
  if (alpha == 2.0) kfac = log(C.emax / C.emin) * etoalpha;
  else kfac = (pow(C.emax, 2.0 - alpha) - pow(C.emin, 2.0 - alpha)) * etoalpha / (2.0 - alpha);
  K_out[i] = C.lum[i] / kfac;
  alpha_out[i] = alpha;
}

extern "C" int pxb_powerlaw_norm(const pxb_powerlaw_cells* cells, double* K_out, double* alpha_out,
                              void* stream) {
  PXB_REQUIRE(cells && cells->ncell >= 0, "bad argument");
  if (cells->ncell == 0) return PXB_OK;
  PXB_REQUIRE(cells->lum && K_out && alpha_out, "NULL argument");
  k_powerlaw_K<<<(unsigned)((cells->ncell + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*cells, K_out, alpha_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

static int check_model_cells(const pxb_model* model, const pxb_thermal_cells* c) {
  PXB_REQUIRE(model && c, "NULL model/cells");
  PXB_REQUIRE(c->nvar == model->dev.nvar, "cells.nvar=%d does not match the table's nvar=%d",
              c->nvar, model->dev.nvar);
  PXB_REQUIRE(c->ncell == 0 || (c->kT && c->em), "kT/em arrays are required");
  PXB_REQUIRE(!model->dev.density_dependent || c->ncell == 0 || c->nH, "density dependent model needs nH");
  return PXB_OK;
}

// workspace of the bucketed shifted path: [keys | par | hist,start,cursor | n_items,ticket | items]
static int spec_bucket_count(const DevModel& D) { return D.prim.nrows + (D.has_fb ? D.fb.nrows : 0); }
static int64_t spec_bucket_bytes(const DevModel& D, int64_t ncell) {
  const int64_t nb = spec_bucket_count(D);
  const int64_t items = ncell / SB_CHUNK + nb + 1;
  return ((ncell + 1) / 2 + 5 * ncell + (3 * nb + 2 + 1) / 2 + 1 + items) * (int64_t)sizeof(double);
}

static int64_t spectrum_weight_count(const DevModel& D) {
  return (int64_t)(2 + D.nvar) * (D.prim.nrows + (D.has_fb ? D.fb.nrows : 0));
}

extern "C" int pxb_thermal_spectrum_workspace_bytes(const pxb_model* model, int64_t ncell,
                                                    int64_t* bytes) {
  PXB_REQUIRE(model && bytes && ncell >= 0, "bad argument");
  // [list count, pad | row weights | list of cells that need the per-cell path]
  *bytes = (2 + spectrum_weight_count(model->dev) + ncell) * (int64_t)sizeof(double) +
           spec_bucket_bytes(model->dev, ncell);
  return PXB_OK;
}

extern "C" int pxb_thermal_spectrum(const pxb_model* model, const pxb_thermal_cells* cells,
                                    const double* shift, int64_t nnew, const double* ebnew,
                                    double* spec_out, void* workspace, int64_t workspace_bytes,
                                    void* stream) {
  int rc = check_model_cells(model, cells);
  if (rc) return rc;
  PXB_REQUIRE(nnew >= 1 && ebnew && spec_out, "bad output grid");
  if (cells->ncell == 0) return PXB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const DevModel& D = model->dev;
  const size_t smem = ((size_t)D.nbins + 1 + nnew + PXB_MAX_VAR + SPEC_THREADS / 32) * sizeof(double);
  PXB_REQUIRE(smem <= 227 * 1024, "nbins + output bins = %lld needs %zu B of shared memory (max 227 KB)",
              (long long)(D.nbins + nnew), smem);
  PXB_CUDA(cudaFuncSetAttribute(k_thermal_spectrum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PXB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_thermal_spectrum, SPEC_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)pxb_sm_count() * per_sm;
  if (grid > cells->ncell) grid = cells->ncell;
  // Without a shift and with a workspace the spectrum is assembled from per-row weights; only
  // the cells whose clip matters take the per-cell path.
  const bool linear = shift == nullptr && workspace != nullptr &&
                      (D.prim.nonneg || (D.has_fb && D.fb.nonneg));
  if (linear) {
    int64_t need = 0;
    pxb_thermal_spectrum_workspace_bytes(model, cells->ncell, &need);
    PXB_REQUIRE(workspace_bytes >= need, "spectrum workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)need);
    const int64_t nw = spectrum_weight_count(D);
    unsigned long long* count = (unsigned long long*)workspace;
    double* W = (double*)workspace + 2;
    int64_t* list = (int64_t*)(W + nw);
    PXB_CUDA(cudaMemsetAsync(workspace, 0, (size_t)(2 + nw) * sizeof(double), st));
    const size_t wsm = (size_t)nw * sizeof(double);
    const int use_smem = wsm <= 96 * 1024 ? 1 : 0;
    if (use_smem)
      PXB_CUDA(cudaFuncSetAttribute(k_spectrum_weights, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm));
    int64_t g = (cells->ncell + 255) / 256;
    const int64_t cap = (int64_t)pxb_sm_count() * 4;
    if (g > cap) g = cap;
    k_spectrum_weights<<<(unsigned)g, 256, use_smem ? wsm : 0, st>>>(D, *cells, W, list, count, use_smem);
    PXB_LAUNCH_CHECK();
    const size_t fsm = ((size_t)D.nbins + 1 + 32) * sizeof(double);
    PXB_CUDA(cudaFuncSetAttribute(k_spectrum_from_weights, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
    k_spectrum_from_weights<<<1, 1024, fsm, st>>>(D, cells->nvar, W, (int)nnew, ebnew, spec_out);
    PXB_LAUNCH_CHECK();
    k_thermal_spectrum<<<(unsigned)grid, SPEC_THREADS, smem, st>>>(D, *cells, shift, (int)nnew, ebnew,
                                                                  spec_out, list, count);
    PXB_LAUNCH_CHECK();
    return PXB_OK;
  }
  // Doppler-shifted, linear model bins, no variable elements: cells bucketed by table row
  const int ngrp = (int)((nnew + 31) / 32);
  const size_t bsm = ((size_t)(D.nbins + 1) * 4 + (size_t)ngrp * 64 + 2) * sizeof(double);
  const bool bucketed = shift != nullptr && workspace != nullptr && D.nvar == 0 &&
                        D.edges_uniform && !D.binscale_log && bsm <= 227 * 1024 &&
                        cells->ncell < (1LL << 31) && (D.prim.pre_n || (D.has_fb && D.fb.pre_n));
  if (bucketed) {
    int64_t need = 0;
    pxb_thermal_spectrum_workspace_bytes(model, cells->ncell, &need);
    PXB_REQUIRE(workspace_bytes >= need, "spectrum workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)need);
    const int64_t n = cells->ncell, nw = spectrum_weight_count(D);
    unsigned long long* count = (unsigned long long*)workspace;
    int64_t* list = (int64_t*)((double*)workspace + 2 + nw);
    SpecBuckets B;
    B.nb = spec_bucket_count(D);
    double* p = (double*)(list + n);
    B.key = (int*)p;                p += (n + 1) / 2;
    B.par = p;                      p += 5 * n;
    B.hist = (unsigned*)p;
    B.start = B.hist + B.nb;
    B.cursor = B.start + B.nb;
    B.n_items = B.cursor + B.nb;
    B.ticket = B.n_items + 1;       p += (3 * B.nb + 2 + 1) / 2 + 1;
    B.items = (int2*)p;
    PXB_CUDA(cudaMemsetAsync(workspace, 0, 2 * sizeof(double), st));
    PXB_CUDA(cudaMemsetAsync(B.hist, 0, (size_t)(3 * B.nb + 2) * sizeof(unsigned), st));
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)pxb_sm_count() * 8;
    if (g > cap) g = cap;
    k_spec_bucket_count<<<(unsigned)g, 256, 0, st>>>(D, *cells, B, list, count);
    PXB_LAUNCH_CHECK();
    k_spec_bucket_plan<<<1, 32, 0, st>>>(B);
    PXB_LAUNCH_CHECK();
    k_spec_bucket_fill<<<(unsigned)g, 256, 0, st>>>(D, *cells, B, shift);
    PXB_LAUNCH_CHECK();
    int64_t gr = n / SB_CHUNK + B.nb;
    if (gr > pxb_sm_count()) gr = pxb_sm_count();
    // one metallicity for the call: the metal rows are folded into the staged rows
    if (!cells->zmet) {
      auto kern = k_spec_bucket_run<true, 2>;
      PXB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
      kern<<<(unsigned)gr, SB_THREADS, bsm, st>>>(D, B, n, cells->zmet_const, (int)nnew, ebnew, spec_out);
    } else {
      auto kern = k_spec_bucket_run<false, 1>;
      PXB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
      kern<<<(unsigned)gr, SB_THREADS, bsm, st>>>(D, B, n, 0.0, (int)nnew, ebnew, spec_out);
    }
    PXB_LAUNCH_CHECK();
    // the cells whose clip matters (or with four rows): per-cell kernel over the list
    k_thermal_spectrum<<<(unsigned)grid, SPEC_THREADS, smem, st>>>(D, *cells, shift, (int)nnew, ebnew,
                                                                  spec_out, list, count);
    PXB_LAUNCH_CHECK();
    return PXB_OK;
  }
  k_thermal_spectrum<<<(unsigned)grid, SPEC_THREADS, smem, st>>>(D, *cells, shift, (int)nnew, ebnew,
                                                                spec_out, nullptr, nullptr);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_thermal_band_workspace_bytes(int64_t ncell, int64_t* bytes) {
  PXB_REQUIRE(bytes && ncell >= 0, "bad argument");
  *bytes = (2 + ncell) * (int64_t)sizeof(int64_t);   // [list count, pad | list]
  return PXB_OK;
}

extern "C" int pxb_thermal_band(const pxb_model* model, const pxb_thermal_cells* cells, int use_energy,
                                double emin, double emax, const double* shift, double* out,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = check_model_cells(model, cells);
  if (rc) return rc;
  if (cells->ncell == 0) return PXB_OK;
  PXB_REQUIRE(out, "out is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const DevModel& D = model->dev;
  const int64_t nwarp = cells->ncell;
  const int64_t blocks = (nwarp + 7) / 8;
  PXB_REQUIRE(blocks < (1LL << 31), "too many cells for one call");
  const bool fast = workspace != nullptr && (D.prim.pre_n || (D.has_fb && D.fb.pre_n));
  if (fast) {
    int64_t need = 0;
    pxb_thermal_band_workspace_bytes(cells->ncell, &need);
    PXB_REQUIRE(workspace_bytes >= need, "band workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)need);
    unsigned long long* count = (unsigned long long*)workspace;
    int64_t* list = (int64_t*)workspace + 2;
    PXB_CUDA(cudaMemsetAsync(workspace, 0, 2 * sizeof(int64_t), st));
    k_thermal_band_fast<<<(unsigned)((cells->ncell + 255) / 256), 256, 0, st>>>(
        D, *cells, use_energy, emin, emax, shift, out, list, count);
    PXB_LAUNCH_CHECK();
    k_thermal_band<<<(unsigned)blocks, 256, 0, st>>>(D, *cells, use_energy, emin, emax, shift, out, list, count);
    PXB_LAUNCH_CHECK();
    return PXB_OK;
  }
  k_thermal_band<<<(unsigned)blocks, 256, 0, st>>>(D, *cells, use_energy, emin, emax, shift, out,
                                                  nullptr, nullptr);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

// ---- band flux per cell from tabulated row fluxes (row f2) -------------------------------
// scipy.interpolate.interp1d(kind="linear") as make_fluxf builds it (spectral_models.py:109-131):
// idx = clip(searchsorted(x, xn), 1, n-1); slope = (y_hi - y_lo) / (x_hi - x_lo);
// y = slope * (xn - x_lo) + y_lo -- same operations in the same order, no fused multiply-add.
__device__ __forceinline__ int interp1d_hi(const double* __restrict__ x, int n, double xn) {
  int lo = 0, hi = n;   // first index with x[idx] >= xn  (searchsorted side="left")
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(x + mid) < xn) lo = mid + 1; else hi = mid;
  }
  return min(max(lo, 1), n - 1);
}
__device__ __forceinline__ double interp1d_eval(const double* __restrict__ x,
                                                const double* __restrict__ y, int hi, double xn) {
  const double xl = __ldg(x + hi - 1), xh = __ldg(x + hi);
  const double yl = __ldg(y + hi - 1), yh = __ldg(y + hi);
  const double slope = __ddiv_rn(__dsub_rn(yh, yl), __dsub_rn(xh, xl));
  return __dadd_rn(__dmul_rn(slope, __dsub_rn(xn, xl)), yl);
}

// The same interpolant the way scipy evaluates it for an N-D y (interp1d._call_linear, used for
// the variable-element fluxes): y = ((xn - x_lo)/(x_hi - x_lo)) * y_hi + ((x_hi - xn)/(x_hi - x_lo)) * y_lo.
// (1-D y goes through np.interp, which is the slope form above.)
__device__ __forceinline__ double interp1d_eval_nd(const double* __restrict__ x,
                                                   const double* __restrict__ y, int hi, double xn) {
  const double xl = __ldg(x + hi - 1), xh = __ldg(x + hi);
  const double yl = __ldg(y + hi - 1), yh = __ldg(y + hi);
  const double dx = __dsub_rn(xh, xl);
  return __dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(xn, xl), dx), yh),
                   __dmul_rn(__ddiv_rn(__dsub_rn(xh, xn), dx), yl));
}

__global__ void __launch_bounds__(256)
k_thermal_flux(const DevModel M, const __grid_constant__ pxb_thermal_cells C,
               const pxb_band_flux B, double* __restrict__ out,
               unsigned long long* __restrict__ n_outside) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C.ncell) return;
  const double kT = C.kT[i];
  if (!cell_cut(C, i, kT)) {
    out[i] = 0.0;
    return;
  }
  const double xh = cell_xh(C, i);
  const double Z = cell_metal(C, i, xh);
  const double nH = C.nH ? C.nH[i] : 0.0;
  bool use_prim = true;
  if (M.density_dependent && M.has_fb)
    use_prim = (kT >= M.kT_lo) && (kT <= M.kT_hi) && (nH >= M.nH_lo) && (nH <= M.nH_hi);
  double tot;
  if (use_prim && M.density_dependent) {
    // _get_flux_2d, spectral_models.py:516-537 (indices from searchsorted - 1, no clipping: the
    // cell is inside the table), then / nH (:500-504)
    const DevTable& t = M.prim;
    const double lkT = log10(kT * M.K_per_keV), lnH = log10(nH);
    int ti = 0, di = 0;
    {
      int lo = 0, hi = t.nx;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(t.xnodes + mid) < lkT) lo = mid + 1; else hi = mid; }
      ti = lo - 1;
      lo = 0; hi = t.ny;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(t.ynodes + mid) < lnH) lo = mid + 1; else hi = mid; }
      di = lo - 1;
    }
    // searchsorted - 1 is -1 exactly on the first node: numpy then wraps to the LAST interval
    // (negative index); the reference inherits that, a lookup there is meaningless -> count it
    if (ti < 0 || di < 0 || ti > t.nx - 2 || di > t.ny - 2) {
      atomicAdd(n_outside, 1ull);
      out[i] = __longlong_as_double(0x7ff8000000000000LL);
      return;
    }
    const double dT = (lkT - __ldg(t.xnodes + ti)) / (__ldg(t.xnodes + ti + 1) - __ldg(t.xnodes + ti));
    const double dn = (lnH - __ldg(t.ynodes + di)) / (__ldg(t.ynodes + di + 1) - __ldg(t.ynodes + di));
    const int i1 = (di + 1) * t.nx + ti + 1, i2 = (di + 1) * t.nx + ti, i3 = di * t.nx + ti + 1,
              i4 = di * t.nx + ti;
    const double dx1 = __dmul_rn(dT, dn), dx2 = dn - dx1, dx3 = dT - dx1, dx4 = 1.0 + dx1 - dT - dn;
    auto bil = [&](const double* f) {
      return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx1, __ldg(f + i1)), __dmul_rn(dx2, __ldg(f + i2))),
                                 __dmul_rn(dx3, __ldg(f + i3))),
                       __dmul_rn(dx4, __ldg(f + i4)));
    };
    tot = bil(B.cflux) / nH;
    tot = __dadd_rn(tot, __dmul_rn(Z, bil(B.mflux) / nH));
    double vs = 0.0;
    for (int k = 0; k < C.nvar; ++k)
      vs = __dadd_rn(vs, __dmul_rn(cell_elem(C, k, i, xh), bil(B.vflux + (size_t)k * t.nrows) / nH));
    if (C.nvar) tot = __dadd_rn(tot, vs);
  } else {
    const DevTable& t = use_prim ? M.prim : M.fb;
    const int lt = use_prim ? M.log_t : M.fb_log_t;
    const double* cf = use_prim ? B.cflux : B.fb_cflux;
    const double* mf = use_prim ? B.mflux : B.fb_mflux;
    const double* vf = use_prim ? B.vflux : B.fb_vflux;
    const double x = lt ? log10(kT * M.K_per_keV) : kT;
    if (!(x >= __ldg(t.xnodes)) || !(x <= __ldg(t.xnodes + t.nx - 1))) {   // interp1d raises here
      atomicAdd(n_outside, 1ull);
      out[i] = __longlong_as_double(0x7ff8000000000000LL);
      return;
    }
    const int hi = interp1d_hi(t.xnodes, t.nx, x);
    tot = interp1d_eval(t.xnodes, cf, hi, x);
    tot = __dadd_rn(tot, __dmul_rn(Z, interp1d_eval(t.xnodes, mf, hi, x)));
    double vs = 0.0;
    for (int k = 0; k < C.nvar; ++k)
      vs = __dadd_rn(vs, __dmul_rn(cell_elem(C, k, i, xh),
                                   interp1d_eval_nd(t.xnodes, vf + (size_t)k * t.nrows, hi, x)));
    if (C.nvar) tot = __dadd_rn(tot, vs);
  }
  out[i] = __dmul_rn(tot, cell_norm(C, i));
}

extern "C" int pxb_thermal_flux(const pxb_model* model, const pxb_thermal_cells* cells,
                                const pxb_band_flux* band, double* out, int64_t* n_outside_out,
                                void* stream) {
  PXB_REQUIRE(model && cells && band && n_outside_out, "NULL argument");
  PXB_REQUIRE(cells->nvar == model->dev.nvar, "cells.nvar does not match the table");
  if (cells->ncell == 0) return PXB_OK;
  PXB_REQUIRE(out && cells->kT && cells->em && band->cflux && band->mflux, "NULL argument");
  PXB_REQUIRE(model->dev.nvar == 0 || band->vflux, "variable-element fluxes missing");
  PXB_REQUIRE(!model->dev.density_dependent || cells->nH, "density dependent model needs nH");
  PXB_REQUIRE(!(model->dev.density_dependent && model->dev.has_fb) || (band->fb_cflux && band->fb_mflux),
              "fallback-table fluxes missing");
  k_thermal_flux<<<(unsigned)((cells->ncell + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      model->dev, *cells, *band, out, (unsigned long long*)n_outside_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_powerlaw_spectrum(int64_t ncell, const double* alpha, const double* K,
                                     const double* shift, int64_t nbins, const double* emid,
                                     double* spec_out, void* stream) {
  PXB_REQUIRE(ncell >= 0 && nbins >= 1, "bad sizes");
  if (ncell == 0) return PXB_OK;
  PXB_REQUIRE(alpha && K && emid && spec_out, "NULL argument");
  const dim3 grid((unsigned)((nbins + 255) / 256), (unsigned)((ncell + SLAB - 1) / SLAB));
  PXB_REQUIRE(grid.y <= 65535, "too many cells for one call (max %d)", 65535 * SLAB);
  k_powerlaw_spectrum<<<grid, 256, 0, (cudaStream_t)stream>>>(ncell, alpha, K, shift, nbins, emid, spec_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_line_spectrum(int64_t ncell, double e0, const double* sigma, const double* N,
                                 const double* shift, int64_t nbins, const double* ee, int64_t ng,
                                 const double* gx, const double* gpdf, double* spec_out,
                                 void* stream) {
  PXB_REQUIRE(ncell >= 0 && nbins >= 1 && ng >= 2, "bad sizes");
  if (ncell == 0) return PXB_OK;
  PXB_REQUIRE(sigma && N && ee && gx && gpdf && spec_out, "NULL argument");
  const dim3 grid((unsigned)((nbins + 255) / 256), (unsigned)((ncell + SLAB - 1) / SLAB));
  PXB_REQUIRE(grid.y <= 65535, "too many cells for one call (max %d)", 65535 * SLAB);
  k_line_spectrum<<<grid, 256, 0, (cudaStream_t)stream>>>(ncell, e0, sigma, N, shift, nbins, ee, ng, gx,
                                                         gpdf, spec_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}
