// pipeline.cu -- the photon pipeline: energy sampling, projection, and both fused in one pass.
//
// Reference stages:
//   generate ... ThermalSourceModel.process_data photons branch, energies part,
//                pyxsim/source_models/thermal_sources/base.py:462-492,522-523
//   project .... the projection tile loop, pyxsim/photon_list.py:686-799 (doppler_shift,
//                absorb_photons, scatter_events / scatter_events_allsky, pixel_to_cel, eobs[det],
//                flux/emin/emax)
//
// ONE kernel template does all of it, in three modes that share every line of per-photon code:
//   MODE_SAMPLE   energies of the active cells' photons -> the photon list   (pxb_thermal_sample)
//   MODE_PROJECT  photon list -> event list                                  (pxb_project)
//   MODE_FUSED    active cells -> event list, no photon list in between      (pxb_make_events)
// so `make_events` produces bit for bit the events of `make_photons` + `project_photons`.
//
// Work decomposition.  The photons of a cell are taken in PAIRS (2m, 2m+1): a pair shares its
// cell look-up, its per-cell data and its three Philox blocks (see projection.cuh), and a lane
// owns one pair per round.  Pairs are numbered cell by cell (qoff = exclusive prefix of
// ceil(N/2) over the active cells); a warp owns a tile of 128 consecutive pair slots (<= 256
// photons, four rounds) that it draws from a global ticket, so tiles are started in increasing
// order over the whole grid; the CTAs are persistent (one per SM, 32 warps) and keep the four
// alias rows of the temperature bin being worked on in shared memory (TMA bulk copies, one
// mbarrier, re-decided every 16 tiles per warp); the event list keeps the photon order through
// a warp-wide decoupled look-back over per-tile survivor counts (the reference's cursor `k`,
// sky_functions.pyx:87-105).
// DRAM traffic is the algorithmic minimum: per-cell records in, (energies in/out,) events out.
#include "sampling.cuh"
#include "projection.cuh"

enum { MODE_SAMPLE = 0, MODE_PROJECT = 1, MODE_FUSED = 2 };

constexpr int PL_WARPS = 32;                       // warps of a persistent CTA holding staged rows
#ifndef PXB_STATS_IN_SCATTER
#define PXB_STATS_IN_SCATTER 0
#endif
#ifndef PXB_SMALL_WARPS
#define PXB_SMALL_WARPS 8
#endif
constexpr int PL_WARPS_SMALL = PXB_SMALL_WARPS;     // warps of a CTA that stages nothing (32 warps per SM)
constexpr int PL_SMALL_PER_SM = 32 / PL_WARPS_SMALL;
static_assert(PL_WARPS_SMALL * PL_SMALL_PER_SM == 32 && PL_WARPS_SMALL < PL_WARPS, "small CTAs fill 32 warps per SM");
constexpr int PL_EPOCH = 16;                       // tiles per warp between two staging decisions

// Tile geometry per mode.  A warp tile is ROUNDS x 32 pair slots.  The sampler keeps nothing
// between rounds and takes 128-slot tiles; the modes that project keep a tile's energies in
// shared memory from the absorption decisions to the scatter step, DOUBLE buffered (a tile whose
// predecessors have not all reported yet is finished after the next tile's decisions instead of
// waiting): 2 x 2 KB per warp at 128 slots.  (PXB_PROJ_ROUNDS = 2 builds the 64-slot tiles of the
// first version: every per-tile cost -- ticket, barriers, cell starts, statistics -- is paid twice
// as often.)
#ifndef PXB_PROJ_ROUNDS
#define PXB_PROJ_ROUNDS 4
#endif
static_assert(PXB_PROJ_ROUNDS >= 1 && PXB_PROJ_ROUNDS <= 4, "one byte per round in a 32-bit word");
template <int MODE>
struct Tiling {
  static constexpr int ROUNDS = MODE == 0 ? 4 : PXB_PROJ_ROUNDS;
  static constexpr int TSLOTS = 32 * ROUNDS;       // pair slots per warp tile
  static constexpr int MAXC = TSLOTS + 4;          // cell starts staged per tile
  static constexpr int NBUF = MODE == 0 ? 1 : 2;
  // one tile buffer of a warp in shared memory: [energies of the tile's photons (projecting modes)
  // | MAXC cell starts | first cell c0 of the tile], addressed as 32-bit shared-memory offsets
  static constexpr unsigned EBYTES = MODE == 0 ? 0u : 2u * TSLOTS * (unsigned)sizeof(double);
  static constexpr unsigned BSTRIDE = EBYTES + MAXC * 4u + 16u;
  static_assert(BSTRIDE % 16 == 0 && (EBYTES + MAXC * 4u) % 16 == 0, "16-byte aligned tile buffers");
};
static inline int tslots_of(int mode) { return mode == 0 ? Tiling<0>::TSLOTS : Tiling<1>::TSLOTS; }

// everything the photon kernels need of an active cell: 64 bytes, four 16-byte loads
struct __align__(16) CellRec {
  long long gid;     // global cell id: the RNG counter key
  double shift;      // Doppler factor (project / fused)
  long long poff;    // first photon of the cell in the energy array (sample / project)
  long long nph;     // photons of the cell
  int z1;            // FastCell: first table row
  int meta;          // FastCell: see sampling.cuh
  long long c0, c1, c2;   // FastCell: cnt[3] (split modes) or the bits of cum[3] (probability mode)
};
static_assert(sizeof(CellRec) == 64, "CellRec layout");

struct PipeArgs {
  const CellRec* recs;
  const CellGeom* geom;
  const int64_t* qoff;        // [n_cells + 1] pair offsets
  const int64_t* tile_cell;   // [n_tiles + 1] cell holding the first slot of every tile
  const int64_t* idx;         // chunk-local index of every active cell (sample / fused)
  const int64_t* tabcum;      // general split side array
  int tab_stride;
  int use_staging;
  int64_t n_cells, n_pairs, n_tiles, n_photons;
  unsigned int* ticket;
  double* energies;           // sample: out; project: in
  double* xsky;
  double* ysky;
  double* eobs;
  double* los;
  int64_t* n_events;
  unsigned long long* bstatus;   // [n_batches] look-back words (a batch = the tiles of one CTA step)
  double* part_sum;           // [n_tiles] per-tile partial statistics
  double* part_min;
  double* part_max;
  int off_w;                  // byte offset of the per-warp tile buffers in dynamic shared memory
  int f32;                    // events are written as float (xsky / ysky minus sky0 / sky1)
  double sky0, sky1;
};

// ---------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier
// ---------------------------------------------------------------------------------------
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

// shared memory through 32-bit addresses (volatile: ordered among themselves and kept where
// they are written; the buffers are private to a warp, __syncwarp() separates writers and readers)
__device__ __forceinline__ int lds_s32(unsigned a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_s32(unsigned a, int v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v));
}
__device__ __forceinline__ void lds_f64x2(unsigned a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ void sts_f64x2(unsigned a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y));
}
__device__ __forceinline__ void lds_s64x2(unsigned a, long long& x, long long& y) {
  asm volatile("ld.shared.v2.s64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(a));
}
__device__ __forceinline__ void sts_s64x2(unsigned a, long long x, long long y) {
  asm volatile("st.shared.v2.s64 [%0], {%1, %2};" ::"r"(a), "l"(x), "l"(y));
}

// ---------------------------------------------------------------------------------------
// decoupled look-back over per-tile survivor counts
// ---------------------------------------------------------------------------------------
constexpr unsigned long long ST_AGG = 1ull << 62;   // tile aggregate available
constexpr unsigned long long ST_INC = 2ull << 62;   // inclusive prefix available
constexpr unsigned long long ST_MASK = 3ull << 62;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned int ld_ticket(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v));
}

// Exclusive prefix of the survivor counts of all earlier BATCHES (a batch = the tiles a CTA works
// on together, one per warp) -- decoupled look-back over one status word per batch: the CTA publishes
// its batch's count (ST_AGG) as soon as the absorption decisions of the batch are made, walks
// back over the earlier batches' words, adding counts until it meets an inclusive prefix, and
// publishes its own inclusive prefix (ST_INC).  One warp per CTA does this; with one batch per SM
// in flight the walk is a handful of 32-word windows.  Resumable: with blocking = false it returns
// false as soon as a needed batch has not reported yet, keeping its progress in (j, prefix).
// Progress: a CTA publishes its count before it looks back, batches are handed out in increasing
// order, and a batch whose count is missing is being worked on by a resident CTA whose decision
// phase waits for nothing.
__device__ __forceinline__ bool lookback_run(const unsigned long long* status, int64_t& j,
                                             int64_t& prefix, int lane, bool blocking) {
  unsigned nap = 32;
  for (;;) {
    if (j < 0) return true;                                   // walked past batch 0
    const int64_t idx = j - lane;
    const unsigned long long sv = (idx >= 0) ? ld_status(status + idx) : ST_INC;
    const unsigned inc = __ballot_sync(0xffffffffu, (sv & ST_MASK) == ST_INC);
    const unsigned emp = __ballot_sync(0xffffffffu, (sv & ST_MASK) == 0ull);
    const int first = inc ? (__ffs(inc) - 1) : 32;           // nearest batch with an inclusive prefix
    const unsigned need = first >= 31 ? 0xffffffffu : ((2u << first) - 1u);
    if (emp & need) {                                        // a needed batch has not reported yet
      if (!blocking) return false;
      __nanosleep(nap);
      nap = min(nap * 2u, 1024u);
      continue;
    }
    int64_t v = (lane <= first) ? (int64_t)(sv & ~ST_MASK) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    prefix += v;
    if (first < 32) {
      j = -1;
      return true;
    }
    j -= 32;
  }
}

// ---------------------------------------------------------------------------------------
// per-cell preparation
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void rec_set_fast(CellRec& r, const FastCell& fc) {
  r.z1 = fc.z1;
  r.meta = fc.meta;
  r.c0 = fc.cnt[0];
  r.c1 = fc.cnt[1];
  r.c2 = fc.cnt[2];
}

// every pair tile that starts inside the cell's slot range starts in this cell
__device__ __forceinline__ void mark_tiles(int64_t* tile_cell, int64_t qa, int64_t qb, int64_t k,
                                           int tslots) {
  PXB_DASSERT(qa >= 0 && qb >= qa);
  for (int64_t t = (qa + tslots - 1) / tslots; t * tslots < qb; ++t) tile_cell[t] = k;
}

// Sampling side: one thread per ACTIVE cell computes the cell's mixture weights and, where it
// pays, the multinomial split of its photon count over its components (sampling.cuh), and writes
// the cell's record.  Cells the mixture path cannot take (extrapolating beyond the table, negative
// abundance) are appended to the general kernel's list; cells that want the general split are
// listed for the compacted second pass.
__global__ void __launch_bounds__(256)
k_prep_sample(const DevModel M, const __grid_constant__ pxb_thermal_cells C, int64_t n_active,
              const int64_t* __restrict__ idx, const int64_t* __restrict__ poff,
              const int64_t* __restrict__ qoff, CellRec* __restrict__ recs,
              int64_t* __restrict__ tile_cell, int64_t* __restrict__ list,
              unsigned long long* __restrict__ list_count, int64_t* __restrict__ tabcum,
              int64_t* __restrict__ glist, unsigned long long* __restrict__ glist_count,
              int split_factor, int tslots, int tab_stride) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool want_general = false;
  if (k < n_active) {
    const int64_t a = poff[k], b = poff[k + 1];
    const int64_t i = idx[k];
    FastCell fc = fast_setup<false>(M, C, i, b - a, tabcum ? tabcum + k * tab_stride : nullptr, split_factor);
    want_general = (fc.meta & 128) != 0;
    fc.meta &= ~128;
    CellRec r;
    r.gid = (long long)cell_gid(C, i);
    r.shift = 1.0;
    r.poff = a;
    r.nph = b - a;
    rec_set_fast(r, fc);
    recs[k] = r;
    if (!(fc.meta & 16) && b > a) {
      const unsigned long long slot = atomicAdd(list_count, 1ull);
      list[slot] = k;
    }
    mark_tiles(tile_cell, qoff[k], qoff[k + 1], k, tslots);
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

// second pass of the sampling-side preparation: the general multinomial split for the cells that
// asked for it (all lanes busy with binomial draws instead of one in a warp)
__global__ void __launch_bounds__(128)
k_prep_split(const DevModel M, const __grid_constant__ pxb_thermal_cells C,
             const int64_t* __restrict__ idx, const int64_t* __restrict__ poff,
             CellRec* __restrict__ recs, int64_t* __restrict__ tabcum, int tab_stride,
             const int64_t* __restrict__ glist, const unsigned long long* __restrict__ glist_count) {
  const unsigned long long n = *glist_count;
  for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (unsigned long long)gridDim.x * blockDim.x) {
    const int64_t k = glist[j];
    const FastCell fc = fast_setup<true>(M, C, idx[k], poff[k + 1] - poff[k], tabcum + k * tab_stride, 0);
    CellRec r = recs[k];
    rec_set_fast(r, fc);
    recs[k] = r;
  }
}

// Projection side: Doppler factor, rotated centre and scatter width of every active cell
// (photon_list.py:686-697,733-734,758-760).  INIT: the record is created here (project mode: the
// list comes from a file); otherwise only its Doppler factor is filled in (fused mode).
template <bool INIT>
__global__ void __launch_bounds__(256)
k_prep_project(const __grid_constant__ pxb_photon_list L, const __grid_constant__ pxb_project_params P,
               CellRec* __restrict__ recs, CellGeom* __restrict__ geom,
               int64_t* __restrict__ tile_cell, int tslots) {
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
  if (INIT) {
    CellRec r;
    r.gid = L.cell_id ? L.cell_id[c] : P.cell_id0 + c;
    r.shift = shift;
    r.poff = L.poff[c];
    r.nph = L.poff[c + 1] - L.poff[c];
    r.z1 = 0;
    r.meta = 16;
    r.c0 = r.c1 = r.c2 = 0;
    recs[c] = r;
    mark_tiles(tile_cell, L.qoff[c], L.qoff[c + 1], c, tslots);
  } else {
    recs[c].shift = shift;
  }
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
  geom[c] = g;
}

// ---------------------------------------------------------------------------------------
// energy sampling, general path: one CTA per declined cell
// ---------------------------------------------------------------------------------------
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
    // sampling: one Philox block -> the photon pair (2p, 2p+1)
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

// ---------------------------------------------------------------------------------------
// the photon kernel
// ---------------------------------------------------------------------------------------
// SPEC = 0: every option is read from the parameter blocks.  SPEC = 1 / 2: the common option set
// compiled in -- sampling: plain 1-D table without variable elements (every eligible cell is in
// four-component split mode), inverse-CDF sampling, uniformly spaced linear bins; projection:
// wabs with a foreground column only, cells, TAN sky, no sigma_pos, on-axis (1) / off-axis (2)
// normal.  Same results as SPEC = 0, fewer instructions and registers.
// WARPS: warps per CTA -- 32 (one persistent CTA per SM, room for the staged rows) or 8 (four CTAs
// per SM, nothing staged: while one CTA sits at a barrier the others issue).
template <int MODE, int SPEC, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == PL_WARPS ? 1 : PL_SMALL_PER_SM)
k_pipeline(const DevModel M, const __grid_constant__ pxb_thermal_cells C,
           const __grid_constant__ pxb_project_params P, const __grid_constant__ PhiloxKeys RKE,
           const __grid_constant__ PhiloxKeys RKP, const __grid_constant__ PipeArgs A,
           const SkyConst K) {
  constexpr int ROUNDS = Tiling<MODE>::ROUNDS, TSLOTS = Tiling<MODE>::TSLOTS;
  constexpr int MAXC = Tiling<MODE>::MAXC, NBUF = Tiling<MODE>::NBUF;
  constexpr unsigned EBYTES = Tiling<MODE>::EBYTES, BSTRIDE = Tiling<MODE>::BSTRIDE;
  constexpr unsigned HDR = EBYTES + MAXC * 4u;        // first cell of the buffer's tile (16 bytes)
  // dynamic shared memory: [4 staged alias rows | per-warp tile buffers (Tiling<MODE>)]
  extern __shared__ __align__(128) unsigned char dyn[];
  const DevTable& T = M.prim;
  uint2* s_alias = reinterpret_cast<uint2*>(dyn);
  __shared__ uint64_t s_bar;
  __shared__ int s_want;
  __shared__ long long s_batch[2];                     // batch per buffer (-1: none)
  __shared__ long long s_base[2];                      // its exclusive prefix, once known
  __shared__ int s_cnt[2][WARPS];                      // survivors of its tiles
  __shared__ long long s_lbj, s_lbprefix;              // warp 0: look-back progress of the pending batch
  __shared__ WabsFast s_wabs;

  // (lane and the warp's buffer address are pinned in registers: read through volatile asm, the
  // compiler cannot re-derive them from the thread index at every use)
  unsigned lane, wbase;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
  const int wid = threadIdx.x >> 5;
  {
    const unsigned w0 = smem_u32(dyn) + (unsigned)A.off_w + (unsigned)wid * (unsigned)(NBUF * BSTRIDE);
    asm volatile("mov.u32 %0, %1;" : "=r"(wbase) : "r"(w0));
  }
  if (MODE != MODE_SAMPLE && P.absorb_kind == PXB_ABS_WABS) wabs_fast_init(s_wabs);
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  int tag = -1;
  uint32_t parity = 0;
  const bool absorb = MODE != MODE_SAMPLE && P.absorb_kind != PXB_ABS_NONE;
  // rows are staged by the big CTAs only
  const bool staging = WARPS == PL_WARPS && MODE != MODE_PROJECT && A.use_staging;

  // (cell relative to c0, pair index inside the cell) of slot r*32 + lane of a tile: from the
  // packed ranks and the staged cell starts -- or, relpack == REL_SEARCH (lists with empty cells,
  // which the ballot look-up cannot rank; a first cell that starts 2^30 slots before the tile), by
  // bisection in the global offsets
  constexpr unsigned REL_SEARCH = 0xffffffffu;          // (no rank byte is 255: MAXC < 255)
  auto locate = [&](unsigned relpack, unsigned bb, int64_t tile, int64_t c0, int r, int srel,
                    int& rel, int64_t& m) {
    if (relpack != REL_SEARCH) {
      rel = (int)((relpack >> (8 * r)) & 255u);
      m = (int64_t)(srel - lds_s32(bb + EBYTES + 4u * (unsigned)rel));
    } else {
      const int64_t slot0 = tile * TSLOTS;
      const int nslots = (int)min((int64_t)TSLOTS, A.n_pairs - slot0);
      const int64_t c1 = (tile + 1 < A.n_tiles) ? __ldg(A.tile_cell + tile + 1) : A.n_cells - 1;
      const int64_t slot = slot0 + min(srel, nslots - 1);
      int64_t lo = c0, hi = c1 + 1;   // qoff[lo] <= slot < qoff[hi]
      while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(A.qoff + mid) <= slot) lo = mid; else hi = mid;
      }
      rel = (int)(lo - c0);
      m = slot0 + srel - __ldg(A.qoff + lo);
    }
  };

  // ---- a tile's decisions: energies (sampled or loaded), Doppler shift, absorption -------------
  // (writes the tile's energies, cell starts and first cell into buffer `buf`; returns its survivor
  // count and the per-lane survivor bits / cell ranks the scatter step needs)
  auto decide = [&](int64_t tile, int buf, int& run, unsigned& bits, unsigned& rels) {
    const unsigned bb = wbase + (unsigned)buf * BSTRIDE;
    const int64_t slot0 = tile * TSLOTS;
    const int nslots = (int)min((int64_t)TSLOTS, A.n_pairs - slot0);
    const int64_t c0 = __ldg(A.tile_cell + tile);
    // cell starts of the tile, relative to its first slot (the first cell's is <= 0)
    int nc;
    {
      const int64_t c1 = (tile + 1 < A.n_tiles) ? __ldg(A.tile_cell + tile + 1) : A.n_cells - 1;
      const int64_t ncl = c1 - c0 + 1;
      nc = ncl <= MAXC ? (int)ncl : 0;
    }
    __syncwarp();
    bool search = nc == 0;
    {
      bool bad = false;
      for (int q = (int)lane; q < nc; q += 32) {
        const int64_t a = __ldg(A.qoff + c0 + q) - slot0, b = __ldg(A.qoff + c0 + q + 1) - slot0;
        sts_s32(bb + EBYTES + 4u * (unsigned)q, (int)a);
        bad |= (a == b) || a < -(int64_t)(1 << 30);      // empty cell; start too far back for 32 bits
      }
      if (__any_sync(0xffffffffu, bad)) search = true;
      if (MODE != MODE_SAMPLE && lane == 0) sts_s64x2(bb + HDR, c0, 0ll);
    }
    __syncwarp();
    // Rank of the cell of every pair slot of the tile among the staged cells, for all rounds at
    // once (one byte per round; MAXC < 255): the number of cells starting at or before the slot,
    // minus one -- a ballot for the cells that start before the round's 32 slots plus a warp-wide
    // OR of start bits inside them.
    unsigned relpack = REL_SEARCH;
    if (!search) {
      unsigned cnt = 0;
      for (int qb = 0; qb < nc; qb += 32) {
        const int q = qb + (int)lane;
        const int qs = q < nc ? lds_s32(bb + EBYTES + 4u * (unsigned)q) : 0x7fffffff;
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
          const unsigned before = __ballot_sync(0xffffffffu, qs < r * 32);
          const unsigned d = (unsigned)(qs - r * 32);
          const unsigned mask = __reduce_or_sync(0xffffffffu, d < 32u ? (1u << d) : 0u);
          cnt += (unsigned)(__popc(before) + __popc(mask & ((2u << lane) - 1u))) << (8 * r);
        }
      }
      relpack = cnt - (0x01010101u >> (8 * (4 - ROUNDS)));   // (every byte in use is >= 1)
    }

    unsigned detpack = 0;
#if !PXB_STATS_IN_SCATTER
    double tsum = 0.0, tmin = 1.0e99, tmax = -1.0e99;
#endif
#pragma unroll 1
    for (int r = 0; r < ROUNDS; ++r) {
      if (r * 32 >= nslots) break;
      const int srel = r * 32 + (int)lane;
      const bool valid = srel < nslots;
      int rel;
      int64_t m;
      locate(relpack, bb, tile, c0, r, srel, rel, m);
      PXB_DASSERT(rel >= 0 && c0 + rel < A.n_cells && (!valid || m >= 0));
      const CellRec* rp = A.recs + (c0 + rel);
      const longlong2 q0 = __ldg(reinterpret_cast<const longlong2*>(rp));
      const longlong2 q1 = __ldg(reinterpret_cast<const longlong2*>(rp) + 1);
      const uint64_t gid = (uint64_t)q0.x;
      const uint64_t d0 = 2ull * (uint64_t)m;
      const bool hasA = valid;
      const bool hasB = valid && (int64_t)(d0 + 1) < q1.y;
      double eA = 0.0, eB = 0.0;
      bool elig = true;
      if (MODE != MODE_PROJECT) {
        const longlong2 q2 = __ldg(reinterpret_cast<const longlong2*>(rp) + 2);
        const longlong2 q3 = __ldg(reinterpret_cast<const longlong2*>(rp) + 3);
        const int z1 = (int)(q2.x & 0xffffffffll), meta = (int)(q2.x >> 32);
        const long long cn0 = q2.y, cn1 = q3.x, cn2 = q3.y;
        elig = (meta & 16) != 0;
        if (valid && elig) {
          if (SPEC != 0 || (meta & (32 | 64))) {
            // split modes: the photon's component follows from its index inside the cell
            const Philox4 E = pxb_rand4(RKE, STREAM_ENERGY, gid, (uint64_t)m);
            const bool prim = SPEC != 0 || !(meta & 8);
            const DevTable* tb = prim ? &M.prim : &M.fb;
            const bool four = SPEC != 0 || (meta & 32);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              // (a cell's last pair may lack its second photon: that lane half repeats the first
              // one's look-ups instead of indexing past the cell's components)
              // (eligible cells have fewer than 2^31 photons, fast_setup: 32-bit compares)
              const int dd = (int)d0 + ((h && hasB) ? 1 : 0);
              const int comp = (dd >= (int)cn0) + (dd >= (int)cn1) + (dd >= (int)cn2);
              int tab, row;
              if (four) {
                // four components: (row0,cosmic), (row0,metal), (row1,cosmic), (row1,metal)
                tab = comp & 1;
                row = z1 + (comp >> 1);
              } else {
                // general split: comp is the ROW, the table comes from the side array
                const long long din = dd - (comp == 0 ? 0 : (comp == 1 ? (int)cn0 : (comp == 2 ? (int)cn1 : (int)cn2)));
                const int ntab = 2 + C.nvar;
                const int64_t* tc = A.tabcum + (c0 + rel) * A.tab_stride + comp * ntab;
                tab = 0;
                if (ntab <= 8) {
                  for (int t2 = 0; t2 < ntab - 1; ++t2) tab += (din >= __ldg(tc + t2)) ? 1 : 0;
                } else {
                  int lo = 0, hi = ntab - 1;     // first tab with din < tc[tab]
                  while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (din >= __ldg(tc + mid)) lo = mid + 1; else hi = mid;
                  }
                  tab = lo;
                }
                row = fast_row(tb, z1, meta & 7, comp);
              }
              const uint32_t w0 = h ? E.z : E.x, w1 = h ? E.w : E.y;
              uint32_t v;
              int j;
              PXB_DASSERT(tab >= 0 && tab < 2 + C.nvar && row >= 0 && row < tb->nrows);
              const unsigned dr = (unsigned)(row - tag);
              if (WARPS == PL_WARPS && prim && tab < 2 && tag >= 0 && dr < 2u)   // staged row: shared-memory gather
                j = alias_pick(s_alias, (unsigned)(tab * 2 + (int)dr) * (unsigned)T.alias_stride,
                               (unsigned)T.nbins, w0, w1, v);
              else
                j = alias_pick(tb->alias, alias_row_index(tb, tab, row), (unsigned)T.nbins, w0, w1, v);
              const double e = alias_energy<(SPEC != 0) ? 1 : 0>(M, j, v);
              if (h) eB = e; else eA = e;
            }
          } else {
            // probability mode: every photon picks its component itself (one block per photon)
            FastCell fc;
            fc.z1 = z1;
            fc.meta = meta;
            fc.cnt[0] = cn0; fc.cnt[1] = cn1; fc.cnt[2] = cn2;
            const int64_t i = __ldg(A.idx + c0 + rel);
            const double* th = (meta & 256)
                ? reinterpret_cast<const double*>(A.tabcum + (c0 + rel) * A.tab_stride) : nullptr;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (h && !hasB) break;
              const Philox4 rn = pxb_rand4(RKE, STREAM_ENERGY, gid, d0 + h);
              const double e = fast_draw<0>(M, C, fc, i, th, u53(rn.x, rn.y), rn.z, rn.w);
              if (h) eB = e; else eA = e;
            }
          }
          if (SPEC == 0 && M.binscale_log) {
            eA = exp10(eA);
            eB = exp10(eB);
          }
        }
      }
      if (MODE == MODE_SAMPLE) {
        if (elig) {
          double* dst = A.energies + q1.x + (int64_t)d0;
          PXB_DASSERT(!valid || (q1.x >= 0 && q1.x + (int64_t)d0 + (hasB ? 1 : 0) < A.n_photons));
          if (hasB && (((uintptr_t)dst) & 15) == 0) {
            asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(dst), "d"(eA), "d"(eB));
          } else {
            if (hasA) st_stream(dst, eA);
            if (hasB) st_stream(dst + 1, eB);
          }
        }
        continue;
      }
      if (MODE == MODE_PROJECT) {
        const double* src = A.energies + q1.x + (int64_t)d0;
        PXB_DASSERT(!valid || (q1.x >= 0 && q1.x + (int64_t)d0 + (hasB ? 1 : 0) < A.n_photons));
        if (hasA) eA = ld_stream(src);
        if (hasB) eB = ld_stream(src + 1);
      }
      const double shift = __longlong_as_double(q0.y);
      eA *= shift;
      eB *= shift;
      bool detA = hasA && elig, detB = hasB && elig;
      if (absorb) {
        const Philox4 Aw = pxb_rand4(RKP, STREAM_PROJ_A, gid, (uint64_t)m);
        Philox4 Dw{0u, 0u, 0u, 0u};
        if (SPEC == 0 && P.nH_cell) Dw = pxb_rand4(RKP, STREAM_PROJ_D, gid, (uint64_t)m);
        if (detA) detA = absorb_decision<SPEC>(P, RKP, s_wabs, c0 + rel, gid, d0, Aw.x, Dw.x, eA);
        if (detB) detB = absorb_decision<SPEC>(P, RKP, s_wabs, c0 + rel, gid, d0 + 1, Aw.y, Dw.y, eB);
      }
      sts_f64x2(bb + 16u * (unsigned)srel, eA, eB);
#if !PXB_STATS_IN_SCATTER
      if (detA) {
        tsum += eA;
        tmin = eA < tmin ? eA : tmin;   // (energies are never NaN: plain compares)
        tmax = eA > tmax ? eA : tmax;
      }
      if (detB) {
        tsum += eB;
        tmin = eB < tmin ? eB : tmin;
        tmax = eB > tmax ? eB : tmax;
      }
#endif
      run += __popc(__ballot_sync(0xffffffffu, detA)) + __popc(__ballot_sync(0xffffffffu, detB));
      detpack |= ((detA ? 1u : 0u) | (detB ? 2u : 0u)) << (2 * r);
    }

    if (MODE != MODE_SAMPLE) {
#if !PXB_STATS_IN_SCATTER
      tsum = warp_sum(tsum);
      tmin = warp_min(tmin);
      tmax = warp_max(tmax);
      if (lane == 0) {
        A.part_sum[tile] = tsum;
        A.part_min[tile] = tmin;
        A.part_max[tile] = tmax;
      }
#endif
      bits = detpack;
      rels = relpack;
    }
  };

  // ---- a tile's scatter step: survivors -> their final, order-preserving slots ------------------
  auto scatter = [&](int64_t ptile, int pbuf, int64_t base, unsigned detpack, unsigned relpack) {
    const unsigned bb = wbase + (unsigned)pbuf * BSTRIDE;
    long long c0, unused;
    lds_s64x2(bb + HDR, c0, unused);
    // (float events: half the bytes -- the float arrays are indexed from the same pointers)
    const unsigned lt = (1u << lane) - 1u;
    int done = 0;
#if PXB_STATS_IN_SCATTER
    double tsum = 0.0, tmin = 1.0e99, tmax = -1.0e99;   // the tile's survivors: flux / emin / emax partials
#endif
#pragma unroll 1
    for (int r = 0; r < ROUNDS; ++r) {
      const bool detA = (detpack >> (2 * r)) & 1u, detB = (detpack >> (2 * r + 1)) & 1u;
      const unsigned balA = __ballot_sync(0xffffffffu, detA), balB = __ballot_sync(0xffffffffu, detB);
      if (!(balA | balB)) continue;
      const int o0 = done + __popc(balA & lt) + __popc(balB & lt);
      done += __popc(balA) + __popc(balB);
      if (!(detA || detB)) continue;
      const int srel = r * 32 + (int)lane;
      int rel;
      int64_t m;
      locate(relpack, bb, ptile, c0, r, srel, rel, m);
      const uint64_t gid = (uint64_t)__ldg(&A.recs[c0 + rel].gid);
      const uint64_t d0 = 2ull * (uint64_t)m;
      const CellGeom g = ld_geom(A.geom + c0 + rel);
      const Philox4 Bw = pxb_rand4(RKP, STREAM_PROJ_B, gid, (uint64_t)m);
      // the third scatter word lives in the absorption block: recomputed only where it is
      // used (off-axis cells, particle kernels, all-sky)
      Philox4 Aw{0u, 0u, 0u, 0u};
      if (SPEC == 2 || (SPEC == 0 && !(P.observer == 0 && P.data_type == 0 && P.vn_axis >= 0)))
        Aw = pxb_rand4(RKP, STREAM_PROJ_A, gid, (uint64_t)m);
      double eA, eB;
      lds_f64x2(bb + 16u * (unsigned)srel, eA, eB);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!(h ? detB : detA)) continue;
        double xs, ys, los;
        const uint32_t w0 = h ? Bw.z : Bw.x, w1 = h ? Bw.w : Bw.y, w2 = h ? Aw.w : Aw.z;
        if (SPEC != 0 || P.observer == 0)
          scatter_external<SPEC>(P, RKP, g, K, w0, w1, w2, gid, d0 + h, xs, ys, los);
        else
          scatter_allsky(P, RKP, g, w0, w1, w2, gid, d0 + h, xs, ys, los);
        const int64_t o = base + o0 + (h && detA ? 1 : 0);
        PXB_DASSERT(o >= 0 && o < A.n_photons && srel < TSLOTS);   // events never outnumber photons
        const double ev = h ? eB : eA;
#if PXB_STATS_IN_SCATTER
        tsum += ev;
        tmin = ev < tmin ? ev : tmin;   // (energies are never NaN: plain compares)
        tmax = ev > tmax ? ev : tmax;
#endif
        if (!A.f32) {
          st_stream(A.xsky + o, xs);
          st_stream(A.ysky + o, ys);
          st_stream(A.eobs + o, ev);
          if (A.los) st_stream(A.los + o, los);
        } else {
          // sky offsets from the fp64 centre and energies in single precision
          __stcs(reinterpret_cast<float*>(A.xsky) + o, (float)(xs - A.sky0));
          __stcs(reinterpret_cast<float*>(A.ysky) + o, (float)(ys - A.sky1));
          __stcs(reinterpret_cast<float*>(A.eobs) + o, (float)ev);
          if (A.los) __stcs(reinterpret_cast<float*>(A.los) + o, (float)los);
        }
      }
    }
#if PXB_STATS_IN_SCATTER
    tsum = warp_sum(tsum);
    tmin = warp_min(tmin);
    tmax = warp_max(tmax);
    if (lane == 0) {
      A.part_sum[ptile] = tsum;
      A.part_min[ptile] = tmin;
      A.part_max[ptile] = tmax;
    }
#endif
  };

  // the rows of the temperature bin the grid is working in are (re)staged; every thread of the CTA
  // calls this at a point where no warp is sampling (tnext: the tile about to be handed out)
  auto restage = [&](int64_t tnext) {
    if (threadIdx.x == 0) {
      const CellRec* f0 = A.recs + __ldg(A.tile_cell + min(tnext, A.n_tiles - 1));
      const int meta0 = __ldg(&f0->meta);
      s_want = ((meta0 & 7) == 2 && !(meta0 & 8)) ? __ldg(&f0->z1) : -1;
    }
    __syncthreads();     // one decision for the whole CTA
    const int w = s_want;
    if (w >= 0 && w != tag) {
      const uint32_t row_bytes = (uint32_t)T.alias_stride * (uint32_t)sizeof(uint2);
      if (threadIdx.x == 0) {
        mbar_expect_tx(&s_bar, 4u * row_bytes);
#pragma unroll
        for (int slot = 0; slot < 4; ++slot)   // slot = 2 * table + (row - w)
          tma_bulk_g2s(s_alias + (size_t)slot * T.alias_stride,
                       alias_row(&T, slot >> 1, w + (slot & 1)), row_bytes, &s_bar);
      }
      mbar_wait(&s_bar, parity);
      parity ^= 1u;
      tag = w;
    }
  };

  if (MODE == MODE_SAMPLE) {
    // Sampling keeps nothing between tiles: every WARP draws its tiles one at a time from the
    // global ticket; the warps of a CTA only meet at the staging epochs.
    __shared__ int s_live;
    if (threadIdx.x == 0) s_live = WARPS;
    __syncthreads();
    bool out = false;
    for (int it = 0;; ++it) {
      if (staging && (it & (PL_EPOCH - 1)) == 0) {
        __syncthreads();     // every warp is done with the previously staged rows
        if (s_live == 0) break;
        restage((int64_t)ld_ticket(A.ticket));
      }
      if (out) {
        if (!staging) break;
        continue;
      }
      unsigned int tk = 0;
      if (lane == 0) tk = atomicAdd(A.ticket, 1u);
      const int64_t tile = (int64_t)__shfl_sync(0xffffffffu, tk, 0);
      if (tile >= A.n_tiles) {
        out = true;
        if (lane == 0) atomicSub(&s_live, 1);
        if (!staging) break;
        continue;
      }
      int run = 0;
      unsigned bits = 0, rels = 0;
      decide(tile, 0, run, bits, rels);
    }
    return;
  }

  // Projecting modes.  The CTA works on BATCHES of WARPS consecutive tiles (one per warp) that
  // it draws from the global ticket, in two half-steps per iteration separated by CTA barriers:
  //   decide(batch k) | scatter(batch k-1)
  // The survivor count of batch k is published right after its decisions; its exclusive prefix
  // over all earlier batches is looked up by warp 0 -- first without blocking, and for good one
  // iteration later, just before the scatter step that needs it -- so the latency of the
  // look-back hides behind the next batch's decisions, nobody spins on a per-warp basis, and the
  // barriers keep the warps of an SM in step (the scheduler's fixed warp priorities would
  // otherwise let some warps starve while everybody waits for their tiles).
  int buf = 0;
  bool have_pend = false;      // this warp's tile of the previous batch still has its scatter step due
  unsigned pend_bits = 0, pend_rels = 0;
  bool lbdone = true, cta_out = false, pend_batch = false;
  for (int it = 0;; ++it) {
    if (threadIdx.x == 0) s_batch[buf] = cta_out ? -1ll : (long long)atomicAdd(A.ticket, 1u);
    __syncthreads();     // (1) the batch is known; every warp is done with the previous iteration
    const int64_t batch = s_batch[buf];
    if (batch < 0 || batch * WARPS >= A.n_tiles) cta_out = true;
    if (staging && !cta_out && (it & (PL_EPOCH - 1)) == 0) restage(batch * WARPS);
    const int64_t tile = batch * WARPS + wid;
    const bool have_new = !cta_out && tile < A.n_tiles;
    int run = 0;
    unsigned bits = 0, rels = 0;
    if (have_new) decide(tile, buf, run, bits, rels);
    if (lane == 0) s_cnt[buf][wid] = run;
    if (wid == 0 && pend_batch && !lbdone) {
      // the pending batch's prefix, for good now: it has had a whole batch's time
      int64_t lbj = s_lbj, lbprefix = s_lbprefix;
      lookback_run(A.bstatus, lbj, lbprefix, lane, true);
      if (lane == 0) {
        int tot = 0;
        for (int w = 0; w < WARPS; ++w) tot += s_cnt[buf ^ 1][w];
        st_status(A.bstatus + s_batch[buf ^ 1], ST_INC | (unsigned long long)(lbprefix + tot));
        s_base[buf ^ 1] = lbprefix;
      }
      lbdone = true;
    }
    __syncthreads();     // (2) the new batch's counts and the pending batch's prefix are in place
    if (wid == 0 && !cta_out) {
      // publish the new batch's count; a first, non-blocking look at its predecessors
      const int cw = lane < WARPS ? s_cnt[buf][lane] : 0;
      const int tot = __reduce_add_sync(0xffffffffu, cw);
      if (lane == 0) st_status(A.bstatus + batch, (batch == 0 ? ST_INC : ST_AGG) | (unsigned long long)tot);
      int64_t lbj = batch - 1, lbprefix = 0;
      lbdone = lookback_run(A.bstatus, lbj, lbprefix, lane, false);
      if (lane == 0) {
        if (lbdone) {
          if (batch > 0) st_status(A.bstatus + batch, ST_INC | (unsigned long long)(lbprefix + tot));
          s_base[buf] = lbprefix;
        } else {
          s_lbj = lbj;
          s_lbprefix = lbprefix;
        }
      }
    }
    if (have_pend) {
      int before = 0, mine = 0;
      {
        const int cw = lane < WARPS ? s_cnt[buf ^ 1][lane] : 0;
        before = __reduce_add_sync(0xffffffffu, lane < wid ? cw : 0);
        mine = __shfl_sync(0xffffffffu, cw, wid);
      }
      const int64_t ptile = s_batch[buf ^ 1] * WARPS + wid;
      const int64_t base = s_base[buf ^ 1] + before;
      if (ptile == A.n_tiles - 1 && lane == 0) *A.n_events = base + mine;
      if (mine > 0) scatter(ptile, buf ^ 1, base, pend_bits, pend_rels);
#if PXB_STATS_IN_SCATTER
      else if (lane == 0) {
        A.part_sum[ptile] = 0.0;
        A.part_min[ptile] = 1.0e99;
        A.part_max[ptile] = -1.0e99;
      }
#endif
    }
    if (cta_out) break;          // nothing new was decided: the batch just scattered was the last
    have_pend = have_new;
    pend_batch = true;
    pend_bits = bits;
    pend_rels = rels;
    buf ^= 1;
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

// =======================================================================================
// host
// =======================================================================================
constexpr int STATS_BLOCKS = 256;

static inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

// side array of the general split: cumulative counts per (row, table) of every active cell
static int split_stride(const DevModel& D) {
  if (D.nvar == 0) return 0;        // no variable elements: no general split (see fast_setup)
  return (D.density_dependent ? 4 : 2) * (2 + D.nvar);
}

// workspace layout shared by the three entry points (a section is empty when its mode does not
// use it)
struct PipeLayout {
  int64_t ntiles;
  int tslots;          // pair slots per tile of the mode (Tiling<MODE>::TSLOTS)
  int64_t off_hdr, off_recs, off_geom, off_tilecell, off_bstatus, off_part, off_red;
  int64_t off_list, off_tabcum, off_glist, total;
};

static PipeLayout pipe_layout(const DevModel* D, int64_t n_cells, int64_t n_pairs, bool sample,
                              bool project) {
  PipeLayout l;
  l.tslots = tslots_of(project ? MODE_PROJECT : MODE_SAMPLE);
  l.ntiles = (n_pairs + l.tslots - 1) / l.tslots;
  int64_t o = 0;
  l.off_hdr = o;      o += 256;                                  // counters + ticket
  const int64_t nbatch = (l.ntiles + PL_WARPS_SMALL - 1) / PL_WARPS_SMALL;   // (enough for either CTA size)
  l.off_bstatus = o;  o += project ? align256(nbatch * 8) : 0;   // (header + look-back words: one memset)
  l.off_recs = o;     o += align256(n_cells * (int64_t)sizeof(CellRec));
  l.off_geom = o;     o += project ? align256(n_cells * (int64_t)sizeof(CellGeom)) : 0;
  l.off_tilecell = o; o += align256((l.ntiles + 1) * 8);
  l.off_part = o;     o += project ? align256(3 * l.ntiles * 8) : 0;
  l.off_red = o;      o += project ? align256(3 * STATS_BLOCKS * 8) : 0;
  const int64_t stride = (sample && D) ? split_stride(*D) : 0;
  l.off_list = o;     o += sample ? align256(n_cells * 8) : 0;
  l.off_tabcum = o;   o += stride ? align256(n_cells * stride * 8) : 0;
  l.off_glist = o;    o += stride ? align256(n_cells * 8) : 0;
  l.total = o;
  return l;
}

// header words at off_hdr: [0] declined-cell count, [1] ticket (uint), [2] general-split count
struct PipeHeader {
  unsigned long long* list_count;
  unsigned int* ticket;
  unsigned long long* glist_count;
};
static PipeHeader pipe_header(char* ws, const PipeLayout& l) {
  unsigned long long* h = (unsigned long long*)(ws + l.off_hdr);
  return PipeHeader{h, (unsigned int*)(h + 1), h + 2};
}

// shared-memory plan of a k_pipeline launch
struct SmemPlan {
  int use_staging, off_w, warps;
  size_t total;
};
// mode project stages nothing and runs small CTAs; sampling and the fused mode run one big CTA per
// SM with the four alias rows of the temperature bin staged when they fit (small_fused: the fused
// mode with small CTAs and the rows left to L1/L2 instead)
static SmemPlan smem_plan(const DevModel* D, int mode, bool small_fused = false) {
  SmemPlan s;
  s.warps = (mode == MODE_PROJECT || (mode == MODE_FUSED && small_fused)) ? PL_WARPS_SMALL : PL_WARPS;
  const size_t w_bytes = mode == MODE_SAMPLE
                             ? (size_t)s.warps * Tiling<0>::NBUF * Tiling<0>::BSTRIDE
                             : (size_t)s.warps * Tiling<1>::NBUF * Tiling<1>::BSTRIDE;
  size_t rows = 0;
  s.use_staging = 0;
  if (s.warps == PL_WARPS && mode != MODE_PROJECT && D && !D->density_dependent && !D->has_fb) {
    rows = 4 * (size_t)D->prim.alias_stride * sizeof(uint2);
    // rows are staged when the four of them fit beside the per-warp scratch
    if (rows + w_bytes + 2048 <= 227 * 1024) s.use_staging = 1;
    else rows = 0;
  }
  rows = (rows + 127) & ~(size_t)127;
  s.off_w = (int)rows;
  s.total = rows + w_bytes;
  return s;
}

// which variant of the photon kernel the last pxb_thermal_sample / pxb_project / pxb_make_events of
// this thread launched (diagnostic: lets a test assert that the path it means to test was taken)
static thread_local int g_last_mode = -1, g_last_spec = -1, g_last_staged = -1;
extern "C" int pxb_last_pipeline_launch(int* mode, int* spec, int* staged) {
  if (mode) *mode = g_last_mode;
  if (spec) *spec = g_last_spec;
  if (staged) *staged = g_last_staged;
  return g_last_mode < 0 ? PXB_ERR_INVALID : PXB_OK;
}

template <int MODE, int SPEC, int WARPS>
static int launch_pipeline_w(const DevModel& D, const pxb_thermal_cells& C, const pxb_project_params& P,
                             PipeArgs A, const SmemPlan& sp, cudaStream_t st) {
  static thread_local size_t attr_set = 0;
  if (attr_set != sp.total + 1) {   // once per (process, shared-memory size), not per call
    PXB_CUDA(cudaFuncSetAttribute(k_pipeline<MODE, SPEC, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sp.total));
    attr_set = sp.total + 1;
  }
  // persistent CTAs: one per SM (big CTAs) or four per SM (small ones)
  int64_t grid = (int64_t)pxb_sm_count() * (WARPS == PL_WARPS ? 1 : PL_SMALL_PER_SM);
  const int64_t enough = (A.n_tiles + WARPS - 1) / WARPS;
  if (grid > enough) grid = enough;
  const SkyConst K = make_sky_const(&P);
  g_last_mode = MODE;
  g_last_spec = SPEC;
  g_last_staged = MODE != MODE_PROJECT && A.use_staging;
  k_pipeline<MODE, SPEC, WARPS><<<(unsigned)grid, WARPS * 32, sp.total, st>>>(
      D, C, P, pxb_make_keys(C.seed), pxb_make_keys(P.seed), A, K);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}
template <int MODE, int SPEC>
static int launch_pipeline(const DevModel& D, const pxb_thermal_cells& C, const pxb_project_params& P,
                           PipeArgs A, const SmemPlan& sp, cudaStream_t st) {
  if (MODE != MODE_SAMPLE && sp.warps == PL_WARPS_SMALL)
    return launch_pipeline_w<MODE, SPEC, (MODE == MODE_SAMPLE ? PL_WARPS : PL_WARPS_SMALL)>(D, C, P, A, sp, st);
  return launch_pipeline_w<MODE, SPEC, PL_WARPS>(D, C, P, A, sp, st);
}

static bool sample_is_common(const DevModel& D, const pxb_thermal_cells* cells) {
  return D.method == 0 && !D.binscale_log && D.edges_uniform && D.nvar == 0 &&
         !D.density_dependent && !D.has_fb && !(cells->generic_kernels & 1);
}
// The fused kernel pays where sampling is cheap enough for the step to be bound by the photon
// list's memory traffic -- the common sampler.  With the option-generic sampler (variable elements,
// 2-D or log-binned tables, accept-reject) the step is bound by table gathers and the two
// stages, each with its own CTA shape, are faster (measured at the named sizes: amr 37.8 vs 41.3 ms,
// sph 29.3 vs 35.7 ms per step on 8 GPUs); callers ask here which route to take.
extern "C" int pxb_model_fusion_pays(const pxb_model* m) {
  if (!m) return 0;
  pxb_thermal_cells none{};
  return (m->dev.fast_ok && sample_is_common(m->dev, &none)) ? 1 : 0;
}
// 0: generic; 1: common on-axis; 2: common off-axis
static int project_spec(const pxb_project_params* P) {
  if (P->generic_kernels) return 0;
  const bool common = P->absorb_kind == PXB_ABS_WABS && !P->nH_cell && P->nH > 0.0 &&
                      P->observer == 0 && P->data_type == 0 && !P->has_sigma_pos &&
                      P->sky_mode == PXB_SKY_TAN;
  if (!common) return 0;
  return P->vn_axis >= 0 ? 1 : 2;
}

// ---- sampling -----------------------------------------------------------------------------
extern "C" int pxb_thermal_sample_workspace_bytes(const pxb_model* model, int64_t n_active,
                                                  int64_t n_pairs, int64_t* bytes) {
  PXB_REQUIRE(model && bytes && n_active >= 0 && n_pairs >= 0, "bad argument");
  *bytes = pipe_layout(&model->dev, n_active, n_pairs, true, false).total;
  return PXB_OK;
}

// sampling-side preparation shared by pxb_thermal_sample and pxb_make_events
static int prep_sample(const pxb_model* model, const pxb_thermal_cells* cells, int64_t n_active,
                       const int64_t* idx, const int64_t* poff, const int64_t* qoff, char* ws,
                       const PipeLayout& lay, int* tab_stride_out, int64_t** tabcum_out,
                       cudaStream_t st) {
  const DevModel& D = model->dev;
  const PipeHeader H = pipe_header(ws, lay);
  CellRec* recs = (CellRec*)(ws + lay.off_recs);
  int64_t* tile_cell = (int64_t*)(ws + lay.off_tilecell);
  int64_t* list = (int64_t*)(ws + lay.off_list);
  const int tab_stride = split_stride(D);
  int64_t* tabcum = tab_stride ? (int64_t*)(ws + lay.off_tabcum) : nullptr;
  int64_t* glist = tab_stride ? (int64_t*)(ws + lay.off_glist) : nullptr;
  if (cells->split_factor < 0) tabcum = nullptr;   // per-photon component choice
  // (photons per component from which the multinomial split is taken instead of the per-photon
  // choice from stored cumulative probabilities; measured on the amr configuration, 38 photons per
  // cell on average: sampling stage 59.2 ms at 6, 46.9 at 16, 43.6 at 40, 43.2 never splitting)
  const int split_factor = cells->split_factor > 0 ? cells->split_factor : 40;
  k_prep_sample<<<(unsigned)((n_active + 255) / 256), 256, 0, st>>>(
      D, *cells, n_active, idx, poff, qoff, recs, tile_cell, list, H.list_count, tabcum, glist,
      H.glist_count, split_factor, lay.tslots, tab_stride);
  PXB_LAUNCH_CHECK();
  if (tabcum) {
    int64_t gs = (n_active + 127) / 128;
    if (gs > (int64_t)pxb_sm_count() * 16) gs = (int64_t)pxb_sm_count() * 16;
    k_prep_split<<<(unsigned)gs, 128, 0, st>>>(D, *cells, idx, poff, recs, tabcum, tab_stride, glist,
                                               H.glist_count);
    PXB_LAUNCH_CHECK();
  }
  *tab_stride_out = tab_stride;
  *tabcum_out = tabcum;
  return PXB_OK;
}

extern "C" int pxb_thermal_sample(const pxb_model* model, const pxb_thermal_cells* cells,
                                  int64_t n_active, const int64_t* idx, const int64_t* nph,
                                  const int64_t* poff, const int64_t* qoff, int64_t n_photons,
                                  int64_t n_pairs, double* energies, void* workspace,
                                  int64_t workspace_bytes, void* stream) {
  int rc = pxb_check_cells(model, cells);
  if (rc) return rc;
  if (n_active == 0 || n_photons == 0) return PXB_OK;
  PXB_REQUIRE(idx && nph && poff && qoff && energies, "NULL argument");
  PXB_REQUIRE(n_pairs > 0 && 2 * n_pairs >= n_photons, "n_pairs does not match the photon count");
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
  int64_t grid = (int64_t)pxb_sm_count() * cta_per_sm;
  if (grid > n_active) grid = n_active;
  int64_t* list = nullptr;
  unsigned long long* count = nullptr;
  if (D.fast_ok) {
    const PipeLayout lay = pipe_layout(&D, n_active, n_pairs, true, false);
    PXB_REQUIRE(workspace && workspace_bytes >= lay.total, "sample workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)lay.total);
    char* ws = (char*)workspace;
    const PipeHeader H = pipe_header(ws, lay);
    PXB_REQUIRE(lay.ntiles < (1LL << 32) - (1 << 20), "too many photons for one call");
    PXB_CUDA(cudaMemsetAsync(ws + lay.off_hdr, 0, 256, st));
    int tab_stride = 0;
    int64_t* tabcum = nullptr;
    rc = prep_sample(model, cells, n_active, idx, poff, qoff, ws, lay, &tab_stride, &tabcum, st);
    if (rc) return rc;
    list = (int64_t*)(ws + lay.off_list);
    count = H.list_count;
    const SmemPlan sp = smem_plan(&D, MODE_SAMPLE);
    PipeArgs A{};
    A.recs = (const CellRec*)(ws + lay.off_recs);
    A.qoff = qoff;
    A.tile_cell = (const int64_t*)(ws + lay.off_tilecell);
    A.idx = idx;
    A.tabcum = tabcum;
    A.tab_stride = tab_stride;
    A.use_staging = sp.use_staging;
    A.n_cells = n_active;
    A.n_pairs = n_pairs;
    A.n_photons = n_photons;
    A.n_tiles = lay.ntiles;
    A.ticket = H.ticket;
    A.energies = energies;
    A.off_w = sp.off_w;
    pxb_project_params P0{};
    if (sample_is_common(D, cells)) rc = launch_pipeline<MODE_SAMPLE, 1>(D, *cells, P0, A, sp, st);
    else rc = launch_pipeline<MODE_SAMPLE, 0>(D, *cells, P0, A, sp, st);
    if (rc) return rc;
  }
  k_thermal_sample_cta<<<(unsigned)grid, SAMPLE_THREADS, smem, st>>>(
      D, *cells, n_active, idx, nph, poff, energies, list, (const int64_t*)count);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

// ---- projection ---------------------------------------------------------------------------
extern "C" int pxb_project_workspace_bytes(int64_t n_cells, int64_t n_pairs, int64_t* bytes) {
  PXB_REQUIRE(bytes && n_pairs >= 0 && n_cells >= 0, "bad argument");
  *bytes = pipe_layout(nullptr, n_cells, n_pairs, false, true).total;
  return PXB_OK;
}

static int empty_events(int64_t* n_events_out, double* stats_out, cudaStream_t st) {
  PXB_CUDA(cudaMemsetAsync(n_events_out, 0, sizeof(int64_t), st));
  const double init[3] = {0.0, 1.0e99, -1.0e99};
  PXB_CUDA(cudaMemcpyAsync(stats_out, init, sizeof(init), cudaMemcpyHostToDevice, st));
  PXB_CUDA(cudaStreamSynchronize(st));
  return PXB_OK;
}

static int reduce_stats(char* ws, const PipeLayout& lay, double* stats_out, cudaStream_t st) {
  double* ps = (double*)(ws + lay.off_part);
  double* red = (double*)(ws + lay.off_red);
  const int64_t per_block = (lay.ntiles + STATS_BLOCKS - 1) / STATS_BLOCKS;
  const int nb = (int)((lay.ntiles + per_block - 1) / per_block);
  k_reduce_stats<<<nb, 256, 0, st>>>(ps, ps + lay.ntiles, ps + 2 * lay.ntiles, lay.ntiles, red, per_block);
  PXB_LAUNCH_CHECK();
  k_reduce_stats_final<<<1, 32, 0, st>>>(red, nb, stats_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

static void fill_project_args(PipeArgs& A, char* ws, const PipeLayout& lay, const PipeHeader& H,
                              const pxb_photon_list* L, const pxb_project_params* P, double* xsky,
                              double* ysky, double* eobs, double* los, int64_t* n_events_out) {
  A.recs = (const CellRec*)(ws + lay.off_recs);
  A.geom = (const CellGeom*)(ws + lay.off_geom);
  A.qoff = L->qoff;
  A.tile_cell = (const int64_t*)(ws + lay.off_tilecell);
  A.n_cells = L->n_cells;
  A.n_pairs = L->n_pairs;
  A.n_photons = L->n_photons;
  A.n_tiles = lay.ntiles;
  A.ticket = H.ticket;
  A.xsky = xsky;
  A.ysky = ysky;
  A.eobs = eobs;
  A.los = P->save_los ? los : nullptr;
  A.n_events = n_events_out;
  A.f32 = P->events_f32 != 0;
  // float events carry sky OFFSETS from the centre wherever the sky coordinates are angles about it
  const bool about_centre = P->observer == 0 && P->sky_mode != PXB_SKY_PHYS;
  A.sky0 = about_centre ? P->sky_center[0] : 0.0;
  A.sky1 = about_centre ? P->sky_center[1] : 0.0;
  A.bstatus = (unsigned long long*)(ws + lay.off_bstatus);
  A.part_sum = (double*)(ws + lay.off_part);
  A.part_min = A.part_sum + lay.ntiles;
  A.part_max = A.part_min + lay.ntiles;
}

static int check_project_inputs(const pxb_photon_list* L, const pxb_project_params* P, double* xsky,
                                double* ysky, double* eobs, double* los, bool need_energy) {
  PXB_REQUIRE(xsky && ysky && eobs, "output arrays missing");
  PXB_REQUIRE(!P->save_los || los, "save_los needs the los array");
  PXB_REQUIRE(L->poff && L->qoff && L->x && L->y && L->z, "photon list incomplete");
  PXB_REQUIRE(!need_energy || L->energy, "photon list has no energies");
  PXB_REQUIRE(L->n_pairs > 0 && 2 * L->n_pairs >= L->n_photons, "n_pairs does not match the photon count");
  PXB_REQUIRE(P->no_shifting || (L->vx && L->vy && L->vz), "velocities missing");
  PXB_REQUIRE(P->observer == 1 || P->sky_mode == PXB_SKY_PHYS || P->D_A_kpc > 0.0,
              "D_A must be positive for sky coordinates");
  return PXB_OK;
}

extern "C" int pxb_project(const pxb_photon_list* L, const pxb_project_params* P, double* xsky,
                           double* ysky, double* eobs, double* los, int64_t* n_events_out,
                           double* stats_out, void* workspace, int64_t workspace_bytes,
                           void* stream) {
  PXB_REQUIRE(L && n_events_out && stats_out, "NULL argument");
  int rc = pxb_check_project_params(P);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (L->n_photons == 0 || L->n_cells == 0) return empty_events(n_events_out, stats_out, st);
  rc = check_project_inputs(L, P, xsky, ysky, eobs, los, true);
  if (rc) return rc;
  const PipeLayout lay = pipe_layout(nullptr, L->n_cells, L->n_pairs, false, true);
  PXB_REQUIRE(workspace && workspace_bytes >= lay.total,
              "projection workspace too small (%lld < %lld)", (long long)workspace_bytes,
              (long long)lay.total);
  PXB_REQUIRE(lay.ntiles < (1LL << 32) - (1 << 20), "too many photons for one projection call");
  char* ws = (char*)workspace;
  const PipeHeader H = pipe_header(ws, lay);
  // header + look-back status words are adjacent: one memset
  PXB_CUDA(cudaMemsetAsync(ws + lay.off_hdr, 0, (size_t)(lay.off_recs - lay.off_hdr), st));
  k_prep_project<true><<<(unsigned)((L->n_cells + 255) / 256), 256, 0, st>>>(
      *L, *P, (CellRec*)(ws + lay.off_recs), (CellGeom*)(ws + lay.off_geom),
      (int64_t*)(ws + lay.off_tilecell), lay.tslots);
  PXB_LAUNCH_CHECK();
  const SmemPlan sp = smem_plan(nullptr, MODE_PROJECT);
  PipeArgs A{};
  fill_project_args(A, ws, lay, H, L, P, xsky, ysky, eobs, los, n_events_out);
  A.energies = const_cast<double*>(L->energy);
  A.off_w = sp.off_w;
  static const DevModel D0{};
  static const pxb_thermal_cells C0{};
  const int spec = project_spec(P);
  if (spec == 1) rc = launch_pipeline<MODE_PROJECT, 1>(D0, C0, *P, A, sp, st);
  else if (spec == 2) rc = launch_pipeline<MODE_PROJECT, 2>(D0, C0, *P, A, sp, st);
  else rc = launch_pipeline<MODE_PROJECT, 0>(D0, C0, *P, A, sp, st);
  if (rc) return rc;
  return reduce_stats(ws, lay, stats_out, st);
}

// ---- fused: active cells -> events ---------------------------------------------------------
extern "C" int pxb_make_events_workspace_bytes(const pxb_model* model, int64_t n_cells,
                                               int64_t n_pairs, int64_t* bytes) {
  PXB_REQUIRE(model && bytes && n_pairs >= 0 && n_cells >= 0, "bad argument");
  *bytes = pipe_layout(&model->dev, n_cells, n_pairs, true, true).total;
  return PXB_OK;
}

extern "C" int pxb_make_events(const pxb_model* model, const pxb_thermal_cells* cells,
                               const int64_t* idx, const pxb_photon_list* L,
                               const pxb_project_params* P, double* xsky, double* ysky,
                               double* eobs, double* los, int64_t* n_events_out, double* stats_out,
                               int64_t* n_declined_out, void* workspace, int64_t workspace_bytes,
                               void* stream) {
  PXB_REQUIRE(L && n_events_out && stats_out && n_declined_out && idx, "NULL argument");
  int rc = pxb_check_cells(model, cells);
  if (rc) return rc;
  rc = pxb_check_project_params(P);
  if (rc) return rc;
  const DevModel& D = model->dev;
  PXB_REQUIRE(D.fast_ok, "the fused path needs non-negative spectral tables (use the two-stage path)");
  cudaStream_t st = (cudaStream_t)stream;
  if (L->n_photons == 0 || L->n_cells == 0) {
    PXB_CUDA(cudaMemsetAsync(n_declined_out, 0, sizeof(int64_t), st));
    return empty_events(n_events_out, stats_out, st);
  }
  rc = check_project_inputs(L, P, xsky, ysky, eobs, los, false);
  if (rc) return rc;
  const PipeLayout lay = pipe_layout(&D, L->n_cells, L->n_pairs, true, true);
  PXB_REQUIRE(workspace && workspace_bytes >= lay.total,
              "make_events workspace too small (%lld < %lld)", (long long)workspace_bytes,
              (long long)lay.total);
  PXB_REQUIRE(lay.ntiles < (1LL << 32) - (1 << 20), "too many photons for one call");
  char* ws = (char*)workspace;
  const PipeHeader H = pipe_header(ws, lay);
  PXB_CUDA(cudaMemsetAsync(ws + lay.off_hdr, 0, (size_t)(lay.off_recs - lay.off_hdr), st));
  int tab_stride = 0;
  int64_t* tabcum = nullptr;
  rc = prep_sample(model, cells, L->n_cells, idx, L->poff, L->qoff, ws, lay, &tab_stride, &tabcum, st);
  if (rc) return rc;
  k_prep_project<false><<<(unsigned)((L->n_cells + 255) / 256), 256, 0, st>>>(
      *L, *P, (CellRec*)(ws + lay.off_recs), (CellGeom*)(ws + lay.off_geom),
      (int64_t*)(ws + lay.off_tilecell), lay.tslots);
  PXB_LAUNCH_CHECK();
  // small CTAs with the table rows left to L1/L2 measure faster than one big CTA per SM with
  // staged rows (23.0 vs 24.6 ms on the bench workload); generic_kernels bit 1 selects the latter
  const SmemPlan sp = smem_plan(&D, MODE_FUSED, (cells->generic_kernels & 2) == 0);
  PipeArgs A{};
  fill_project_args(A, ws, lay, H, L, P, xsky, ysky, eobs, los, n_events_out);
  A.idx = idx;
  A.tabcum = tabcum;
  A.tab_stride = tab_stride;
  A.use_staging = sp.use_staging;
  A.off_w = sp.off_w;
  const int pspec = project_spec(P);
  const bool scommon = sample_is_common(D, cells);
  if (scommon && pspec == 1) rc = launch_pipeline<MODE_FUSED, 1>(D, *cells, *P, A, sp, st);
  else if (scommon && pspec == 2) rc = launch_pipeline<MODE_FUSED, 2>(D, *cells, *P, A, sp, st);
  else rc = launch_pipeline<MODE_FUSED, 0>(D, *cells, *P, A, sp, st);
  if (rc) return rc;
  // cells the mixture sampler declined have produced no events: the caller must fall back to the
  // two-stage path when this count is not zero
  PXB_CUDA(cudaMemcpyAsync(n_declined_out, H.list_count, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  return reduce_stats(ws, lay, stats_out, st);
}
