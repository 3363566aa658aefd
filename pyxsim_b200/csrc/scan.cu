// scan.cu -- photon counts -> exclusive offsets, active-cell ranks, active-cell gather.
//
// Replaces the running cursors of the reference's Python loops
// (pyxsim/source_models/thermal_sources/base.py:474-491,517-520) and the per-chunk gather
// of pyxsim/photon_list.py:423-451 with a reduce / scan-of-tiles / scan pass (counts are
// read twice = 16 B/cell, HBM-bound) and one fused gather kernel.
#include "common.cuh"

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

struct I64x2 { int64_t a, b, c; };  // (photons, active cells, photon pairs)

__device__ __forceinline__ I64x2 block_excl_scan(I64x2 v, I64x2& total) {
  __shared__ int64_t sa[SCAN_THREADS / 32], sb[SCAN_THREADS / 32], sc[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int64_t ia = warp_incl_scan_i64(v.a, lane), ib = warp_incl_scan_i64(v.b, lane),
          ic = warp_incl_scan_i64(v.c, lane);
  __syncthreads();
  if (lane == 31) { sa[wid] = ia; sb[wid] = ib; sc[wid] = ic; }
  __syncthreads();
  int64_t ba = 0, bb = 0, bc = 0, ta = 0, tb = 0, tc = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; ++w) {
    if (w < wid) { ba += sa[w]; bb += sb[w]; bc += sc[w]; }
    ta += sa[w]; tb += sb[w]; tc += sc[w];
  }
  total = I64x2{ta, tb, tc};
  return I64x2{ba + ia - v.a, bb + ib - v.b, bc + ic - v.c};
}

__device__ __forceinline__ void load_items(const int64_t* __restrict__ counts, int64_t n,
                                           int64_t first, int64_t (&c)[SCAN_ITEMS]) {
  if (first + SCAN_ITEMS <= n) {
    const longlong2* p = reinterpret_cast<const longlong2*>(counts + first);  // 64 B aligned
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS / 2; ++q) {
      longlong2 v = __ldg(p + q);
      c[2 * q] = v.x; c[2 * q + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) c[q] = (first + q < n) ? counts[first + q] : 0;
  }
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_reduce(const int64_t* __restrict__ counts, int64_t n, int64_t* __restrict__ tile_ph,
              int64_t* __restrict__ tile_act, int64_t* __restrict__ tile_pr) {
  const int64_t first = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int64_t c[SCAN_ITEMS];
  load_items(counts, n, first, c);
  I64x2 v{0, 0, 0};
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; ++q) { v.a += c[q]; v.b += (c[q] > 0); v.c += (c[q] + 1) >> 1; }
  I64x2 tot;
  block_excl_scan(v, tot);
  if (threadIdx.x == 0) { tile_ph[blockIdx.x] = tot.a; tile_act[blockIdx.x] = tot.b; tile_pr[blockIdx.x] = tot.c; }
}

// single CTA: in-place exclusive scan of the per-tile sums; totals[0..1] = grand totals
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tiles(int64_t* __restrict__ tile_ph, int64_t* __restrict__ tile_act,
             int64_t* __restrict__ tile_pr, int64_t ntiles, int64_t* __restrict__ totals) {
  I64x2 carry{0, 0, 0};
  for (int64_t base = 0; base < ntiles; base += SCAN_THREADS) {
    const int64_t t = base + threadIdx.x;
    I64x2 v{0, 0, 0};
    if (t < ntiles) { v.a = tile_ph[t]; v.b = tile_act[t]; v.c = tile_pr[t]; }
    I64x2 tot;
    I64x2 ex = block_excl_scan(v, tot);
    if (t < ntiles) { tile_ph[t] = carry.a + ex.a; tile_act[t] = carry.b + ex.b; tile_pr[t] = carry.c + ex.c; }
    carry.a += tot.a; carry.b += tot.b; carry.c += tot.c;
  }
  if (threadIdx.x == 0) { totals[0] = carry.a; totals[1] = carry.b; totals[2] = carry.c; }
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_final(const int64_t* __restrict__ counts, int64_t n, const int64_t* __restrict__ tile_ph,
             const int64_t* __restrict__ tile_act, const int64_t* __restrict__ tile_pr,
             int64_t* __restrict__ offsets, int64_t* __restrict__ arank,
             int64_t* __restrict__ pairoff, const int64_t* __restrict__ totals) {
  const int64_t first = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int64_t c[SCAN_ITEMS];
  load_items(counts, n, first, c);
  I64x2 v{0, 0, 0};
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; ++q) { v.a += c[q]; v.b += (c[q] > 0); v.c += (c[q] + 1) >> 1; }
  I64x2 tot;
  I64x2 ex = block_excl_scan(v, tot);
  int64_t a = tile_ph[blockIdx.x] + ex.a, b = tile_act[blockIdx.x] + ex.b, p = tile_pr[blockIdx.x] + ex.c;
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; ++q) {
    if (first + q < n) {
      offsets[first + q] = a;
      arank[first + q] = b;
      if (pairoff) pairoff[first + q] = p;
    }
    a += c[q]; b += (c[q] > 0); p += (c[q] + 1) >> 1;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    offsets[n] = totals[0];
    arank[n] = totals[1];
    if (pairoff) pairoff[n] = totals[2];
  }
}

extern "C" int pxb_scan_workspace_bytes(int64_t n, int64_t* bytes) {
  PXB_REQUIRE(bytes && n >= 0, "bad argument");
  const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  *bytes = 3 * (ntiles + 1) * (int64_t)sizeof(int64_t);
  return PXB_OK;
}

extern "C" int pxb_scan_counts(const int64_t* counts, int64_t n, int64_t* offsets, int64_t* arank,
                               int64_t* pairoff, int64_t* totals, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  PXB_REQUIRE(n >= 0 && offsets && arank && totals, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    PXB_CUDA(cudaMemsetAsync(totals, 0, 3 * sizeof(int64_t), st));
    PXB_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int64_t), st));
    PXB_CUDA(cudaMemsetAsync(arank, 0, sizeof(int64_t), st));
    if (pairoff) PXB_CUDA(cudaMemsetAsync(pairoff, 0, sizeof(int64_t), st));
    return PXB_OK;
  }
  PXB_REQUIRE(counts, "counts is NULL");
  PXB_REQUIRE(((uintptr_t)counts & 15) == 0, "counts must be 16-byte aligned");
  int64_t need = 0;
  pxb_scan_workspace_bytes(n, &need);
  PXB_REQUIRE(workspace && workspace_bytes >= need, "scan workspace too small (%lld < %lld)",
              (long long)workspace_bytes, (long long)need);
  const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  PXB_REQUIRE(ntiles < (1LL << 31), "too many cells");
  int64_t* tile_ph = (int64_t*)workspace;
  int64_t* tile_act = tile_ph + ntiles + 1;
  int64_t* tile_pr = tile_act + ntiles + 1;
  k_scan_reduce<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(counts, n, tile_ph, tile_act, tile_pr);
  PXB_LAUNCH_CHECK();
  k_scan_tiles<<<1, SCAN_THREADS, 0, st>>>(tile_ph, tile_act, tile_pr, ntiles, totals);
  PXB_LAUNCH_CHECK();
  k_scan_final<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(counts, n, tile_ph, tile_act, tile_pr,
                                                         offsets, arank, pairoff, totals);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

// ---------------------------------------------------------------------------------------
// active-cell gather: photon_list.py:423-451
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double unwrap(double pos, double dx, double le, double re, double dw,
                                         int periodic) {
  if (periodic) {
    if (pos + dx < le) pos += dw;
    else if (pos - dx > re) pos -= dw;
  }
  return pos;
}

__global__ void __launch_bounds__(256)
k_compact_cells(int64_t n, const int64_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                const int64_t* __restrict__ arank, const int64_t* __restrict__ pairoff,
                int64_t* __restrict__ qoff_out, const double* __restrict__ x,
                const double* __restrict__ y, const double* __restrict__ z,
                const double* __restrict__ vx, const double* __restrict__ vy,
                const double* __restrict__ vz, const double* __restrict__ dx,
                const __grid_constant__ pxb_geometry G, int64_t* __restrict__ idx_out,
                int64_t* __restrict__ nph_out, int64_t* __restrict__ poff_out,
                double* __restrict__ xo, double* __restrict__ yo, double* __restrict__ zo,
                double* __restrict__ vxo, double* __restrict__ vyo, double* __restrict__ vzo,
                double* __restrict__ dxo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    poff_out[arank[n]] = offsets[n];
    if (qoff_out) qoff_out[arank[n]] = pairoff[n];
  }
  if (i >= n) return;
  const int64_t c = ld_stream_i64(counts + i);
  if (c <= 0) return;
  const int64_t k = arank[i];
  PXB_DASSERT(k >= 0 && k < arank[n] && offsets[i] + c <= offsets[n]);
  idx_out[k] = i;
  nph_out[k] = c;
  poff_out[k] = offsets[i];
  if (qoff_out) qoff_out[k] = pairoff[i];
  const double w = dx ? dx[i] : 0.0;
  if (xo) {
    // NB the reference tests the *cell edge* against the object bounds (pos +/- dx)
    xo[k] = unwrap(x[i], w, G.le[0], G.re[0], G.dw[0], G.periodic[0]) - G.c[0];
    yo[k] = unwrap(y[i], w, G.le[1], G.re[1], G.dw[1], G.periodic[1]) - G.c[1];
    zo[k] = unwrap(z[i], w, G.le[2], G.re[2], G.dw[2], G.periodic[2]) - G.c[2];
  }
  if (vxo) {
    vxo[k] = vx[i] - G.bulk_v[0];
    vyo[k] = vy[i] - G.bulk_v[1];
    vzo[k] = vz[i] - G.bulk_v[2];
  }
  if (dxo) dxo[k] = w;
}

extern "C" int pxb_compact_cells(int64_t n, const int64_t* counts, const int64_t* offsets,
                                 const int64_t* arank, const int64_t* pairoff, const double* x,
                                 const double* y, const double* z, const double* vx,
                                 const double* vy, const double* vz, const double* dx,
                                 const pxb_geometry* geom, int64_t* idx_out, int64_t* nph_out,
                                 int64_t* poff_out, int64_t* qoff_out, double* xo, double* yo,
                                 double* zo, double* vxo, double* vyo, double* vzo, double* dxo,
                                 void* stream) {
  PXB_REQUIRE(!qoff_out || pairoff, "qoff_out needs pairoff");
  PXB_REQUIRE(n >= 0 && offsets && arank && idx_out && nph_out && poff_out && geom, "bad argument");
  PXB_REQUIRE(!xo || (x && y && z && yo && zo), "position arrays incomplete");
  PXB_REQUIRE(!vxo || (vx && vy && vz && vyo && vzo), "velocity arrays incomplete");
  const int64_t blocks = n == 0 ? 1 : (n + 255) / 256;
  PXB_REQUIRE(n == 0 || counts, "counts is NULL");
  k_compact_cells<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      n, counts, offsets, arank, pairoff, qoff_out, x, y, z, vx, vy, vz, dx, *geom, idx_out, nph_out,
      poff_out, xo, yo, zo, vxo, vyo, vzo, dxo);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

// ---------------------------------------------------------------------------------------
// compute_radius: pyxsim/source_models/sources.py:77-87
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_radius2(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
          const double* __restrict__ z, const __grid_constant__ pxb_geometry G,
          double* __restrict__ r2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // The reference shifts ALL THREE coordinates of a cell when any periodic axis is out of
  // bounds along axis a, each by dw[a] (pos[:, tfl] += dw[i]); reproduced literally.
  double p[3] = {x[i], y[i], z[i]};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (G.periodic[a]) {
      const bool tfl = p[a] < G.le[a];
      const bool tfr = p[a] > G.re[a];
      if (tfl) { p[0] += G.dw[a]; p[1] += G.dw[a]; p[2] += G.dw[a]; }
      if (tfr) { p[0] -= G.dw[a]; p[1] -= G.dw[a]; p[2] -= G.dw[a]; }
    }
  }
  const double cm_per_kpc = 3.0856775809623245e21;
  const double ax = p[0] - G.c[0], ay = p[1] - G.c[1], az = p[2] - G.c[2];
  r2[i] = (ax * ax + ay * ay + az * az) * (cm_per_kpc * cm_per_kpc);
}

extern "C" int pxb_radius2(int64_t n, const double* x, const double* y, const double* z,
                           const pxb_geometry* geom, double* r2_out, void* stream) {
  PXB_REQUIRE(n >= 0 && geom, "bad argument");
  if (n == 0) return PXB_OK;
  PXB_REQUIRE(x && y && z && r2_out, "NULL array");
  k_radius2<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, x, y, z, *geom, r2_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}
