// tiles.cuh -- photon-tile helpers shared by the thread-per-photon kernels.
//
// The reference walks photons with a running cursor per cell
// (e.g. pyxsim/lib/sky_functions.pyx:28-34,87-105).  Here a CTA owns a contiguous tile of
// photons; it finds the cells that intersect the tile with two warp-wide 32-ary searches of
// the exclusive offsets, stages those offsets in shared memory, and every photon finds its
// cell with a short shared-memory binary search.
#pragma once
#include "common.cuh"

// largest c in [0, n) with poff[c] <= p (requires poff[0] <= p); whole warp participates
__device__ __forceinline__ int64_t warp_search_le(const int64_t* __restrict__ poff, int64_t n,
                                                  int64_t p, int lane) {
  int64_t lo = 0, hi = n;
  while (hi - lo > 1) {
    const int64_t step = (hi - lo + 31) >> 5;
    const int64_t probe = lo + (int64_t)(lane + 1) * step;
    const bool le = (probe < hi) && (__ldg(poff + probe) <= p);
    const int k = __popc(__ballot_sync(0xffffffffu, le));
    const int64_t nlo = lo + (int64_t)k * step;
    const int64_t nhi = lo + (int64_t)(k + 1) * step;
    lo = nlo;
    hi = nhi < hi ? nhi : hi;
  }
  return lo;
}

// Per-tile map photon -> cell.  TILE = photons per tile; s_off has TILE + 1 entries.
template <int TILE>
struct TileMap {
  int64_t c0;   // first cell intersecting the tile
  int nc;       // number of cells intersecting it (0 => use the global fallback)
  const int64_t* poff;
  int64_t n_cells;

  // all threads of the CTA call this; s_c has 2 entries of scratch
  __device__ void build(const int64_t* __restrict__ poff_, int64_t n_cells_, int64_t p0,
                        int64_t p1, int64_t* s_off, int64_t* s_c) {
    poff = poff_;
    n_cells = n_cells_;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (wid == 0) {
      int64_t c = warp_search_le(poff, n_cells, p0, lane);
      if (lane == 0) s_c[0] = c;
    } else if (wid == 1) {
      int64_t c = warp_search_le(poff, n_cells, p1 - 1, lane);
      if (lane == 0) s_c[1] = c;
    }
    __syncthreads();
    c0 = s_c[0];
    const int64_t span = s_c[1] - c0 + 1;
    nc = (span <= TILE) ? (int)span : 0;
    for (int q = threadIdx.x; q <= nc && nc > 0; q += blockDim.x) s_off[q] = __ldg(poff + c0 + q);
    __syncthreads();
  }

  // cell of photon p (tile-local result is c0 + q)
  __device__ __forceinline__ int64_t cell_of(int64_t p, const int64_t* s_off) const {
    if (nc == 1) return c0;
    if (nc > 0) {
      int lo = 0, hi = nc;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_off[mid] <= p) lo = mid; else hi = mid;
      }
      return c0 + lo;
    }
    // fallback: tile spans more cells than photons (photon list with many empty cells)
    int64_t lo = 0, hi = n_cells;
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(poff + mid) <= p) lo = mid; else hi = mid;
    }
    return lo;
  }
};
