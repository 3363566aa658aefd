// common.cuh -- shared device/host helpers for libpxb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/pxb.h"

// Index / capacity assertions of the kernels: compiled in by `python -m pyxsim_b200.build
// --debug-checks` (libpxb_dbg.so, loaded with PXB_LIBRARY=...), nothing otherwise.  A failed
// assertion traps the kernel and surfaces as cudaErrorAssert at the next synchronisation.
#ifdef PXB_DEBUG_CHECKS
#include <assert.h>
#define PXB_DASSERT(c) assert(c)
#else
#define PXB_DASSERT(c) ((void)0)
#endif

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpxb is written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------
void pxb_set_error(const char* fmt, ...);

#define PXB_CUDA(call)                                                                 \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess) {                                                           \
      pxb_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,               \
                    cudaGetErrorString(_e));                                           \
      return PXB_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

#define PXB_REQUIRE(cond, ...)                                                         \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      pxb_set_error(__VA_ARGS__);                                                      \
      return PXB_ERR_INVALID;                                                          \
    }                                                                                  \
  } while (0)

void pxb_count_launch();   // api.cu: process-wide kernel launch counter (pxb_launch_count)

#define PXB_LAUNCH_CHECK()                                                             \
  do {                                                                                 \
    pxb_count_launch();                                                                \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      pxb_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,           \
                    cudaGetErrorString(_e));                                           \
      return PXB_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

int pxb_sm_count();  // cached SM count of the current device (148 on B200)

// global id of cell i of a chunk (pxb_thermal_cells.cell_id0 / id_block / id_stride and the same
// triple of the power-law and line structs): contiguous, or block-cyclic shard of a larger dataset
template <class Cells>
__device__ __forceinline__ uint64_t cell_gid(const Cells& C, int64_t i) {
  if (C.id_block > 0) return (uint64_t)(C.cell_id0 + (i / C.id_block) * C.id_stride + i % C.id_block);
  return (uint64_t)(C.cell_id0 + i);
}

// ---------------------------------------------------------------------------------------
// RNG: Philox4x32-10 (Salmon et al. 2011), counter-based.
//   key     = 64-bit user seed
//   counter = (draw_lo, draw_hi, cell_lo, cell_hi16<<16 | stream)
// so every (cell, draw, stream) triple owns an independent 128-bit block: results do not
// depend on launch geometry, chunking or GPU count.
// ---------------------------------------------------------------------------------------
enum PxbStream : uint32_t {
  STREAM_POISSON = 1,
  STREAM_ENERGY = 2,
  STREAM_PROJ_A = 3,   // absorption + first scatter draws
  STREAM_PROJ_B = 4,   // remaining scatter draws
  STREAM_PROJ_C = 5,   // sigma_pos smoothing
  STREAM_PROJ_D = 6,   // internal absorption
  STREAM_POWERLAW = 7,
  STREAM_LINE = 8,
  STREAM_SPLIT = 9,    // multinomial split of a cell's photons over its table rows
  STREAM_PROJ_R = 10,  // low word of an absorption uniform whose top word does not decide
};

struct Philox4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                 uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// The ten round keys of a call depend on the seed only.  Kernels that draw per photon receive
// them precomputed as a __grid_constant__ parameter: the round XORs then take their key straight
// from the constant bank and the twenty key-schedule additions per call disappear.  Same
// numbers as philox4x32_10 above.
struct PhiloxKeys {
  uint32_t k[20];   // k[2r] = seed_lo + r*W0, k[2r+1] = seed_hi + r*W1
};
static inline PhiloxKeys pxb_make_keys(uint64_t seed) {
  PhiloxKeys K;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    K.k[2 * r] = k0;
    K.k[2 * r + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return K;
}
__device__ __forceinline__ Philox4 pxb_rand4(const PhiloxKeys& K, uint32_t stream, uint64_t cell,
                                             uint64_t draw) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t c0 = (uint32_t)draw, c1 = (uint32_t)(draw >> 32), c2 = (uint32_t)cell,
           c3 = (((uint32_t)(cell >> 32)) << 16) | (stream & 0xFFFFu);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ K.k[2 * r];
    const uint32_t n2 = hi0 ^ c3 ^ K.k[2 * r + 1];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ Philox4 pxb_rand4(uint64_t seed, uint32_t stream, uint64_t cell,
                                             uint64_t draw) {
  return philox4x32_10((uint32_t)draw, (uint32_t)(draw >> 32), (uint32_t)cell,
                       (((uint32_t)(cell >> 32)) << 16) | (stream & 0xFFFFu),
                       (uint32_t)seed, (uint32_t)(seed >> 32));
}

// uniform in [0,1) from 52 random bits placed in the mantissa of a double in [1,2)
// (cheaper than an int64 -> double conversion; granularity 2^-52)
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  return __hiloint2double(0x3FF00000 | (hi >> 12), (hi << 20) | (lo >> 12)) - 1.0;
}
// uniform in (0,1]  (safe for log)
__device__ __forceinline__ double u53_open0(uint32_t hi, uint32_t lo) {
  return 2.0 - __hiloint2double(0x3FF00000 | (hi >> 12), (hi << 20) | (lo >> 12));
}

// two standard normals from one Philox block (Box-Muller)
__device__ __forceinline__ void normal2(const Philox4& r, double& n0, double& n1) {
  double u1 = u53_open0(r.x, r.y);
  double u2 = u53(r.z, r.w);
  double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  n0 = rad * c;
  n1 = rad * s;
}

// Sequential uniform stream for one cell (Poisson sampler): draw d -> two doubles per block.
struct CellStream {
  uint64_t seed, cell;
  uint32_t stream;
  uint64_t block;
  int have;
  double spare;
  __device__ CellStream(uint64_t s, uint32_t st, uint64_t c)
      : seed(s), cell(c), stream(st), block(0), have(0), spare(0.0) {}
  __device__ double next() {
    if (have) {
      have = 0;
      return spare;
    }
    Philox4 r = pxb_rand4(seed, stream, cell, block++);
    spare = u53(r.z, r.w);
    have = 1;
    return u53(r.x, r.y);
  }
};

// log(k!) for an integer-valued k >= 0 (the only way the samplers below use lgamma): exact
// table below 16, Stirling's series above (next term < 7e-15 absolute at k = 16) -- one log and
// one division instead of the general-argument lgamma.
__device__ const double g_logfact[16] = {
    0.0, 0.0, 0.69314718055994530942, 1.79175946922805500081, 3.17805383034794561965,
    4.78749174278204599425, 6.57925121201010099506, 8.52516136106541430017,
    10.60460290274525022842, 12.80182748008146961121, 15.10441257307551529523,
    17.50230784587388583929, 19.98721449566188614952, 22.55216385312342288557,
    25.19122118273868150009, 27.89927138384089156609};
__device__ __forceinline__ double log_factorial(double k) {
  if (k < 16.0) return g_logfact[(int)k];
  const double x = k + 1.0;
  const double r = 1.0 / x, r2 = r * r;
  const double series = r * (1.0 / 12.0 - r2 * (1.0 / 360.0 - r2 * (1.0 / 1260.0 - r2 * (1.0 / 1680.0))));
  return (k + 0.5) * log(x) - x + 0.91893853320467274178 + series;
}

// Poisson(lam): multiplication method below 10 (Knuth), transformed rejection PTRS above
// (Hoermann 1993, "The transformed rejection method for generating Poisson random variables").
__device__ inline int64_t poisson_draw(double lam, CellStream& rs) {
  if (!(lam > 0.0) || !isfinite(lam)) return 0;
  if (lam < 10.0) {
    double enlam = exp(-lam);
    int64_t k = 0;
    double prod = rs.next();
    while (prod > enlam) {
      ++k;
      prod *= rs.next();
    }
    return k;
  }
  double slam = sqrt(lam);
  double loglam = log(lam);
  double b = 0.931 + 2.53 * slam;
  double a = -0.059 + 0.02483 * b;
  double invalpha = 1.1239 + 1.1328 / (b - 3.4);
  double vr = 0.9277 - 3.6224 / (b - 2.0);
  for (int it = 0; it < 1000; ++it) {
    double U = rs.next() - 0.5;
    double V = rs.next();
    double us = 0.5 - fabs(U);
    double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return (int64_t)kf;
    if (kf < 0.0 || (us < 0.013 && V > us)) continue;
    if (log(V) + log(invalpha) - log(a / (us * us) + b) <=
        -lam + kf * loglam - log_factorial(kf))
      return (int64_t)kf;
  }
  return (int64_t)floor(lam + 0.5);  // unreachable in practice
}

// Binomial(n, p): inversion (BINV, Kachitvichyanukul & Schmeiser 1988) for n*min(p,1-p) < 10,
// transformed rejection BTRS (Hoermann 1993, "The generation of binomial random variates")
// above.  Exact up to floating point; validated against scipy.stats.binom in the tests.
__device__ inline int64_t binomial_draw(int64_t n, double p, CellStream& rs) {
  if (n <= 0 || !(p > 0.0)) return 0;
  if (p >= 1.0) return n;
  const bool flip = p > 0.5;
  const double pp = flip ? 1.0 - p : p;
  const double nd = (double)n;
  int64_t k = 0;
  if (nd * pp < 10.0) {
    const double q = 1.0 - pp, sr = pp / q, a = (nd + 1.0) * sr;
    const double r0 = exp(nd * log1p(-pp));
    for (int attempt = 0; attempt < 64; ++attempt) {
      double u = rs.next(), r = r0;
      int64_t x = 0;
      while (u > r && x <= n) {
        u -= r;
        ++x;
        r *= (a / (double)x - sr);
      }
      if (x <= n) { k = x; break; }
    }
  } else {
    const double spq = sqrt(nd * pp * (1.0 - pp));
    const double b = 1.15 + 2.53 * spq;
    const double a = -0.0873 + 0.0248 * b + 0.01 * pp;
    const double c = nd * pp + 0.5;
    const double vr = 0.92 - 4.2 / b;
    const double alpha = (2.83 + 5.1 / b) * spq;
    const double lpq = log(pp / (1.0 - pp));
    const double m = floor((nd + 1.0) * pp);
    const double h = log_factorial(m) + log_factorial(nd - m);
    k = (int64_t)floor(nd * pp + 0.5);
    for (int it = 0; it < 1000; ++it) {
      const double u = rs.next() - 0.5;
      double v = rs.next();
      const double us = 0.5 - fabs(u);
      const double kf = floor((2.0 * a / us + b) * u + c);
      if (us >= 0.07 && v <= vr) { k = (int64_t)kf; break; }
      if (kf < 0.0 || kf > nd) continue;
      v = log(v * alpha / (a / (us * us) + b));
      if (v <= h - log_factorial(kf) - log_factorial(nd - kf) + (kf - m) * lpq) { k = (int64_t)kf; break; }
    }
  }
  return flip ? n - k : k;
}

// ---------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------
// np.digitize(x, nodes) - 1 clamped to [0, n-2]  (spectral_models.py:35-37): largest i with
// nodes[i] <= x, clamped.
__device__ __forceinline__ int node_index(const double* __restrict__ nodes, int n, double x) {
  int lo = 0, hi = n;  // count of nodes <= x
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(nodes + mid) <= x) lo = mid + 1; else hi = mid;
  }
  int i = lo - 1;
  i = max(0, min(i, n - 2));
  return i;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Warp-wide min / max of doubles with the integer reduction unit: an order-preserving map to
// unsigned 64-bit keys, then two 32-bit REDUX (high words, then low words among the lanes that
// hold the winning high word) -- ~10 instructions instead of five shuffle + fmin rounds.
__device__ __forceinline__ unsigned long long pxb_ordered_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
}
__device__ __forceinline__ double pxb_from_ordered_key(unsigned long long k) {
  return __longlong_as_double((long long)(k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull)));
}
__device__ __forceinline__ double warp_min(double v) {
  const unsigned long long k = pxb_ordered_key(v);
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
  return pxb_from_ordered_key(((unsigned long long)mh << 32) | ml);
}
__device__ __forceinline__ double warp_max(double v) {
  const unsigned long long k = pxb_ordered_key(v);
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return pxb_from_ordered_key(((unsigned long long)mh << 32) | ml);
}
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ int64_t warp_incl_scan_i64(int64_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int64_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// streaming (read-once / write-once) global accesses: keep L1 for the tables
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int64_t ld_stream_i64(const int64_t* p) {
  int64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(double* p, double v) {
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v));
}
