// events.cu -- event-list consumers ("next" row f4): the binning the reference's EventList does
// on the host with astropy WCS + np.histogram (pyxsim/event_list.py:254-318 write_fits_image,
// :320-385 write_spectrum, :79-180 write_fits_file), on the device where the events already are.
#include "common.cuh"

constexpr double EV_PI = 3.14159265358979323846;
constexpr double EV_D2R = EV_PI / 180.0;
constexpr double EV_R2D = 180.0 / EV_PI;

// Forward gnomonic (TAN) projection, the closed form of wcs_world2pix for the header the
// reference builds (event_list.py:106-112,281-287: CTYPE RA---TAN/DEC--TAN, CRVAL = sky_center,
// CRPIX = (nx+1)/2, CDELT = (-dtheta, +dtheta), origin 1):
//   D = sin d0 sin d + cos d0 cos d cos(a - a0)
//   x = R2D cos d sin(a - a0) / D,   y = R2D (cos d0 sin d - sin d0 cos d cos(a - a0)) / D
//   pix = crpix + (x, y) / cdelt
struct TanHeader {
  double ra0_rad, sin_d0, cos_d0, crpix, inv_cdelt1, inv_cdelt2;
};

__device__ __forceinline__ bool cel_to_pixel(const TanHeader& H, double ra_deg, double dec_deg,
                                             double& px, double& py) {
  double sd, cd, sa, ca;
  sincos(dec_deg * EV_D2R, &sd, &cd);
  sincos(ra_deg * EV_D2R - H.ra0_rad, &sa, &ca);
  const double D = H.sin_d0 * sd + H.cos_d0 * cd * ca;
  const double x = EV_R2D * cd * sa / D;
  const double y = EV_R2D * (H.cos_d0 * sd - H.sin_d0 * cd * ca) / D;
  px = H.crpix + x * H.inv_cdelt1;
  py = H.crpix + y * H.inv_cdelt2;
  return D > 0.0;   // D <= 0: the point is on the far hemisphere, no TAN image
}

static TanHeader make_header(double ra0_deg, double dec0_deg, double crpix, double cdelt1,
                             double cdelt2) {
  TanHeader H;
  H.ra0_rad = ra0_deg * EV_D2R;
  H.sin_d0 = sin(dec0_deg * EV_D2R);
  H.cos_d0 = cos(dec0_deg * EV_D2R);
  H.crpix = crpix;
  H.inv_cdelt1 = 1.0 / cdelt1;
  H.inv_cdelt2 = 1.0 / cdelt2;
  return H;
}

__global__ void __launch_bounds__(256)
k_cel_to_pixel(int64_t n, const double* __restrict__ ra, const double* __restrict__ dec,
               const TanHeader H, double* __restrict__ px, double* __restrict__ py) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x, y;
  if (!cel_to_pixel(H, ra[i], dec[i], x, y)) {
    x = __longlong_as_double(0x7ff8000000000000LL);
    y = x;
  }
  px[i] = x;
  py[i] = y;
}

// np.histogram2d(xx, yy, bins=[edges, edges]) with edges = 0.5, 1.5, ..., nx + 0.5 for the
// events with emin <= eobs <= emax (event_list.py:277-293).  Bin i holds edges[i] <= v <
// edges[i+1], the last bin its right edge too; v - 0.5 is exact in fp64, so floor() is the
// searchsorted.  image is the TRANSPOSE the reference writes (H.T: row = y pixel).
__global__ void __launch_bounds__(256)
k_event_image(int64_t n, const double* __restrict__ xsky, const double* __restrict__ ysky,
              const double* __restrict__ eobs, double emin, double emax, const TanHeader H, int nx,
              unsigned long long* __restrict__ image) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double e = ld_stream(eobs + i);
    if (!(e >= emin && e <= emax)) continue;
    double px, py;
    if (!cel_to_pixel(H, ld_stream(xsky + i), ld_stream(ysky + i), px, py)) continue;
    const double fx = px - 0.5, fy = py - 0.5, top = (double)nx;
    if (!(fx >= 0.0 && fx <= top && fy >= 0.0 && fy <= top)) continue;
    const int ix = fx == top ? nx - 1 : (int)fx;
    const int iy = fy == top ? nx - 1 : (int)fy;
    PXB_DASSERT(ix >= 0 && ix < nx && iy >= 0 && iy < nx);
    atomicAdd(image + (size_t)iy * nx + ix, 1ull);
  }
}

// np.histogram(eobs, bins=edges) (event_list.py:344-349): nchan bins between edges[0] and
// edges[nchan] (device array, so the very np.linspace values decide), last bin closed.
// Counts go to a per-CTA shared-memory histogram first.
constexpr int SPEC_SMEM_BINS = 8192;
__global__ void __launch_bounds__(256)
k_event_spectrum(int64_t n, const double* __restrict__ eobs, int nchan,
                 const double* __restrict__ edges, unsigned long long* __restrict__ counts) {
  __shared__ unsigned int s_cnt[SPEC_SMEM_BINS];
  const bool priv = nchan <= SPEC_SMEM_BINS;
  if (priv) {
    for (int j = threadIdx.x; j < nchan; j += blockDim.x) s_cnt[j] = 0u;
    __syncthreads();
  }
  const double first = __ldg(edges), last = __ldg(edges + nchan);
  const double norm = (double)nchan / (last - first);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // a CTA flushes long before a 32-bit counter could wrap: at most 2^31 events per CTA pass
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double e = ld_stream(eobs + i);
    if (!(e >= first && e <= last)) continue;
    int j = (int)((e - first) * norm);
    j = min(max(j, 0), nchan - 1);
    // the arithmetic guess can be one off: the stored edges decide (numpy does the same)
    if (e < __ldg(edges + j)) --j;
    else if (j < nchan - 1 && e >= __ldg(edges + j + 1)) ++j;
    PXB_DASSERT(j >= 0 && j < nchan);
    if (priv) atomicAdd(s_cnt + j, 1u);
    else atomicAdd(counts + j, 1ull);
  }
  if (priv) {
    __syncthreads();
    for (int j = threadIdx.x; j < nchan; j += blockDim.x)
      if (s_cnt[j]) atomicAdd(counts + j, (unsigned long long)s_cnt[j]);
  }
}

extern "C" int pxb_cel_to_pixel(int64_t n, const double* ra_deg, const double* dec_deg,
                                double ra0_deg, double dec0_deg, double crpix, double cdelt1,
                                double cdelt2, double* xpix_out, double* ypix_out, void* stream) {
  PXB_REQUIRE(n >= 0 && cdelt1 != 0.0 && cdelt2 != 0.0, "bad argument");
  if (n == 0) return PXB_OK;
  PXB_REQUIRE(ra_deg && dec_deg && xpix_out && ypix_out, "NULL argument");
  k_cel_to_pixel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      n, ra_deg, dec_deg, make_header(ra0_deg, dec0_deg, crpix, cdelt1, cdelt2), xpix_out, ypix_out);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_event_image(int64_t n, const double* xsky, const double* ysky, const double* eobs,
                               double emin, double emax, double ra0_deg, double dec0_deg, int nx,
                               double dtheta_deg, uint64_t* image, void* stream) {
  PXB_REQUIRE(n >= 0 && nx >= 1 && dtheta_deg > 0.0, "bad argument");
  if (n == 0) return PXB_OK;
  PXB_REQUIRE(xsky && ysky && eobs && image, "NULL argument");
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)pxb_sm_count() * 32;
  if (g > cap) g = cap;
  k_event_image<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(
      n, xsky, ysky, eobs, emin, emax,
      make_header(ra0_deg, dec0_deg, 0.5 * (nx + 1), -dtheta_deg, dtheta_deg), nx,
      (unsigned long long*)image);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}

extern "C" int pxb_event_spectrum(int64_t n, const double* eobs, int nchan, const double* edges,
                                  uint64_t* counts, void* stream) {
  PXB_REQUIRE(n >= 0 && nchan >= 1, "bad argument");
  if (n == 0) return PXB_OK;
  PXB_REQUIRE(eobs && edges && counts, "NULL argument");
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)pxb_sm_count() * 8;
  if (g > cap) g = cap;
  k_event_spectrum<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(n, eobs, nchan, edges,
                                                                  (unsigned long long*)counts);
  PXB_LAUNCH_CHECK();
  return PXB_OK;
}
