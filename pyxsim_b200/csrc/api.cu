// api.cu -- error plumbing, version and device queries of libpxb.
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[1024] = "";

void pxb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* pxb_last_error(void) { return g_err; }

#include <atomic>
static std::atomic<long long> g_launches{0};
void pxb_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" int64_t pxb_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }
extern "C" void pxb_launch_count_reset(void) { g_launches.store(0, std::memory_order_relaxed); }

extern "C" int pxb_version(void) { return 100; }  // 0.1.0

int pxb_sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

extern "C" int pxb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0, n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    pxb_set_error("no CUDA device visible (%s): libpxb has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return PXB_ERR_NO_DEVICE;
  }
  PXB_CUDA(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  PXB_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  PXB_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  PXB_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = a;
  if (cc_major) *cc_major = b;
  if (cc_minor) *cc_minor = c;
  return PXB_OK;
}

// sizeof() of every ABI struct, so bindings can verify their layout without a GPU
extern "C" int pxb_struct_sizes(int64_t* out, int n) {
  const int64_t v[] = {(int64_t)sizeof(pxb_table_desc),    (int64_t)sizeof(pxb_model_desc),
                       (int64_t)sizeof(pxb_thermal_cells), (int64_t)sizeof(pxb_geometry),
                       (int64_t)sizeof(pxb_powerlaw_cells), (int64_t)sizeof(pxb_line_cells),
                       (int64_t)sizeof(pxb_photon_list),   (int64_t)sizeof(pxb_project_params),
                       (int64_t)sizeof(pxb_band_flux)};
  const int m = (int)(sizeof(v) / sizeof(v[0]));
  if (!out || n < m) {
    pxb_set_error("pxb_struct_sizes needs room for %d entries", m);
    return PXB_ERR_INVALID;
  }
  for (int i = 0; i < m; ++i) out[i] = v[i];
  return m;
}
