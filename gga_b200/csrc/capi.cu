// C-ABI plumbing: version, error string, device info.
#include <stdarg.h>
#include <atomic>
#include <stdio.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void gga_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* gga_last_error(void) { return g_err; }
extern "C" int gga_version(void) { return 100; }

// Read-only device attributes, cached per device.  Filled at most a few times (a race between two
// first callers writes the same values); the flag is published with release / read with acquire so a
// reader never pairs a set flag with an unwritten value.
static int g_sm[64], g_smem[64];
static std::atomic<bool> g_have[64];

static int dev_attr(const int* cache, int fallback) {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return fallback;
  if (!g_have[d].load(std::memory_order_acquire)) {
    int sm = 0, smem = 0;
    if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, d) != cudaSuccess ||
        cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, d) != cudaSuccess || sm <= 0 || smem <= 0)
      return fallback;
    g_sm[d] = sm;
    g_smem[d] = smem;
    g_have[d].store(true, std::memory_order_release);
  }
  return cache[d];
}

int gga_sm_count() { return dev_attr(g_sm, 148); }
int gga_max_smem_optin() { return dev_attr(g_smem, 232448); }

extern "C" int gga_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
  int d = 0;
  GGA_CHECK_CUDA(cudaGetDevice(&d));
  cudaDeviceProp p;
  GGA_CHECK_CUDA(cudaGetDeviceProperties(&p, d));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (total_mem) *total_mem = p.totalGlobalMem;
  return GGA_OK;
}
