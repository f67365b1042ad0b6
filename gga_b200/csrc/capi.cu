// C-ABI plumbing: version, error string, device info.
#include <stdarg.h>
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

static int g_sm[64], g_smem[64];
static bool g_have[64];

static void fill_dev() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return;
  if (g_have[d]) return;
  cudaDeviceGetAttribute(&g_sm[d], cudaDevAttrMultiProcessorCount, d);
  cudaDeviceGetAttribute(&g_smem[d], cudaDevAttrMaxSharedMemoryPerBlockOptin, d);
  g_have[d] = true;
}

int gga_sm_count() {
  fill_dev();
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < 64 && g_have[d]) ? g_sm[d] : 148;
}

int gga_max_smem_optin() {
  fill_dev();
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < 64 && g_have[d]) ? g_smem[d] : 232448;
}

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
