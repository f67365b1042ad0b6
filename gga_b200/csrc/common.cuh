// Shared helpers for libgga_b200.so (error reporting across the C ABI, small device utils).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gga_b200.h"

void gga_set_error(const char* fmt, ...);

#define GGA_CHECK_CUDA(expr)                                                                 \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      gga_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return GGA_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

#define GGA_REQUIRE(cond, ...)   \
  do {                           \
    if (!(cond)) {               \
      gga_set_error(__VA_ARGS__); \
      return GGA_ERR_INVALID;    \
    }                            \
  } while (0)

static inline cudaStream_t gga_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int gga_sm_count();              // cached per device
int gga_max_smem_optin();        // cached per device
