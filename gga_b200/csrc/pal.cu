// Point-to-Box Alignment distances with their Jacobian, one launch for
// all objects of all frames.
//
// Mirrors CenterHead_GGA.get_distance_single / get_distance_bev
// (/root/reference/mmdet3d/models/dense_heads/centerpoint_head_gga.py:184-248): the reference
// loops in Python over <= 500 objects x 3 tasks x batch with ~20 torch launches each, on ragged
// per-object point lists moved to the device one at a time (:469-470).  Here the lists are one
// CSR array (offsets + packed xy) and one warp serves one object:
//   rotate point and box centre clockwise by rot (structures/utils.py:28-117, 2-D branch):
//       x' = x cos + y sin,  y' = -x sin + y cos                      (products rounded, then added)
//   u = x' - cx', v = y' - cy', hl = w / 2, hh = h / 2
//   min_dis = sum_p min(|u + hl|, |u - hl|, |v + hh|, |v - hh|)       (:209-227, first index on ties)
//   x_dis   = sum_p relu(|u| - 2 hl),   y_dis = sum_p relu(|v| - 2 hh) (:215-219, 228-229)
// and d(min_dis, x_dis, y_dis) / d(cx, cy, w, h, rot) with torch's conventions
// (abs'(0) = 0, relu'(0) = 0, min(dim) routes to the first minimum).
#include "../../include/gga_detmath.h"
#include "common.cuh"

namespace {

constexpr int kPalThreads = 256;

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

constexpr int kChunk = 512;  // points per warp
constexpr int kVals = 18;    // 3 distances + 3 x 5 Jacobian entries

// One warp per (object, 512-point chunk): partial sums to part[obj][chunk][18] (objects with
// thousands of points no longer serialise on one warp; summation order stays fixed).
__global__ void __launch_bounds__(kPalThreads) pal_kernel(const float2* __restrict__ pts,
                                                          const int32_t* __restrict__ offsets,
                                                          const float* __restrict__ box_bev, int n_obj,
                                                          int max_chunks, float* __restrict__ part) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (kPalThreads / 32) + (threadIdx.x >> 5);
  const int obj = (int)(wid / max_chunks), chunk = (int)(wid - (long long)obj * max_chunks);
  if (obj >= n_obj) return;
  const int p0 = __ldg(offsets + obj) + chunk * kChunk, pend = __ldg(offsets + obj + 1);
  if (p0 >= pend && chunk > 0) return;  // chunk 0 always writes (objects without points give zeros)
  const int p1 = min(pend, p0 + kChunk);
  const float cx = __ldg(box_bev + 5 * obj), cy = __ldg(box_bev + 5 * obj + 1), w = __ldg(box_bev + 5 * obj + 2),
              h = __ldg(box_bev + 5 * obj + 3), rot = __ldg(box_bev + 5 * obj + 4);
  double sd, cd;
  gga_sincos_f32(rot, &sd, &cd);
  const float s = __double2float_rn(sd), c = __double2float_rn(cd);
  const float cxr = __fadd_rn(__fmul_rn(cx, c), __fmul_rn(cy, s));
  const float cyr = __fadd_rn(__fmul_rn(cx, -s), __fmul_rn(cy, c));
  const float hl = w / 2.0f, hh = h / 2.0f;
  float a_min = 0.f, a_x = 0.f, a_y = 0.f;
  float jm[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, jx[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, jy[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int i = p0 + lane; i < p1; i += 32) {
    const float2 q = __ldg(pts + i);
    const float px = __fadd_rn(__fmul_rn(q.x, c), __fmul_rn(q.y, s));
    const float py = __fadd_rn(__fmul_rn(q.x, -s), __fmul_rn(q.y, c));
    const float u = __fsub_rn(px, cxr), v = __fsub_rn(py, cyr);
    // du = (-c, -s, 0, 0, v), dv = (s, -c, 0, 0, -u) w.r.t. (cx, cy, w, h, rot)
    const float d0 = __fsub_rn(px, __fsub_rn(cxr, hl)), d1 = __fsub_rn(px, __fadd_rn(cxr, hl));
    const float d2 = __fsub_rn(py, __fsub_rn(cyr, hh)), d3 = __fsub_rn(py, __fadd_rn(cyr, hh));
    float m = fabsf(d0), g = sgn(d0);
    int k = 0;
    if (fabsf(d1) < m) { m = fabsf(d1); g = sgn(d1); k = 1; }
    if (fabsf(d2) < m) { m = fabsf(d2); g = sgn(d2); k = 2; }
    if (fabsf(d3) < m) { m = fabsf(d3); g = sgn(d3); k = 3; }
    a_min += m;
    if (k < 2) {
      jm[0] += g * -c; jm[1] += g * -s; jm[4] += g * v;
      jm[2] += g * (k == 0 ? 0.5f : -0.5f);
    } else {
      jm[0] += g * s; jm[1] += g * -c; jm[4] += g * -u;
      jm[3] += g * (k == 2 ? 0.5f : -0.5f);
    }
    const float ex = __fsub_rn(fabsf(u), __fmul_rn(2.f, hl)), ey = __fsub_rn(fabsf(v), __fmul_rn(2.f, hh));
    if (ex > 0.f) {
      const float gu = sgn(u);
      a_x += ex;
      jx[0] += gu * -c; jx[1] += gu * -s; jx[4] += gu * v; jx[2] += -1.f;
    }
    if (ey > 0.f) {
      const float gv = sgn(v);
      a_y += ey;
      jy[0] += gv * s; jy[1] += gv * -c; jy[4] += gv * -u; jy[3] += -1.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a_min += __shfl_xor_sync(0xffffffffu, a_min, o);
    a_x += __shfl_xor_sync(0xffffffffu, a_x, o);
    a_y += __shfl_xor_sync(0xffffffffu, a_y, o);
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      jm[j] += __shfl_xor_sync(0xffffffffu, jm[j], o);
      jx[j] += __shfl_xor_sync(0xffffffffu, jx[j], o);
      jy[j] += __shfl_xor_sync(0xffffffffu, jy[j], o);
    }
  }
  if (lane == 0) {
    float* o = part + ((size_t)obj * max_chunks + chunk) * kVals;
    o[0] = a_min; o[1] = a_x; o[2] = a_y;
#pragma unroll
    for (int j = 0; j < 5; ++j) { o[3 + j] = jm[j]; o[8 + j] = jx[j]; o[13 + j] = jy[j]; }
  }
}

// Sums the chunk partials of every object in chunk order: thread = (object, value).
__global__ void __launch_bounds__(256) pal_reduce_kernel(const float* __restrict__ part,
                                                         const int32_t* __restrict__ offsets, int n_obj,
                                                         int max_chunks, float* __restrict__ dist,
                                                         float* __restrict__ jac) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int obj = i / kVals, v = i - obj * kVals;
  if (obj >= n_obj) return;
  const int n = __ldg(offsets + obj + 1) - __ldg(offsets + obj);
  const int chunks = n <= 0 ? 1 : (n + kChunk - 1) / kChunk;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += part[((size_t)obj * max_chunks + c) * kVals + v];
  if (v < 3) dist[3 * obj + v] = s;
  else if (jac) jac[15 * obj + (v - 3)] = s;
}

}  // namespace

extern "C" size_t gga_pal_workspace_bytes(int n_obj, int max_points_per_object) {
  if (n_obj <= 0) return 256;
  const int mc = max_points_per_object <= 0 ? 1 : (max_points_per_object + kChunk - 1) / kChunk;
  return (size_t)n_obj * mc * kVals * sizeof(float);
}

extern "C" int gga_point_box_alignment(const float* points_xy, const int32_t* offsets, const float* box_bev,
                                       int n_obj, int max_points_per_object, float* dist, float* jac,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  GGA_REQUIRE(n_obj >= 0, "negative n_obj");
  if (n_obj == 0) return GGA_OK;
  GGA_REQUIRE(offsets && box_bev && dist, "null pointer");
  GGA_REQUIRE((reinterpret_cast<uintptr_t>(points_xy) & 7u) == 0, "points_xy must be 8-byte aligned");
  const int mc = max_points_per_object <= 0 ? 1 : (max_points_per_object + kChunk - 1) / kChunk;
  GGA_REQUIRE(workspace != nullptr && workspace_bytes >= gga_pal_workspace_bytes(n_obj, max_points_per_object),
              "workspace missing or too small (gga_pal_workspace_bytes)");
  const long long warps = (long long)n_obj * mc;
  const int per = kPalThreads / 32;
  GGA_REQUIRE((warps + per - 1) / per < (1ll << 31), "too many (object, chunk) pairs");
  cudaStream_t st = gga_stream(stream);
  float* part = static_cast<float*>(workspace);
  pal_kernel<<<(unsigned)((warps + per - 1) / per), kPalThreads, 0, st>>>(
      reinterpret_cast<const float2*>(points_xy), offsets, box_bev, n_obj, mc, part);
  GGA_CHECK_CUDA(cudaGetLastError());
  pal_reduce_kernel<<<(n_obj * kVals + 255) / 256, 256, 0, st>>>(part, offsets, n_obj, mc, dist, jac);
  GGA_CHECK_CUDA(cudaGetLastError());
  return GGA_OK;
}
