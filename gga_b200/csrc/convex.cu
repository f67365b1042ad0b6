// Convex-polygon membership (contract a6 of SURVEY.md §8a; frustum membership of §8f rank 2)
// and FCAF3D face distances (contract a7).
//
//   _points_in_convex_polygon_3d_jit   /root/reference/mmdet3d/core/bbox/box_np_ops.py:641-675
//       ret[i, j] = all_k ( p.x n[j,k,0] + p.y n[j,k,1] + p.z n[j,k,2] + d[j,k] < 0 )
//       evaluated left to right in the promoted dtype (numba: no contraction); used by
//       points_in_rbbox (:353-376; plane coefficients from surface_equ_3d :617-638) and by
//       tools/data_converter/utils_gga.py:88-101 (frustum membership).  All faces are OPEN
//       (the mmcv contract of membership.cu has a closed z slab) and NaN points are "inside"
//       (sign >= 0 is false) — mirrored.
//   FCAF3DHead._get_face_distances     /root/reference/mmdet3d/models/dense_heads/fcaf3d_head.py:495-520
//       + inside = min > 0 (:566-572).
// The plane coefficients are host-side format work (M boxes); the N x M x S tests run here.
#include "../../include/gga_detmath.h"
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// PT: point type, NT: plane type, CT: the type numba promotes the expression to
template <typename PT, typename NT, typename CT>
__global__ void __launch_bounds__(kThreads) convex_kernel(const PT* __restrict__ pts, int stride,
                                                          const NT* __restrict__ normal, const NT* __restrict__ dd,
                                                          const long long* __restrict__ num_surfaces, int N, int M,
                                                          int S, int tile, uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CT* sn = reinterpret_cast<CT*>(smem_raw);  // [tile][S][4] = (n0, n1, n2, d)
  const int i = blockIdx.x * kThreads + threadIdx.x;
  CT px = 0, py = 0, pz = 0;
  if (i < N) {
    px = (CT)pts[(size_t)i * stride];
    py = (CT)pts[(size_t)i * stride + 1];
    pz = (CT)pts[(size_t)i * stride + 2];
  }
  for (int j0 = 0; j0 < M; j0 += tile) {
    const int nt = min(tile, M - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < nt * S; e += kThreads) {
      const size_t g = (size_t)j0 * S + e;
      sn[4 * e] = (CT)normal[3 * g];
      sn[4 * e + 1] = (CT)normal[3 * g + 1];
      sn[4 * e + 2] = (CT)normal[3 * g + 2];
      sn[4 * e + 3] = (CT)dd[g];
    }
    __syncthreads();
    if (i < N) {
      for (int j = 0; j < nt; ++j) {
        const long long ns = num_surfaces ? num_surfaces[j0 + j] : 9999999ll;
        bool inside = true;
        for (int k = 0; k < S; ++k) {
          if ((long long)k > ns) break;  // sic: the reference tests k > num_surfaces[j]
          const CT* q = sn + 4 * (j * S + k);
          const CT sign = add_rn(add_rn(add_rn(mul_rn(px, q[0]), mul_rn(py, q[1])), mul_rn(pz, q[2])), q[3]);
          if (sign >= (CT)0) { inside = false; break; }
        }
        out[(size_t)i * M + j0 + j] = inside ? 1 : 0;
      }
    }
  }
}

template <typename PT, typename NT, typename CT>
int launch_convex(const void* pts, int stride, const void* normal, const void* d, const long long* ns, int N, int M,
                  int S, uint8_t* out, cudaStream_t st) {
  int tile = (int)(40960 / ((size_t)S * 4 * sizeof(CT)));
  if (tile > 32) tile = 32;
  GGA_REQUIRE(tile >= 1, "too many surfaces per polygon (%d)", S);
  const size_t smem = (size_t)tile * S * 4 * sizeof(CT);
  convex_kernel<PT, NT, CT><<<(N + kThreads - 1) / kThreads, kThreads, smem, st>>>(
      static_cast<const PT*>(pts), stride, static_cast<const NT*>(normal), static_cast<const NT*>(d), ns, N, M, S, tile,
      out);
  GGA_CHECK_CUDA(cudaGetLastError());
  return GGA_OK;
}

__global__ void __launch_bounds__(kThreads) face_dist_kernel(const float* __restrict__ pts,
                                                             const float* __restrict__ boxes, int N, int M,
                                                             float* __restrict__ dist, uint8_t* __restrict__ inside) {
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= (long long)N * M) return;
  const int n = (int)(i / M), m = (int)(i - (long long)n * M);
  const float* b = boxes + 7 * m;
  const float bx = __ldg(b), by = __ldg(b + 1), bz = __ldg(b + 2), dx = __ldg(b + 3), dy = __ldg(b + 4),
              dz = __ldg(b + 5), yaw = __ldg(b + 6);
  double sd, cd;
  gga_sincos_f32(-yaw, &sd, &cd);  // rotation_3d_in_axis(shift, -yaw, axis=2)
  const float s = __double2float_rn(sd), c = __double2float_rn(cd);
  const float sx = __fsub_rn(__ldg(pts + 3 * n), bx), sy = __fsub_rn(__ldg(pts + 3 * n + 1), by),
              sz = __fsub_rn(__ldg(pts + 3 * n + 2), bz);
  // einsum('aij,jka->aik') with rot_mat_T = [[c, s, 0], [-s, c, 0], [0, 0, 1]]
  const float rx = __fadd_rn(__fmul_rn(sx, c), __fmul_rn(sy, -s));
  const float ry = __fadd_rn(__fmul_rn(sx, s), __fmul_rn(sy, c));
  const float cx = __fadd_rn(bx, rx), cy = __fadd_rn(by, ry), cz = __fadd_rn(bz, sz);
  const float hx = dx / 2.f, hy = dy / 2.f, hz = dz / 2.f;
  const float f0 = __fadd_rn(__fsub_rn(cx, bx), hx), f1 = __fsub_rn(__fadd_rn(bx, hx), cx);
  const float f2 = __fadd_rn(__fsub_rn(cy, by), hy), f3 = __fsub_rn(__fadd_rn(by, hy), cy);
  const float f4 = __fadd_rn(__fsub_rn(cz, bz), hz), f5 = __fsub_rn(__fadd_rn(bz, hz), cz);
  if (dist) {
    float* o = dist + i * 6;
    o[0] = f0; o[1] = f1; o[2] = f2; o[3] = f3; o[4] = f4; o[5] = f5;
  }
  if (inside) {
    // torch.min propagates NaN: min > 0 is false if any distance is NaN
    const bool ok = (f0 > 0.f) & (f1 > 0.f) & (f2 > 0.f) & (f3 > 0.f) & (f4 > 0.f) & (f5 > 0.f);
    inside[i] = ok ? 1 : 0;
  }
}

}  // namespace

extern "C" int gga_points_in_convex_polygons(const void* points, int pts_stride, int points_f64, const void* normal,
                                             const void* d, int planes_f64, const int64_t* num_surfaces, int N, int M,
                                             int S, uint8_t* out, void* stream) {
  GGA_REQUIRE(N >= 0 && M >= 0 && S >= 0, "negative size");
  if (N == 0 || M == 0) return GGA_OK;
  GGA_REQUIRE(points && out, "null pointer");
  GGA_REQUIRE(S == 0 || (normal && d), "null plane pointer");
  GGA_REQUIRE(pts_stride >= 3, "pts_stride must be >= 3");
  cudaStream_t st = gga_stream(stream);
  const long long* ns = reinterpret_cast<const long long*>(num_surfaces);
  if (S == 0) {  // no surface rejects anything
    GGA_CHECK_CUDA(cudaMemsetAsync(out, 1, (size_t)N * M, st));
    return GGA_OK;
  }
  if (points_f64) {
    GGA_REQUIRE(planes_f64, "float64 points need float64 planes (numpy promotes both)");
    return launch_convex<double, double, double>(points, pts_stride, normal, d, ns, N, M, S, out, st);
  }
  if (planes_f64) return launch_convex<float, double, double>(points, pts_stride, normal, d, ns, N, M, S, out, st);
  return launch_convex<float, float, float>(points, pts_stride, normal, d, ns, N, M, S, out, st);
}

extern "C" int gga_face_distances(const float* points, const float* boxes, int N, int M, float* dist, uint8_t* inside,
                                  void* stream) {
  GGA_REQUIRE(N >= 0 && M >= 0, "negative size");
  if (N == 0 || M == 0) return GGA_OK;
  GGA_REQUIRE(points && boxes && (dist || inside), "null pointer");
  const long long total = (long long)N * M;
  GGA_REQUIRE((total + kThreads - 1) / kThreads < (1ll << 31), "too many (point, box) pairs");
  face_dist_kernel<<<(unsigned)((total + kThreads - 1) / kThreads), kThreads, 0, gga_stream(stream)>>>(
      points, boxes, N, M, dist, inside);
  GGA_CHECK_CUDA(cudaGetLastError());
  return GGA_OK;
}
