/*
 * TEST INFRASTRUCTURE — CPU oracle for the membership contract.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product path (gga_b200/) never does.
 *
 * Restates the un-vendored mmcv-full op `points_in_boxes_cpu`
 *   (mmcv/ops/csrc/pytorch/cpu/points_in_boxes.cpp, pinned mmcv-full 1.6.0 by
 *    /root/reference/docker/Dockerfile:4-6; range >=1.5.2,<=1.8.0 at
 *    /root/reference/mmdet3d/__init__.py:21-22),
 * which the reference re-exports at mmdet3d/ops/__init__.py:12-13 and calls through
 * mmdet3d/core/bbox/structures/base_box3d.py:534,566.  Arithmetic per SURVEY.md
 * Appendix A.1.  Pinned against the reference's own golden masks
 * (tests/test_utils/test_box3d.py:1683-1797) in tests/test_oracle_membership.py.
 *
 * Literal on purpose: box-major loop, cos/sin re-evaluated for every (box, point)
 * pair that passes the z test, no hoisting — this is also the timed CPU baseline.
 * Build with -O2 -ffp-contract=off (x86-64 baseline wheels have no FMA contraction).
 */
#include <math.h>
#include <stdint.h>

/* rotate the shift (sx, sy) into the box frame: angle -rz */
static inline void lidar_to_local(float sx, float sy, float rz, float* lx, float* ly) {
  float cosa = (float)cos((double)(-rz)), sina = (float)sin((double)(-rz));
  *lx = sx * cosa + sy * (-sina);
  *ly = sx * sina + sy * cosa;
}

/* pt = (x, y, z); box = (cx, cy, cz_bottom, dx, dy, dz, rz) */
static inline int pt_in_box3d(const float* pt, const float* box) {
  float x = pt[0], y = pt[1], z = pt[2];
  float cx = box[0], cy = box[1], cz = box[2];
  float dx = box[3], dy = box[4], dz = box[5], rz = box[6];
  cz += dz / 2.0; /* double add, rounded back to float: centre of the box */
  if (fabsf(z - cz) > dz / 2.0) return 0; /* closed z slab, compared in double */
  float lx, ly;
  lidar_to_local(x - cx, y - cy, rz, &lx, &ly);
  float in_flag = (lx > -dx / 2.0) & (lx < dx / 2.0) & (ly > -dy / 2.0) & (ly < dy / 2.0);
  return (int)in_flag;
}

/* out[t * M + m], int32, like the op's internal [T, M] buffer (the Python wrapper
 * transposes to [M, T]).  nthreads > 1 splits the box loop with OpenMP for the
 * "all host threads" reference arm of bench.py; results are identical. */
void gga_oracle_points_in_boxes_cpu(const float* boxes, const float* pts, int pts_stride, int T,
                                    int M, int32_t* out, int nthreads) {
  (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
#endif
  for (int t = 0; t < T; ++t) {
    for (int m = 0; m < M; ++m) {
      out[(int64_t)t * M + m] = pt_in_box3d(pts + (int64_t)m * pts_stride, boxes + (int64_t)t * 7);
    }
  }
}

/* `_part` contract: index of the first enclosing box (ascending), else -1. */
void gga_oracle_points_in_boxes_part(const float* boxes, const float* pts, int pts_stride, int T,
                                     int M, int32_t* out) {
  for (int m = 0; m < M; ++m) {
    int32_t idx = -1;
    for (int t = 0; t < T; ++t) {
      if (pt_in_box3d(pts + (int64_t)m * pts_stride, boxes + (int64_t)t * 7)) { idx = t; break; }
    }
    out[m] = idx;
  }
}

/* The per-box rotation terms exactly as the loop above evaluates them (libm). */
void gga_oracle_box_sincos(const float* rz, int T, float* cosa, float* sina) {
  for (int t = 0; t < T; ++t) {
    cosa[t] = (float)cos((double)(-rz[t]));
    sina[t] = (float)sin((double)(-rz[t]));
  }
}
