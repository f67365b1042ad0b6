// Parts 2 + 3 — box corners -> projection -> 8-corner min/max -> (clamped) 2D box, and the
// projected-box vs 2D-target IoU / GIoU / L1 loss with its analytic backward, one thread per
// box, ONE launch for all boxes of all frames (the reference spends ~100 tiny torch launches
// per task on this: SURVEY.md §2.2).
//
// Reference functions mirrored (operation order follows them, fp32, no FMA contraction in
// the forward so the host can reason about rounding):
//   corners            /root/reference/mmdet3d/core/bbox/structures/lidar_box3d.py:49-89,
//                      cam_box3d.py:116-157 (+ origin re-base :70-73)
//   rotation           structures/utils.py:28-117   limit_period utils.py:10-25
//   LiDAR->CAM         structures/box_3d_mode.py:117-123,162-173
//   points_cam2img     structures/utils.py:175-214
//   variant A          models/dense_heads/centerpoint_head_gga.py:252-275,317-338
//   variant B          datasets/kitti_dataset_GGA_match.py:713-748, clamp :511-512
//   variant C          models/dense_heads/pgd_head.py:413-427
//   IoU/GIoU formula   core/bbox/iou_calculators/iou3d_calculator.py:281-329 (2 axes), mmdet GIoULoss/IoULoss/L1Loss
// Gradient conventions (verified on torch 2.11 CPU): min/max over
// the 8 corners routes to the first index attaining it; elementwise max/min split exact
// ties 1/2-1/2; clamp(min=0) passes the gradient at exactly 0.
#include "common.cuh"

namespace {

constexpr int kBoxThreads = 128;
constexpr int kMaxPartials = 1024;

// Caller-owned scratch of the deterministic loss reduction (include/gga_b200.h,
// gga_loss_scratch_bytes): an arrival counter the kernel leaves at zero, and the per-block
// partial sums the last block adds in index order.  Nothing is shared between calls.
struct LossScratch {
  unsigned int counter;
  unsigned int pad[3];
  float partial[kMaxPartials];
};

int scratch_of(void* scratch, size_t scratch_bytes, float** partial, unsigned int** counter) {
  GGA_REQUIRE(scratch != nullptr, "loss_sum needs scratch (gga_loss_scratch_bytes() bytes, zeroed once)");
  GGA_REQUIRE(scratch_bytes >= sizeof(LossScratch), "scratch too small: %zu bytes given, %zu needed", scratch_bytes,
              sizeof(LossScratch));
  GGA_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15u) == 0, "scratch must be 16-byte aligned");
  LossScratch* s = static_cast<LossScratch*>(scratch);
  *partial = s->partial;
  *counter = &s->counter;
  return GGA_OK;
}

struct Mat4 {
  float m[16];
};

__device__ __forceinline__ Mat4 load_mat(const float* __restrict__ base, int stride, int idx) {
  Mat4 r;
  const float* p = base + (long long)idx * stride;
#pragma unroll
  for (int i = 0; i < 16; ++i) r.m[i] = __ldg(p + i);
  return r;
}

__device__ __forceinline__ float limit_period_dev(float v, float period) {
  // val - floor(val / period + 0.5) * period, fp32 like torch with a python-float period
  return __fsub_rn(v, __fmul_rn(floorf(__fadd_rn(__fdiv_rn(v, period), 0.5f)), period));
}

// row r of M times (X, Y, Z, 1), summed left to right
__device__ __forceinline__ float row_dot(const Mat4& M, int r, float X, float Y, float Z) {
  float a = __fmul_rn(M.m[4 * r + 0], X);
  a = __fadd_rn(a, __fmul_rn(M.m[4 * r + 1], Y));
  a = __fadd_rn(a, __fmul_rn(M.m[4 * r + 2], Z));
  return __fadd_rn(a, M.m[4 * r + 3]);
}

// Geometry of one box after the mode-specific pre-transform: centre, dims, yaw in the frame
// the corners are generated in ("lidar-like": yaw about z, origin (.5,.5,0); "cam-like":
// yaw about y, origin (.5,1,.5)).
struct BoxGeom {
  float c[3];  // corner-frame box origin point (bottom centre)
  float d[3];  // dims along the corner-frame x, y, z
  float s, co;  // sin / cos of the corner-frame yaw
  bool cam;     // cam-like corners
};

__device__ __forceinline__ void corner_norm(int k, bool cam, float& nx, float& ny, float& nz) {
  const int bx = k >> 2, by = (k >> 1) & 1, bz = by ^ (k & 1);
  nx = (float)bx - 0.5f;
  ny = cam ? (float)by - 1.0f : (float)by - 0.5f;
  nz = cam ? (float)bz - 0.5f : (float)bz;
}

__device__ __forceinline__ void corner_xyz(const BoxGeom& g, int k, float& X, float& Y, float& Z) {
  float nx, ny, nz;
  corner_norm(k, g.cam, nx, ny, nz);
  const float lx = __fmul_rn(g.d[0], nx), ly = __fmul_rn(g.d[1], ny), lz = __fmul_rn(g.d[2], nz);
  if (!g.cam) {  // axis 2: x' = x c - y s, y' = x s + y c
    X = __fadd_rn(__fadd_rn(__fmul_rn(lx, g.co), __fmul_rn(ly, -g.s)), g.c[0]);
    Y = __fadd_rn(__fadd_rn(__fmul_rn(lx, g.s), __fmul_rn(ly, g.co)), g.c[1]);
    Z = __fadd_rn(lz, g.c[2]);
  } else {  // axis 1: x' = x c + z s, z' = -x s + z c
    X = __fadd_rn(__fadd_rn(__fmul_rn(lx, g.co), __fmul_rn(lz, g.s)), g.c[0]);
    Y = __fadd_rn(ly, g.c[1]);
    Z = __fadd_rn(__fadd_rn(__fmul_rn(lx, -g.s), __fmul_rn(lz, g.co)), g.c[2]);
  }
}

__device__ __forceinline__ BoxGeom make_geom(const float* b, int mode, const Mat4* rt) {
  BoxGeom g;
  if (mode == GGA_PROJ_LIDAR_DIRECT) {
    g.cam = false;
    g.c[0] = b[0]; g.c[1] = b[1]; g.c[2] = b[2];
    g.d[0] = b[3]; g.d[1] = b[4]; g.d[2] = b[5];
    g.s = sinf(b[6]); g.co = cosf(b[6]);
  } else if (mode == GGA_PROJ_KITTI_CAM) {
    g.cam = true;
    const float TWO_PI = 6.283185307179586f, HALF_PI = 1.5707963267948966f;
    const float yaw1 = limit_period_dev(b[6], TWO_PI);                           // :713
    g.c[0] = row_dot(*rt, 0, b[0], b[1], b[2]);                                  // box_3d_mode.py:162-169
    g.c[1] = row_dot(*rt, 1, b[0], b[1], b[2]);
    g.c[2] = row_dot(*rt, 2, b[0], b[1], b[2]);
    g.d[0] = b[3]; g.d[1] = b[5]; g.d[2] = b[4];                                 // (dx, dz, dy) :120
    const float yc = limit_period_dev(__fsub_rn(-yaw1, HALF_PI), TWO_PI);        // :122-123
    g.s = sinf(yc); g.co = cosf(yc);
  } else {
    g.cam = true;
    g.c[0] = b[0]; g.c[1] = b[1]; g.c[2] = b[2];
    g.d[0] = b[3]; g.d[1] = b[4]; g.d[2] = b[5];
    if (mode == GGA_PROJ_CAM_CENTER) g.c[1] = __fadd_rn(b[1], __fmul_rn(b[4], 0.5f));  // cam_box3d.py:70-73
    g.s = sinf(b[6]); g.co = cosf(b[6]);
  }
  return g;
}

__device__ __forceinline__ float tie_hi(float a, float b) {  // d max(a,b) / d a
  return a > b ? 1.f : (a == b ? 0.5f : 0.f);
}
__device__ __forceinline__ float tie_lo(float a, float b) {  // d min(a,b) / d a
  return a < b ? 1.f : (a == b ? 0.5f : 0.f);
}

// 2D loss between p (pred) and t (target); returns the per-box loss and d loss / d p, d t.
// L1 is handled by the caller (per side).
__device__ __forceinline__ float iou_family_loss(const float p[4], const float t[4], int kind,
                                                 float eps, float gp[4], float gt[4]) {
  const float pw = p[2] - p[0], ph = p[3] - p[1], tw = t[2] - t[0], th = t[3] - t[1];
  const float area1 = pw * ph, area2 = tw * th;
  const float ltx = fmaxf(p[0], t[0]), lty = fmaxf(p[1], t[1]);
  const float rbx = fminf(p[2], t[2]), rby = fminf(p[3], t[3]);
  const float dw = rbx - ltx, dh = rby - lty;
  const float w = fmaxf(dw, 0.f), h = fmaxf(dh, 0.f);
  const float overlap = w * h;
  const float uni = area1 + area2 - overlap;
  const float uc = fmaxf(uni, eps);
  const float iou = overlap / uc;
  float loss, g_iou, g_uc = 0.f, g_ea = 0.f;  // gradients of the loss
  float ew = 0.f, eh = 0.f, edw = 0.f, edh = 0.f, ea = 0.f, eac = 1.f;
  if (kind == GGA_LOSS_GIOU) {
    const float elx = fminf(p[0], t[0]), ely = fminf(p[1], t[1]);
    const float erx = fmaxf(p[2], t[2]), ery = fmaxf(p[3], t[3]);
    edw = erx - elx; edh = ery - ely;
    ew = fmaxf(edw, 0.f); eh = fmaxf(edh, 0.f);
    ea = ew * eh;
    eac = fmaxf(ea, eps);
    const float giou = iou - (eac - uc) / eac;
    loss = 1.f - giou;
    g_iou = -1.f;
    g_uc = -1.f / eac;              // d(-giou)/d uc via +uc/eac
    g_ea = uc / (eac * eac);        // d(-giou)/d eac
  } else {
    const float ic = fmaxf(iou, eps);  // ious.clamp(min=eps)
    const float pass = iou >= eps ? 1.f : 0.f;
    if (kind == GGA_LOSS_IOU_LINEAR) { loss = 1.f - ic; g_iou = -pass; }
    else if (kind == GGA_LOSS_IOU_SQUARE) { loss = 1.f - ic * ic; g_iou = -2.f * ic * pass; }
    else { loss = -logf(ic); g_iou = -pass / ic; }
  }
  // iou = overlap / uc
  const float g_ov_direct = g_iou / uc;
  g_uc += -g_iou * overlap / (uc * uc);
  const float g_uni = g_uc * tie_hi(uni, eps);
  const float g_ov = g_ov_direct - g_uni;  // union = a1 + a2 - overlap
  const float g_a1 = g_uni, g_a2 = g_uni;
  const float g_w = g_ov * h * (dw >= 0.f ? 1.f : 0.f), g_h = g_ov * w * (dh >= 0.f ? 1.f : 0.f);
  // w = rbx - ltx
  gp[0] = -g_w * tie_hi(p[0], t[0]); gt[0] = -g_w * tie_hi(t[0], p[0]);
  gp[1] = -g_h * tie_hi(p[1], t[1]); gt[1] = -g_h * tie_hi(t[1], p[1]);
  gp[2] = g_w * tie_lo(p[2], t[2]);  gt[2] = g_w * tie_lo(t[2], p[2]);
  gp[3] = g_h * tie_lo(p[3], t[3]);  gt[3] = g_h * tie_lo(t[3], p[3]);
  // areas
  gp[0] += -g_a1 * ph; gp[2] += g_a1 * ph; gp[1] += -g_a1 * pw; gp[3] += g_a1 * pw;
  gt[0] += -g_a2 * th; gt[2] += g_a2 * th; gt[1] += -g_a2 * tw; gt[3] += g_a2 * tw;
  if (kind == GGA_LOSS_GIOU) {
    const float g_e = g_ea * tie_hi(ea, eps);
    const float g_ew = g_e * eh * (edw >= 0.f ? 1.f : 0.f), g_eh = g_e * ew * (edh >= 0.f ? 1.f : 0.f);
    gp[0] += -g_ew * tie_lo(p[0], t[0]); gt[0] += -g_ew * tie_lo(t[0], p[0]);
    gp[1] += -g_eh * tie_lo(p[1], t[1]); gt[1] += -g_eh * tie_lo(t[1], p[1]);
    gp[2] += g_ew * tie_hi(p[2], t[2]);  gt[2] += g_ew * tie_hi(t[2], p[2]);
    gp[3] += g_eh * tie_hi(p[3], t[3]);  gt[3] += g_eh * tie_hi(t[3], p[3]);
  }
  return loss;
}

// Accumulates d(total)/d(box params) for one image-plane coordinate of one corner.
//   axis 0: u = q0/d, axis 1: v = q1/d; gval = d(total)/d(that coordinate).
__device__ __forceinline__ void corner_backward(const float* b, const BoxGeom& g, int mode,
                                                const Mat4& P, const Mat4* rt, float depth_clamp,
                                                int k, int axis, float gval, float gb[7]) {
  if (gval == 0.f) return;
  float X, Y, Z;
  corner_xyz(g, k, X, Y, Z);
  const float q0 = row_dot(P, 0, X, Y, Z), q1 = row_dot(P, 1, X, Y, Z), q2 = row_dot(P, 2, X, Y, Z);
  float d = q2, dmask = 1.f;
  if (mode == GGA_PROJ_LIDAR_DIRECT && depth_clamp > 0.f) {  // torch.maximum(depth, 0.1)
    d = fmaxf(q2, depth_clamp);
    dmask = tie_hi(q2, depth_clamp);
  }
  const float qa = axis == 0 ? q0 : q1;
  const float g_qa = gval / d;
  const float g_q2 = -gval * qa / (d * d) * dmask;
  // d q / d (X, Y, Z)
  const float gX = g_qa * P.m[4 * axis + 0] + g_q2 * P.m[8 + 0];
  const float gY = g_qa * P.m[4 * axis + 1] + g_q2 * P.m[8 + 1];
  const float gZ = g_qa * P.m[4 * axis + 2] + g_q2 * P.m[8 + 2];
  float nx, ny, nz;
  corner_norm(k, g.cam, nx, ny, nz);
  const float lx = g.d[0] * nx, ly = g.d[1] * ny, lz = g.d[2] * nz;
  // gradients w.r.t. the corner-frame centre, dims and yaw
  float gc[3] = {gX, gY, gZ}, gd[3], gyaw;
  if (!g.cam) {
    gd[0] = (gX * g.co + gY * g.s) * nx;
    gd[1] = (-gX * g.s + gY * g.co) * ny;
    gd[2] = gZ * nz;
    gyaw = gX * (-lx * g.s - ly * g.co) + gY * (lx * g.co - ly * g.s);
  } else {
    gd[0] = (gX * g.co - gZ * g.s) * nx;
    gd[1] = gY * ny;
    gd[2] = (gX * g.s + gZ * g.co) * nz;
    gyaw = gX * (-lx * g.s + lz * g.co) + gZ * (-lx * g.co - lz * g.s);
  }
  if (mode == GGA_PROJ_LIDAR_DIRECT || mode == GGA_PROJ_CAM_BOTTOM) {
    gb[0] += gc[0]; gb[1] += gc[1]; gb[2] += gc[2];
    gb[3] += gd[0]; gb[4] += gd[1]; gb[5] += gd[2];
    gb[6] += gyaw;
  } else if (mode == GGA_PROJ_CAM_CENTER) {
    gb[0] += gc[0]; gb[1] += gc[1]; gb[2] += gc[2];
    gb[3] += gd[0]; gb[4] += gd[1] + 0.5f * gc[1]; gb[5] += gd[2];
    gb[6] += gyaw;
  } else {  // KITTI_CAM: xyz_cam = rt[:3,:3] xyz + t; dims (dx,dz,dy); yaw_c = -yaw - pi/2 (mod 2pi)
    const Mat4& R = *rt;
    gb[0] += R.m[0] * gc[0] + R.m[4] * gc[1] + R.m[8] * gc[2];
    gb[1] += R.m[1] * gc[0] + R.m[5] * gc[1] + R.m[9] * gc[2];
    gb[2] += R.m[2] * gc[0] + R.m[6] * gc[1] + R.m[10] * gc[2];
    gb[3] += gd[0]; gb[5] += gd[1]; gb[4] += gd[2];
    gb[6] += -gyaw;
  }
  (void)b;
}

struct Args {
  gga_box_loss_args a;
  float* partial;
  unsigned int* counter;
};

__global__ void __launch_bounds__(kBoxThreads) box_loss_kernel(const Args A) {
  const gga_box_loss_args& a = A.a;
  __shared__ float warp_sum[kBoxThreads / 32];
  __shared__ bool is_last;
  float my_sum = 0.f;
  for (int i = blockIdx.x * kBoxThreads + threadIdx.x; i < a.n; i += gridDim.x * kBoxThreads) {
    float b[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) b[j] = __ldg(a.boxes + (long long)i * 7 + j);
    const int fr = a.frame_of_box ? __ldg(a.frame_of_box + i) : i;
    const Mat4 P = load_mat(a.proj, a.proj_stride, fr);
    Mat4 RT;
    if (a.mode == GGA_PROJ_KITTI_CAM) RT = load_mat(a.rt, a.rt_stride, fr);
    const BoxGeom g = make_geom(b, a.mode, &RT);
    // forward: 8 corners, first index wins on ties (torch CPU min/max(dim))
    float xmin = 0.f, ymin = 0.f, xmax = 0.f, ymax = 0.f;
    int ixmin = 0, iymin = 0, ixmax = 0, iymax = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float X, Y, Z;
      corner_xyz(g, k, X, Y, Z);
      const float q0 = row_dot(P, 0, X, Y, Z), q1 = row_dot(P, 1, X, Y, Z), q2 = row_dot(P, 2, X, Y, Z);
      float d = q2;
      if (a.mode == GGA_PROJ_LIDAR_DIRECT && a.depth_clamp > 0.f) d = fmaxf(q2, a.depth_clamp);
      const float u = __fdiv_rn(q0, d), v = __fdiv_rn(q1, d);
      if (k == 0) { xmin = xmax = u; ymin = ymax = v; }
      else {
        if (u < xmin) { xmin = u; ixmin = k; }
        if (u > xmax) { xmax = u; ixmax = k; }
        if (v < ymin) { ymin = v; iymin = k; }
        if (v > ymax) { ymax = v; iymax = k; }
      }
    }
    const float raw[4] = {xmin, ymin, xmax, ymax};
    // validity + clamp (variant B)
    float H = 0.f, W = 0.f;
    const bool have_img = a.img_hw != nullptr;
    if (have_img) {
      const int fi = a.frame_of_box ? fr : 0;
      H = __ldg(a.img_hw + 2 * fi); W = __ldg(a.img_hw + 2 * fi + 1);
    }
    if (a.valid) {
      bool ok = true;
      if (have_img) ok = (xmin < W) & (ymin < H) & (xmax > 0.f) & (ymax > 0.f);
      if (a.pcd_range) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
          ok = ok & (b[j] > __ldg(a.pcd_range + j)) & (b[j] < __ldg(a.pcd_range + 3 + j));
      }
      a.valid[i] = ok ? 1 : 0;
    }
    if (a.box2d) {
      float4 o = make_float4(xmin, ymin, xmax, ymax);
      if (a.clamp_to_image && have_img) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f);
        o.z = fminf(o.z, W); o.w = fminf(o.w, H);
      }
      reinterpret_cast<float4*>(a.box2d)[i] = o;
    }
    if (a.argidx) {
      reinterpret_cast<uchar4*>(a.argidx)[i] = make_uchar4(ixmin, iymin, ixmax, iymax);
    }
    if (a.loss_kind == GGA_LOSS_NONE) continue;

    float t[4], gp[4] = {0.f, 0.f, 0.f, 0.f}, gt[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const float4 tv = __ldg(reinterpret_cast<const float4*>(a.target) + i);
      t[0] = tv.x; t[1] = tv.y; t[2] = tv.z; t[3] = tv.w;
    }
    const float up = a.grad_loss ? __ldg(a.grad_loss + i) : a.grad_scale;
    if (a.loss_kind == GGA_LOSS_L1) {
      float wk[4] = {1.f, 1.f, 1.f, 1.f};
      if (a.weight) {
        if (a.weight_cols == 4) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(a.weight) + i);
          wk[0] = wv.x; wk[1] = wv.y; wk[2] = wv.z; wk[3] = wv.w;
        } else {
          wk[0] = wk[1] = wk[2] = wk[3] = __ldg(a.weight + i);
        }
      }
      float l[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float df = raw[k] - t[k];
        l[k] = fabsf(df);
        const float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
        gp[k] = up * wk[k] * sg;
        gt[k] = -gp[k];
        my_sum += l[k] * wk[k];
      }
      if (a.loss) reinterpret_cast<float4*>(a.loss)[i] = make_float4(l[0], l[1], l[2], l[3]);
    } else {
      float wi = 1.f;
      if (a.weight) {
        if (a.weight_cols == 4) {  // mmdet: weight.mean(-1)
          const float4 wv = __ldg(reinterpret_cast<const float4*>(a.weight) + i);
          wi = (((wv.x + wv.y) + wv.z) + wv.w) / 4.f;
        } else {
          wi = __ldg(a.weight + i);
        }
      }
      const float l = iou_family_loss(raw, t, a.loss_kind, a.eps, gp, gt);
      if (a.loss) a.loss[i] = l;
      my_sum += l * wi;
      const float s = up * wi;
#pragma unroll
      for (int k = 0; k < 4; ++k) { gp[k] *= s; gt[k] *= s; }
    }
    if (a.grad_box2d) reinterpret_cast<float4*>(a.grad_box2d)[i] = make_float4(gp[0], gp[1], gp[2], gp[3]);
    if (a.grad_target) reinterpret_cast<float4*>(a.grad_target)[i] = make_float4(gt[0], gt[1], gt[2], gt[3]);
    if (a.grad_boxes) {
      float gb[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      corner_backward(b, g, a.mode, P, &RT, a.depth_clamp, ixmin, 0, gp[0], gb);
      corner_backward(b, g, a.mode, P, &RT, a.depth_clamp, iymin, 1, gp[1], gb);
      corner_backward(b, g, a.mode, P, &RT, a.depth_clamp, ixmax, 0, gp[2], gb);
      corner_backward(b, g, a.mode, P, &RT, a.depth_clamp, iymax, 1, gp[3], gb);
#pragma unroll
      for (int j = 0; j < 7; ++j) a.grad_boxes[(long long)i * 7 + j] = gb[j];
    }
  }
  if (!a.loss_sum) return;
  // deterministic reduction: per-block partials, the last block adds them in index order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_sum += __shfl_xor_sync(0xffffffffu, my_sum, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = my_sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kBoxThreads / 32; ++w) s += warp_sum[w];
    A.partial[blockIdx.x] = s;
    __threadfence();
    const unsigned int done = atomicAdd(A.counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    float s = 0.f;
    for (unsigned int k = 0; k < gridDim.x; ++k) s += __ldcg(A.partial + k);
    *a.loss_sum = s;
    if (a.loss_accum) atomicAdd(a.loss_accum, s);  // steps on different streams may share the accumulator
    *A.counter = 0u;
  }
}

__global__ void __launch_bounds__(kBoxThreads) box_backward_kernel(
    const float* __restrict__ boxes, const float* __restrict__ proj, int proj_stride,
    const float* __restrict__ rt, int rt_stride, const int32_t* __restrict__ frame_of_box,
    const uint8_t* __restrict__ argidx, const float* __restrict__ grad_box2d, float* grad_boxes,
    int n, int mode, float depth_clamp) {
  const int i = blockIdx.x * kBoxThreads + threadIdx.x;
  if (i >= n) return;
  float b[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) b[j] = __ldg(boxes + (long long)i * 7 + j);
  const int fr = frame_of_box ? __ldg(frame_of_box + i) : i;
  const Mat4 P = load_mat(proj, proj_stride, fr);
  Mat4 RT;
  if (mode == GGA_PROJ_KITTI_CAM) RT = load_mat(rt, rt_stride, fr);
  const BoxGeom g = make_geom(b, mode, &RT);
  const uchar4 ai = reinterpret_cast<const uchar4*>(argidx)[i];
  const float4 gv = reinterpret_cast<const float4*>(grad_box2d)[i];
  float gb[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  corner_backward(b, g, mode, P, &RT, depth_clamp, ai.x, 0, gv.x, gb);
  corner_backward(b, g, mode, P, &RT, depth_clamp, ai.y, 1, gv.y, gb);
  corner_backward(b, g, mode, P, &RT, depth_clamp, ai.z, 0, gv.z, gb);
  corner_backward(b, g, mode, P, &RT, depth_clamp, ai.w, 1, gv.w, gb);
#pragma unroll
  for (int j = 0; j < 7; ++j) grad_boxes[(long long)i * 7 + j] = gb[j];
}

// 2D loss on given boxes
__global__ void __launch_bounds__(kBoxThreads) box2d_loss_kernel(
    const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ weight,
    int weight_cols, const float* __restrict__ grad_loss, int n, int kind, float eps, float grad_scale,
    float* loss, float* loss_sum, float* grad_pred, float* grad_target, float* partial,
    unsigned int* counter) {
  __shared__ float warp_sum[kBoxThreads / 32];
  __shared__ bool is_last;
  float my_sum = 0.f;
  for (int i = blockIdx.x * kBoxThreads + threadIdx.x; i < n; i += gridDim.x * kBoxThreads) {
    const float4 pv = __ldg(reinterpret_cast<const float4*>(pred) + i);
    const float4 tv = __ldg(reinterpret_cast<const float4*>(target) + i);
    const float p[4] = {pv.x, pv.y, pv.z, pv.w}, t[4] = {tv.x, tv.y, tv.z, tv.w};
    float gp[4] = {0.f, 0.f, 0.f, 0.f}, gt[4] = {0.f, 0.f, 0.f, 0.f};
    const float up = grad_loss ? __ldg(grad_loss + i) : grad_scale;
    float wk[4] = {1.f, 1.f, 1.f, 1.f};
    if (weight) {
      if (weight_cols == 4) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(weight) + i);
        wk[0] = wv.x; wk[1] = wv.y; wk[2] = wv.z; wk[3] = wv.w;
      } else {
        wk[0] = wk[1] = wk[2] = wk[3] = __ldg(weight + i);
      }
    }
    if (kind == GGA_LOSS_L1) {
      float l[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float df = p[k] - t[k];
        l[k] = fabsf(df);
        const float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
        gp[k] = up * wk[k] * sg;
        gt[k] = -gp[k];
        my_sum += l[k] * wk[k];
      }
      if (loss) reinterpret_cast<float4*>(loss)[i] = make_float4(l[0], l[1], l[2], l[3]);
    } else {
      const float wi = (weight && weight_cols == 4) ? (((wk[0] + wk[1]) + wk[2]) + wk[3]) / 4.f : wk[0];
      const float l = iou_family_loss(p, t, kind, eps, gp, gt);
      if (loss) loss[i] = l;
      my_sum += l * wi;
      const float s = up * wi;
#pragma unroll
      for (int k = 0; k < 4; ++k) { gp[k] *= s; gt[k] *= s; }
    }
    if (grad_pred) reinterpret_cast<float4*>(grad_pred)[i] = make_float4(gp[0], gp[1], gp[2], gp[3]);
    if (grad_target) reinterpret_cast<float4*>(grad_target)[i] = make_float4(gt[0], gt[1], gt[2], gt[3]);
  }
  if (!loss_sum) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_sum += __shfl_xor_sync(0xffffffffu, my_sum, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = my_sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kBoxThreads / 32; ++w) s += warp_sum[w];
    partial[blockIdx.x] = s;
    __threadfence();
    is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    float s = 0.f;
    for (unsigned int k = 0; k < gridDim.x; ++k) s += __ldcg(partial + k);
    *loss_sum = s;
    *counter = 0u;
  }
}

// Axis-aligned 3-D IoU / GIoU loss of aligned box pairs (x1, y1, z1, x2, y2, z2): the vendored
// formula of /root/reference/mmdet3d/core/bbox/iou_calculators/iou3d_calculator.py:281-329
// (is_aligned) under AxisAlignedIoULoss (mmdet3d/models/losses/axis_aligned_iou_loss.py:10-82):
// loss = 1 - iou (or 1 - giou).  Same tie / clamp gradient conventions as the 2-D family.
__device__ __forceinline__ float aa3d_loss(const float p[6], const float t[6], bool giou, float eps,
                                           float gp[6], float gt[6]) {
  float pe[3], te[3], dw[3], w[3], lt[3], rb[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    pe[k] = p[3 + k] - p[k];
    te[k] = t[3 + k] - t[k];
    lt[k] = fmaxf(p[k], t[k]);
    rb[k] = fminf(p[3 + k], t[3 + k]);
    dw[k] = rb[k] - lt[k];
    w[k] = fmaxf(dw[k], 0.f);
  }
  const float area1 = pe[0] * pe[1] * pe[2], area2 = te[0] * te[1] * te[2];
  const float overlap = w[0] * w[1] * w[2];
  const float uni = area1 + area2 - overlap;
  const float uc = fmaxf(uni, eps);
  const float iou = overlap / uc;
  float loss = 1.f - iou, g_uc = overlap / (uc * uc), g_ea = 0.f;  // d loss / d uc ; g_iou = -1
  float edw[3] = {0.f, 0.f, 0.f}, ew[3] = {0.f, 0.f, 0.f}, ea = 0.f, eac = 1.f;
  if (giou) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      edw[k] = fmaxf(p[3 + k], t[3 + k]) - fminf(p[k], t[k]);
      ew[k] = fmaxf(edw[k], 0.f);
    }
    ea = ew[0] * ew[1] * ew[2];
    eac = fmaxf(ea, eps);
    loss = 1.f - (iou - (eac - uc) / eac);
    g_uc += -1.f / eac;
    g_ea = uc / (eac * eac);
  }
  const float g_uni = g_uc * tie_hi(uni, eps);
  const float g_ov = -1.f / uc - g_uni;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
    const float g_w = g_ov * w[k1] * w[k2] * (dw[k] >= 0.f ? 1.f : 0.f);
    gp[k] = -g_w * tie_hi(p[k], t[k]) - g_uni * pe[k1] * pe[k2];
    gt[k] = -g_w * tie_hi(t[k], p[k]) - g_uni * te[k1] * te[k2];
    gp[3 + k] = g_w * tie_lo(p[3 + k], t[3 + k]) + g_uni * pe[k1] * pe[k2];
    gt[3 + k] = g_w * tie_lo(t[3 + k], p[3 + k]) + g_uni * te[k1] * te[k2];
    if (giou) {
      const float g_ew = g_ea * tie_hi(ea, eps) * ew[k1] * ew[k2] * (edw[k] >= 0.f ? 1.f : 0.f);
      gp[k] += -g_ew * tie_lo(p[k], t[k]);
      gt[k] += -g_ew * tie_lo(t[k], p[k]);
      gp[3 + k] += g_ew * tie_hi(p[3 + k], t[3 + k]);
      gt[3 + k] += g_ew * tie_hi(t[3 + k], p[3 + k]);
    }
  }
  return loss;
}

__global__ void __launch_bounds__(kBoxThreads) box3d_aa_loss_kernel(
    const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ weight,
    const float* __restrict__ grad_loss, int n, int giou, float eps, float grad_scale, float* loss,
    float* loss_sum, float* grad_pred, float* grad_target, float* partial, unsigned int* counter) {
  __shared__ float warp_sum[kBoxThreads / 32];
  __shared__ bool is_last;
  float my_sum = 0.f;
  for (int i = blockIdx.x * kBoxThreads + threadIdx.x; i < n; i += gridDim.x * kBoxThreads) {
    float p[6], t[6], gp[6], gt[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { p[k] = __ldg(pred + (size_t)i * 6 + k); t[k] = __ldg(target + (size_t)i * 6 + k); }
    const float l = aa3d_loss(p, t, giou != 0, eps, gp, gt);
    const float wi = weight ? __ldg(weight + i) : 1.f;
    const float up = (grad_loss ? __ldg(grad_loss + i) : grad_scale) * wi;
    if (loss) loss[i] = l;
    my_sum += l * wi;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      if (grad_pred) grad_pred[(size_t)i * 6 + k] = gp[k] * up;
      if (grad_target) grad_target[(size_t)i * 6 + k] = gt[k] * up;
    }
  }
  if (!loss_sum) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_sum += __shfl_xor_sync(0xffffffffu, my_sum, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = my_sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kBoxThreads / 32; ++w) s += warp_sum[w];
    partial[blockIdx.x] = s;
    __threadfence();
    is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    float s = 0.f;
    for (unsigned int k = 0; k < gridDim.x; ++k) s += __ldcg(partial + k);
    *loss_sum = s;
    *counter = 0u;
  }
}

int blocks_for(int n) {
  int blocks = (n + kBoxThreads - 1) / kBoxThreads;
  if (blocks > kMaxPartials) blocks = kMaxPartials;
  if (blocks < 1) blocks = 1;
  return blocks;
}

}  // namespace

extern "C" size_t gga_loss_scratch_bytes(void) { return sizeof(LossScratch); }

extern "C" int gga_box_project_loss(const gga_box_loss_args* args, void* stream) {
  GGA_REQUIRE(args != nullptr, "null args");
  const gga_box_loss_args& a = *args;
  GGA_REQUIRE(a.n >= 0, "negative n");
  GGA_REQUIRE(a.mode >= GGA_PROJ_LIDAR_DIRECT && a.mode <= GGA_PROJ_CAM_BOTTOM, "bad mode %d", a.mode);
  GGA_REQUIRE(a.loss_kind >= GGA_LOSS_NONE && a.loss_kind <= GGA_LOSS_L1, "bad loss kind %d", a.loss_kind);
  cudaStream_t st = gga_stream(stream);
  if (a.n == 0) {
    if (a.loss_sum) GGA_CHECK_CUDA(cudaMemsetAsync(a.loss_sum, 0, sizeof(float), st));
    return GGA_OK;
  }
  GGA_REQUIRE(a.boxes && a.proj, "null boxes/proj");
  GGA_REQUIRE(a.proj_stride == 0 || a.proj_stride >= 16, "proj_stride must be 0 or >= 16");
  if (a.mode == GGA_PROJ_KITTI_CAM) {
    GGA_REQUIRE(a.rt != nullptr, "KITTI_CAM mode needs rt (rect @ Trv2c)");
    GGA_REQUIRE(a.rt_stride == 0 || a.rt_stride >= 16, "rt_stride must be 0 or >= 16");
  }
  if (a.loss_kind != GGA_LOSS_NONE) GGA_REQUIRE(a.target != nullptr, "loss needs a target");
  if (a.weight) GGA_REQUIRE(a.weight_cols == 1 || a.weight_cols == 4, "weight_cols must be 1 or 4");
  if (a.loss_kind != GGA_LOSS_L1 && a.loss_kind != GGA_LOSS_NONE)
    GGA_REQUIRE(a.eps > 0.f, "eps must be positive");
  Args A;
  A.a = a;
  A.partial = nullptr;
  A.counter = nullptr;
  if (a.loss_sum) {
    const int rc = scratch_of(a.scratch, a.scratch_bytes, &A.partial, &A.counter);
    if (rc != GGA_OK) return rc;
  }
  box_loss_kernel<<<blocks_for(a.n), kBoxThreads, 0, st>>>(A);
  GGA_CHECK_CUDA(cudaGetLastError());
  return GGA_OK;
}

extern "C" int gga_box_project_backward(const float* boxes, const float* proj, int proj_stride,
                                        const float* rt, int rt_stride, const int32_t* frame_of_box,
                                        const uint8_t* argidx, const float* grad_box2d,
                                        float* grad_boxes, int n, int mode, float depth_clamp,
                                        void* stream) {
  GGA_REQUIRE(n >= 0, "negative n");
  if (n == 0) return GGA_OK;
  GGA_REQUIRE(boxes && proj && argidx && grad_box2d && grad_boxes, "null pointer");
  GGA_REQUIRE(mode >= GGA_PROJ_LIDAR_DIRECT && mode <= GGA_PROJ_CAM_BOTTOM, "bad mode %d", mode);
  if (mode == GGA_PROJ_KITTI_CAM) GGA_REQUIRE(rt != nullptr, "KITTI_CAM mode needs rt");
  box_backward_kernel<<<(n + kBoxThreads - 1) / kBoxThreads, kBoxThreads, 0, gga_stream(stream)>>>(
      boxes, proj, proj_stride, rt, rt_stride, frame_of_box, argidx, grad_box2d, grad_boxes, n, mode,
      depth_clamp);
  GGA_CHECK_CUDA(cudaGetLastError());
  return GGA_OK;
}

extern "C" int gga_box2d_loss(const float* pred, const float* target, const float* weight,
                              int weight_cols, const float* grad_loss, int n, int loss_kind, float eps,
                              float grad_scale, float* loss, float* loss_sum, float* grad_pred,
                              float* grad_target, void* scratch, size_t scratch_bytes, void* stream) {
  GGA_REQUIRE(n >= 0, "negative n");
  GGA_REQUIRE(loss_kind >= GGA_LOSS_GIOU && loss_kind <= GGA_LOSS_L1, "bad loss kind %d", loss_kind);
  cudaStream_t st = gga_stream(stream);
  if (n == 0) {
    if (loss_sum) GGA_CHECK_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(float), st));
    return GGA_OK;
  }
  GGA_REQUIRE(pred && target, "null pred/target");
  if (weight) GGA_REQUIRE(weight_cols == 1 || weight_cols == 4, "weight_cols must be 1 or 4");
  float* partial = nullptr;
  unsigned int* counter = nullptr;
  if (loss_sum) {
    const int rc = scratch_of(scratch, scratch_bytes, &partial, &counter);
    if (rc != GGA_OK) return rc;
  }
  box2d_loss_kernel<<<blocks_for(n), kBoxThreads, 0, st>>>(pred, target, weight, weight_cols, grad_loss,
                                                           n, loss_kind, eps, grad_scale, loss,
                                                           loss_sum, grad_pred, grad_target, partial,
                                                           counter);
  GGA_CHECK_CUDA(cudaGetLastError());
  return GGA_OK;
}

extern "C" int gga_box3d_aa_loss(const float* pred, const float* target, const float* weight,
                                 const float* grad_loss, int n, int giou, float eps, float grad_scale,
                                 float* loss, float* loss_sum, float* grad_pred, float* grad_target,
                                 void* scratch, size_t scratch_bytes, void* stream) {
  GGA_REQUIRE(n >= 0, "negative n");
  cudaStream_t st = gga_stream(stream);
  if (n == 0) {
    if (loss_sum) GGA_CHECK_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(float), st));
    return GGA_OK;
  }
  GGA_REQUIRE(pred && target, "null pred/target");
  GGA_REQUIRE(eps > 0.f, "eps must be positive");
  float* partial = nullptr;
  unsigned int* counter = nullptr;
  if (loss_sum) {
    const int rc = scratch_of(scratch, scratch_bytes, &partial, &counter);
    if (rc != GGA_OK) return rc;
  }
  box3d_aa_loss_kernel<<<blocks_for(n), kBoxThreads, 0, st>>>(pred, target, weight, grad_loss, n, giou, eps,
                                                              grad_scale, loss, loss_sum, grad_pred,
                                                              grad_target, partial, counter);
  GGA_CHECK_CUDA(cudaGetLastError());
  return GGA_OK;
}
