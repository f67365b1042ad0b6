// GGA training-target packing, one launch for all frames and tasks.
//
// Mirrors CenterHead_GGA.get_targets_single / get_targets
// (/root/reference/mmdet3d/models/dense_heads/centerpoint_head_gga.py:343-627) with the Gaussian
// helpers of /root/reference/mmdet3d/core/utils/gaussian.py:6-86.  The reference loops in Python
// over tasks x objects (<= 500) x frames, a few dozen 0-dim tensor ops each; here one CTA packs one
// frame:
//   1. slot k of object i inside its task: objects of a task are concatenated class by class, each
//      class in input order (:426-433, 457-472)  ->  k = #{j : task_j = task_i and
//      (cls_j, j) < (cls_i, i)}; slots >= max_objs are dropped (:515);
//   2. per object, in the dtype T of the pseudo labels (fp32 or fp64, torch type promotion of the
//      0-dim operands): BEV extent in feature-map cells (:546-551), gaussian_radius with the
//      reference's three roots (gaussian.py:58-86, operation order kept), radius =
//      max(min_radius, (int)r) (:557), centre cell = trunc((float)((x - x0) / voxel / factor))
//      (:562-572), range check (:578);
//   3. heatmap[cls] = max(heatmap[cls], g) over the clipped (2r+1)^2 window with
//      g = (float)exp(-(dx^2 + dy^2) / (2 sigma sigma)) in double, sigma = (2r+1)/6
//      (gaussian.py:6-55) — max is order independent, so atomicMax on the bit pattern of the
//      non-negative floats reproduces the sequential result exactly;
//   4. ind / mask / anno_box = (x1, y1, x2, y2, srl) / lidar2img / boundary mask of slot k
//      (:588-617); unused slots keep 0, respectively the frame's base lidar2img (:509-511).
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxFrameObjs = 2048;  // objects of one frame held in shared memory

struct Consts {
  double one_minus, one_plus, b3k, c3k, a3x4;  // Python-evaluated scalars of gaussian_radius
  double pc_x0, pc_y0, voxel_x, voxel_y;       // fp32 values of the reference's CPU tensors, widened
  double factor;
};

template <typename T>
__device__ __forceinline__ T tsqrt(T v);
template <>
__device__ __forceinline__ float tsqrt<float>(float v) { return __fsqrt_rn(v); }
template <>
__device__ __forceinline__ double tsqrt<double>(double v) { return __dsqrt_rn(v); }

// gaussian_radius((height, width), min_overlap) of gaussian.py:58-86 in T arithmetic
template <typename T>
__device__ T radius_of(T height, T width, const Consts& c) {
  const T hw = height + width;
  const T b1 = hw;
  const T c1 = width * height * (T)c.one_minus / (T)c.one_plus;
  const T sq1 = tsqrt<T>(b1 * b1 - (T)4 * c1);
  const T r1 = (b1 + sq1) / (T)2;
  const T b2 = (T)2 * hw;
  const T c2 = (T)c.one_minus * width * height;
  const T sq2 = tsqrt<T>(b2 * b2 - (T)16 * c2);
  const T r2 = (b2 + sq2) / (T)2;
  const T b3 = (T)c.b3k * hw;
  const T c3 = (T)c.c3k * width * height;
  const T sq3 = tsqrt<T>(b3 * b3 - (T)c.a3x4 * c3);
  const T r3 = (b3 + sq3) / (T)2;
  T r = r1;             // Python min(r1, r2, r3): first of the smallest
  if (r2 < r) r = r2;
  if (r3 < r) r = r3;
  return r;
}

struct Draw {
  int cx, cy, radius, channel;
};

template <typename T>
__global__ void __launch_bounds__(kThreads) pack_targets_kernel(const gga_target_args a, const Consts c) {
  __shared__ int16_t s_task[kMaxFrameObjs];
  __shared__ int16_t s_cls[kMaxFrameObjs];
  __shared__ Draw s_draw[kMaxFrameObjs];
  __shared__ int s_ndraw;
  const int f = blockIdx.x, tid = threadIdx.x;
  const int o0 = a.frame_offsets[f], n = a.frame_offsets[f + 1] - o0;
  const int F = a.num_frames, K = a.max_objs;
  if (tid == 0) s_ndraw = 0;
  for (int i = tid; i < n; i += kThreads) {
    const int lab = a.labels[o0 + i];
    const bool ok = lab >= 0 && lab < a.n_classes;
    s_task[i] = ok ? (int16_t)a.class_task[lab] : (int16_t)-1;
    s_cls[i] = ok ? (int16_t)a.class_cls[lab] : (int16_t)0;
  }
  // unused slots: zeros, the frame's base calibration, source index -1
  for (int t = 0; t < a.n_tasks; ++t) {
    const long long row0 = ((long long)t * F + f) * K;
    for (int k = tid; k < K; k += kThreads) {
      a.ind[row0 + k] = 0;
      a.mask[row0 + k] = 0;
      a.src_index[row0 + k] = -1;
      reinterpret_cast<uint32_t*>(a.boundary_mask)[row0 + k] = 0u;
    }
    for (int k = tid; k < K * 5; k += kThreads) a.anno_box[row0 * 5 + k] = 0.f;
    for (int k = tid; k < K * 16; k += kThreads) a.anno_lidar2img[row0 * 16 + k] = a.base_lidar2img[f * 16 + (k & 15)];
  }
  __syncthreads();
  const T* pseudo = reinterpret_cast<const T*>(a.pseudo);
  for (int i = tid; i < n; i += kThreads) {
    const int t = s_task[i];
    if (t < 0) continue;
    const int cl = s_cls[i];
    int k = 0;
    for (int j = 0; j < n; ++j) k += (s_task[j] == t && (s_cls[j] < cl || (s_cls[j] == cl && j < i))) ? 1 : 0;
    if (k >= K) continue;
    const long long slot = ((long long)t * F + f) * K + k;
    const int obj = o0 + i;
    a.src_index[slot] = obj;
    const T* q = pseudo + (long long)obj * 7;
    const T width = q[3] / (T)c.voxel_x / (T)c.factor;
    const T length = q[4] / (T)c.voxel_y / (T)c.factor;
    if (!(width > (T)0 && length > (T)0)) continue;
    const T rr = radius_of<T>(length, width, c);
    int radius = (int)rr;  // Python int(): truncation
    radius = max(a.min_radius, radius);
    const float fx = (float)((q[0] - (T)c.pc_x0) / (T)c.voxel_x / (T)c.factor);
    const float fy = (float)((q[1] - (T)c.pc_y0) / (T)c.voxel_y / (T)c.factor);
    if (!(fx > -1.0f && fx < (float)a.fm_w && fy > -1.0f && fy < (float)a.fm_h)) continue;  // also NaN
    const int cx = (int)fx, cy = (int)fy;  // .to(torch.int32): truncation, (-1, 0) -> 0
    if (!(cx >= 0 && cx < a.fm_w && cy >= 0 && cy < a.fm_h)) continue;
    a.ind[slot] = (int64_t)cy * a.fm_w + cx;
    a.mask[slot] = 1;
    for (int e = 0; e < 16; ++e) a.anno_lidar2img[slot * 16 + e] = a.lidar2img[(long long)obj * 16 + e];
    for (int e = 0; e < 4; ++e) {
      a.boundary_mask[slot * 4 + e] = a.bdry[(long long)obj * 4 + e] ? 0 : 1;
      a.anno_box[slot * 5 + e] = a.boxes_img[(long long)obj * 4 + e];
    }
    a.anno_box[slot * 5 + 4] = a.srl[f * a.n_tasks + t];
    const int d = atomicAdd(&s_ndraw, 1);
    s_draw[d].cx = cx; s_draw[d].cy = cy; s_draw[d].radius = radius;
    s_draw[d].channel = a.task_channel0[t] + cl;
  }
  __syncthreads();
  // one warp per Gaussian, lanes over the clipped window
  const int lane = tid & 31, warp = tid >> 5;
  for (int d = warp; d < s_ndraw; d += kThreads / 32) {
    const Draw w = s_draw[d];
    const int r = w.radius, left = min(w.cx, r), right = min(a.fm_w - w.cx, r + 1);
    const int top = min(w.cy, r), bottom = min(a.fm_h - w.cy, r + 1);
    const int ww = left + right, hh = top + bottom;
    if (ww <= 0 || hh <= 0) continue;
    const double sigma = (double)(2 * r + 1) / 6.0;
    const double den = 2.0 * sigma * sigma;
    int* hm = reinterpret_cast<int*>(a.heatmap) + ((long long)f * a.n_channels + w.channel) * a.fm_h * a.fm_w;
    for (int e = lane; e < ww * hh; e += 32) {
      const int dy = e / ww - top, dx = e - (e / ww) * ww - left;
      const double x = (double)dx, y = (double)dy;
      double g = exp(-(x * x + y * y) / den);
      if (g < 2.220446049250313e-16) g = 0.0;  // h[h < eps * h.max()] = 0 (h.max() = 1 at the centre)
      const float gv = (float)g;
      atomicMax(hm + (long long)(w.cy + dy) * a.fm_w + (w.cx + dx), __float_as_int(gv));
    }
  }
}

}  // namespace

extern "C" int gga_pack_targets(const gga_target_args* args, void* stream) {
  GGA_REQUIRE(args, "null args");
  const gga_target_args& a = *args;
  GGA_REQUIRE(a.num_frames >= 0 && a.n_tasks >= 1 && a.n_classes >= 1 && a.max_objs >= 1, "bad sizes");
  GGA_REQUIRE(a.fm_w >= 1 && a.fm_h >= 1 && a.n_channels >= 1, "bad feature map size");
  GGA_REQUIRE(a.pseudo_dtype == GGA_F32 || a.pseudo_dtype == GGA_F64, "pseudo_dtype must be GGA_F32 or GGA_F64");
  GGA_REQUIRE(a.voxel_x > 0.f && a.voxel_y > 0.f && a.out_size_factor >= 1, "bad voxel size / out_size_factor");
  if (a.num_frames == 0) return GGA_OK;
  GGA_REQUIRE(a.labels && a.frame_offsets && a.boxes_img && a.lidar2img && a.pseudo && a.bdry && a.base_lidar2img &&
                  a.srl && a.class_task && a.class_cls && a.task_channel0,
              "null input pointer");
  GGA_REQUIRE(a.heatmap && a.anno_box && a.ind && a.mask && a.anno_lidar2img && a.boundary_mask && a.src_index,
              "null output pointer");
  GGA_REQUIRE(a.max_frame_objs >= 0 && a.max_frame_objs <= kMaxFrameObjs,
              "at most %d objects per frame (got %d)", kMaxFrameObjs, a.max_frame_objs);
  cudaStream_t st = gga_stream(stream);
  GGA_CHECK_CUDA(cudaMemsetAsync(a.heatmap, 0, (size_t)a.num_frames * a.n_channels * a.fm_h * a.fm_w * sizeof(float), st));
  Consts c;
  const double ov = (double)a.gaussian_overlap;  // the Python float of the config
  c.one_minus = 1 - ov;
  c.one_plus = 1 + ov;
  c.b3k = -2 * ov;
  c.c3k = ov - 1;
  c.a3x4 = 4 * (4 * ov);
  c.pc_x0 = (double)a.pc_x0; c.pc_y0 = (double)a.pc_y0;
  c.voxel_x = (double)a.voxel_x; c.voxel_y = (double)a.voxel_y;
  c.factor = (double)a.out_size_factor;
  if (a.pseudo_dtype == GGA_F32) pack_targets_kernel<float><<<a.num_frames, kThreads, 0, st>>>(a, c);
  else pack_targets_kernel<double><<<a.num_frames, kThreads, 0, st>>>(a, c);
  GGA_CHECK_CUDA(cudaGetLastError());
  return GGA_OK;
}
