// Pseudo-label matching: block-diagonal pairwise 2D IoU + argmax, one CTA per frame.
//
// Mirrors image_box_overlap (/root/reference/mmdet3d/core/evaluation/kitti_utils/eval.py:85-114)
// exactly as tools/utils_pseudo_labels_gga.py:45 feeds it: `boxes` = detections, float32 (the
// .numpy() of convert_valid_bboxes' torch tensor, kitti_dataset_GGA_match.py:752), `query_boxes`
// = annotation boxes, float64.  numba types that call as: detection area in float32, everything
// else in float64, result stored into a float32 array (verified against the reference itself:
// tests/golden/ref_iou.npz `ibo_f32_f64`).  np.argmax(axis=-1) (:60) returns the first maximum.
// calculate_iou_partly (eval.py:343-418) computes whole num_parts x num_parts blocks and keeps
// the per-frame diagonal; only the diagonal is computed here.
#include "common.cuh"

namespace {

__device__ __forceinline__ float overlap_f32_f64(const float4 b, const double* __restrict__ q) {
  const double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  const double qarea = __dmul_rn(__dsub_rn(q2, q0), __dsub_rn(q3, q1));
  const double iw = __dsub_rn(fmin((double)b.z, q2), fmax((double)b.x, q0));
  if (!(iw > 0.0)) return 0.f;
  const double ih = __dsub_rn(fmin((double)b.w, q3), fmax((double)b.y, q1));
  if (!(ih > 0.0)) return 0.f;
  const float barea = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  const double inter = __dmul_rn(iw, ih);
  const double ua = __dsub_rn(__dadd_rn((double)barea, qarea), inter);
  return __double2float_rn(__ddiv_rn(inter, ua));
}

__global__ void __launch_bounds__(128) match_kernel(const float* __restrict__ dt,
                                                    const int32_t* __restrict__ dt_off,
                                                    const double* __restrict__ gt,
                                                    const int32_t* __restrict__ gt_off,
                                                    int32_t* __restrict__ match,
                                                    float* __restrict__ best_iou,
                                                    float* __restrict__ overlaps,
                                                    const int64_t* __restrict__ ov_off) {
  const int f = blockIdx.x;
  const int d0 = dt_off[f], d1 = dt_off[f + 1], g0 = gt_off[f], g1 = gt_off[f + 1];
  const int ng = g1 - g0;
  for (int i = d0 + threadIdx.x; i < d1; i += blockDim.x) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(dt) + i);
    int best = ng > 0 ? 0 : -1;
    float bv = 0.f;
    float* orow = overlaps ? overlaps + ov_off[f] + (int64_t)(i - d0) * ng : nullptr;
    for (int k = 0; k < ng; ++k) {
      const float v = overlap_f32_f64(b, gt + (int64_t)(g0 + k) * 4);
      if (orow) orow[k] = v;
      if (k == 0 || v > bv) { bv = v; best = k; }  // first maximum wins (np.argmax)
    }
    match[i] = best;
    if (best_iou) best_iou[i] = bv;
  }
}

__global__ void overlap_f64_kernel(const double* __restrict__ boxes, int N,
                                   const double* __restrict__ query, int K, int criterion,
                                   double* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * K) return;
  const int n = (int)(idx / K), k = (int)(idx % K);
  const double* b = boxes + (long long)n * 4;
  const double* q = query + (long long)k * 4;
  double r = 0.0;
  const double iw = __dsub_rn(fmin(b[2], q[2]), fmax(b[0], q[0]));
  if (iw > 0.0) {
    const double ih = __dsub_rn(fmin(b[3], q[3]), fmax(b[1], q[1]));
    if (ih > 0.0) {
      const double qarea = __dmul_rn(__dsub_rn(q[2], q[0]), __dsub_rn(q[3], q[1]));
      const double barea = __dmul_rn(__dsub_rn(b[2], b[0]), __dsub_rn(b[3], b[1]));
      const double inter = __dmul_rn(iw, ih);
      double ua;
      if (criterion == -1) ua = __dsub_rn(__dadd_rn(barea, qarea), inter);
      else if (criterion == 0) ua = barea;
      else if (criterion == 1) ua = qarea;
      else ua = 1.0;
      r = __ddiv_rn(inter, ua);
    }
  }
  out[idx] = r;
}

}  // namespace

extern "C" int gga_match_dt_gt(const float* dt, const int32_t* dt_offsets, const double* gt,
                               const int32_t* gt_offsets, int num_frames, int32_t* match,
                               float* best_iou, float* overlaps, const int64_t* ov_offsets,
                               void* stream) {
  GGA_REQUIRE(num_frames >= 0, "negative num_frames");
  if (num_frames == 0) return GGA_OK;
  GGA_REQUIRE(dt_offsets && gt_offsets && match, "null pointer");
  GGA_REQUIRE(!overlaps || ov_offsets, "overlaps needs ov_offsets");
  match_kernel<<<num_frames, 128, 0, gga_stream(stream)>>>(dt, dt_offsets, gt, gt_offsets, match,
                                                           best_iou, overlaps, ov_offsets);
  GGA_CHECK_CUDA(cudaGetLastError());
  return GGA_OK;
}

extern "C" int gga_image_box_overlap_f64(const double* boxes, int N, const double* query, int K,
                                         int criterion, double* out, void* stream) {
  GGA_REQUIRE(N >= 0 && K >= 0, "negative size");
  if (N == 0 || K == 0) return GGA_OK;
  GGA_REQUIRE(boxes && query && out, "null pointer");
  const long long total = (long long)N * K;
  overlap_f64_kernel<<<(unsigned)((total + 255) / 256), 256, 0, gga_stream(stream)>>>(boxes, N, query, K,
                                                                                      criterion, out);
  GGA_CHECK_CUDA(cudaGetLastError());
  return GGA_OK;
}
