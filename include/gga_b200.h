/*
 * gga_b200.h — C ABI of libgga_b200.so: the B200 (sm_100a) implementation of GGA's
 * geometry hot path.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Conventions (mirroring how the reference's ops are called):
 *   - every `const float*` / output pointer is a DEVICE pointer unless the function name
 *     ends in `_host`; the caller owns all buffers; inputs are never written;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream) and the call returns without synchronising, like the mmcv ops run on the
 *     current torch stream (mmcv/ops/points_in_boxes.py; call sites
 *     /root/reference/mmdet3d/core/bbox/structures/base_box3d.py:534,566);
 *   - return value 0 = OK, negative = error (no exceptions cross the ABI; the Python
 *     wrapper turns them into RuntimeError like mmcv's TORCH_CHECK);
 *     gga_last_error() returns a thread-local message for the last failure;
 *   - fp32 data, int32 indices; `num_points` = M and `num_boxes` = T in mmcv's naming.
 */
#ifndef GGA_B200_H_
#define GGA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGA_OK 0
#define GGA_ERR_INVALID (-1)     /* bad argument (shape, stride, null pointer) */
#define GGA_ERR_CUDA (-2)        /* a CUDA runtime call failed */
#define GGA_ERR_UNSUPPORTED (-3) /* valid request this build cannot serve (e.g. too many boxes) */

int gga_version(void);
const char* gga_last_error(void);
/* sm count, compute capability and total memory of the current device */
int gga_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);

/* ------------------------------------------------------------------------------------
 * Part 1 — point -> 3D box membership.
 * Replaces mmcv.ops.points_in_boxes_{all,part,cpu} as re-exported at
 * /root/reference/mmdet3d/ops/__init__.py:12-13 and bound at
 * /root/reference/mmdet3d/core/bbox/structures/base_box3d.py:7.
 * Contract (bit-exact, the CPU one): SURVEY.md Appendix A.1 / oracle/pib_oracle.c.
 *   points : [B, num_points, pts_stride] floats, xyz first (pts_stride 3 = mmcv layout,
 *            4 = (x,y,z,r) KITTI layout, any >= 3 accepted)
 *   boxes  : [B, num_boxes, 7] = (x, y, z_bottom, dx, dy, dz, rz)
 * ---------------------------------------------------------------------------------- */

/* Words per point row of the bit-packed mask for `num_boxes` boxes:
 * 1, 2 or 4 for <= 32 / 64 / 128 boxes, else 8 * ceil(num_boxes / 256).
 * Box t of a point is bit (t & 31) of word (t >> 5); padding bits are zero. */
int gga_pib_row_words(int num_boxes);

/* The membership calls need no scratch memory: the per-frame box index lives in shared memory
 * of the one kernel a call launches, so concurrent calls on different streams share nothing.
 * (mmcv's ops allocate nothing either.) */

/* bits : uint32 [B, num_points, gga_pib_row_words(num_boxes)] */
int gga_points_in_boxes_bits(const float* points, int pts_stride, const float* boxes,
                             uint32_t* bits, int B, int num_points, int num_boxes, void* stream);

/* out : int32 [B, num_points, num_boxes], 0/1 — exact layout of mmcv points_in_boxes_all.
 * Every element is written (no pre-zeroing needed). */
int gga_points_in_boxes_all(const float* points, int pts_stride, const float* boxes, int32_t* out,
                            int B, int num_points, int num_boxes, void* stream);

/* out : int32 [B, num_points], index of the first enclosing box or -1 — mmcv
 * points_in_boxes_part. */
int gga_points_in_boxes_part(const float* points, int pts_stride, const float* boxes,
                             int32_t* out, int B, int num_points, int num_boxes, void* stream);

/* Compact form of bit-packed rows for callers that take the masks to the host: the list of
 * (row, box) pairs of the set bits — at KITTI densities ~25x fewer bytes than the rows.
 *   bits      : uint32 [num_rows, gga_pib_row_words(num_boxes)] (rows of one or several frames)
 *   row_base  : added to the row index written in the pairs (e.g. frame * num_points)
 *   pairs     : int32 [capacity, 2] = (row_base + row, box); order unspecified
 *   count     : int32 [1] device counter; += the number of set bits (exact even beyond `capacity`,
 *               pairs beyond it are dropped); zeroed first when reset_count != 0 */
int gga_pib_hit_list(const uint32_t* bits, int64_t num_rows, int num_boxes, int64_t row_base,
                     int32_t* pairs, int capacity, int32_t* count, int reset_count, void* stream);

/* HOST buffers in and out (the points_in_boxes_cpu signature: CPU tensors), computed on
 * the current device: H2D, kernel, D2H, synchronous; device buffers are allocated internally.
 * out : int32 [B, num_points, num_boxes]. */
int gga_points_in_boxes_all_host(const float* points, int pts_stride, const float* boxes,
                                 int32_t* out, int B, int num_points, int num_boxes);

/* ------------------------------------------------------------------------------------
 * Part 2 + 3 — box corners -> projection -> 8-corner min/max -> (clamped) 2D box, and the
 * projected-box vs 2D-target loss with its backward pass, one launch for n boxes.
 *
 * mode (which reference function is mirrored):
 *   GGA_PROJ_LIDAR_DIRECT : CenterHead_GGA.get_prediction_single,
 *        /root/reference/mmdet3d/models/dense_heads/centerpoint_head_gga.py:252-275,317-338.
 *        boxes = LiDAR (x,y,z_bottom,dx,dy,dz,yaw); proj = lidar2img 4x4; depth = max(q.z, depth_clamp).
 *   GGA_PROJ_KITTI_CAM    : KittiDataset_GGA*.convert_valid_bboxes,
 *        /root/reference/mmdet3d/datasets/kitti_dataset_GGA_match.py:713-748 (+ clamp :511-512).
 *        boxes = LiDAR; rt = rect @ Trv2c 4x4; proj = P2 padded to 4x4; no depth clamp.
 *   GGA_PROJ_CAM_CENTER   : PGDHead.get_proj_bbox2d core,
 *        /root/reference/mmdet3d/models/dense_heads/pgd_head.py:413-427.
 *        boxes = CAM (x,y,z,l,h,w,yaw) with origin (0.5,0.5,0.5); proj = cam2img 4x4.
 *   GGA_PROJ_CAM_BOTTOM   : CameraInstance3DBoxes.corners + points_cam2img
 *        (cam_box3d.py:116-157, utils.py:175-214); boxes = CAM, origin (0.5,1.0,0.5).
 * loss kind:
 *   GGA_LOSS_NONE, GGA_LOSS_GIOU (mmdet GIoULoss: 1 - giou, eps),
 *   GGA_LOSS_IOU_LINEAR / _SQUARE / _LOG (mmdet IoULoss modes), GGA_LOSS_L1 (mmdet L1Loss,
 *   the GGA Boundary-Projection Loss, centerpoint_head_gga.py:714-720).
 * ---------------------------------------------------------------------------------- */
#define GGA_PROJ_LIDAR_DIRECT 0
#define GGA_PROJ_KITTI_CAM 1
#define GGA_PROJ_CAM_CENTER 2
#define GGA_PROJ_CAM_BOTTOM 3

#define GGA_LOSS_NONE 0
#define GGA_LOSS_GIOU 1
#define GGA_LOSS_IOU_LINEAR 2
#define GGA_LOSS_IOU_SQUARE 3
#define GGA_LOSS_IOU_LOG 4
#define GGA_LOSS_L1 5

typedef struct gga_box_loss_args {
  /* inputs */
  const float* boxes;      /* [n, 7] */
  const float* proj;       /* 4x4 row-major; element stride between boxes = proj_stride floats */
  int proj_stride;         /* 16 = one matrix per box (GGA_lidar2img), 0 = one shared matrix;
                              with frame_of_box: matrix index = frame */
  const float* rt;         /* KITTI_CAM only: rect @ Trv2c 4x4, same striding rule as proj */
  int rt_stride;
  const int32_t* frame_of_box; /* optional [n]: selects proj/rt matrix (and img_hw) per frame */
  const float* img_hw;     /* optional [F or 1, 2] = (H, W): clamp + validity (KITTI_CAM) */
  const float* pcd_range;  /* optional [6]: centre-in-range validity (KITTI_CAM) */
  const float* target;     /* [n, 4] 2D boxes (x1,y1,x2,y2); required unless loss NONE */
  const float* weight;     /* optional; [n] (weight_cols 1) or [n, 4] (weight_cols 4) */
  int weight_cols;
  const float* grad_loss;  /* optional [n]: upstream gradient per box (reduction 'none');
                              NULL = every box gets grad_scale */
  int n;
  int mode;                /* GGA_PROJ_* */
  int loss_kind;           /* GGA_LOSS_* */
  int clamp_to_image;      /* 1: box2d output clamped to [0,W]x[0,H] (gradient is that of the raw box) */
  float depth_clamp;       /* LIDAR_DIRECT: 0.1 in the reference; <= 0 disables */
  float eps;               /* IoU/GIoU eps (mmdet GIoULoss default 1e-6) */
  float grad_scale;        /* dL_total / d(sum_i w_i * loss_i), e.g. loss_weight / avg_factor */
  /* outputs (any may be NULL) */
  float* box2d;            /* [n, 4] */
  uint8_t* valid;          /* [n] KITTI_CAM validity (kitti_dataset_GGA_match.py:741-748), else 1 */
  uint8_t* argidx;         /* [n, 4] corner index attaining (xmin, ymin, xmax, ymax), first wins */
  float* loss;             /* [n] unweighted per-box loss (L1: [n, 4] per side) */
  float* loss_sum;         /* [1]  sum_i w_i * loss_i  (deterministic order) */
  float* loss_accum;       /* optional [1]: += the same sum (running total over steps, reduced across ranks
                              at the log interval: gga_kitti_config.py:251-254) */
  float* grad_boxes;       /* [n, 7] d(total)/d(boxes) */
  float* grad_box2d;       /* [n, 4] d(total)/d(box2d) */
  float* grad_target;      /* [n, 4] d(total)/d(target) (PGD passes a prediction as target) */
  /* scratch of the deterministic reduction behind loss_sum (required iff loss_sum != NULL) */
  void* scratch;
  size_t scratch_bytes;
} gga_box_loss_args;

/* Scratch of the loss reductions (gga_box_project_loss, gga_box2d_loss, gga_box3d_aa_loss with a
 * non-NULL loss_sum): caller-owned device memory, 16-byte aligned, at least
 * gga_loss_scratch_bytes() bytes, ZEROED ONCE by the caller before its first use (the kernels
 * leave it zeroed).  Calls that can run concurrently (different streams, parallel graph
 * branches) must use different scratch buffers; calls ordered on one stream may share one.
 * The library itself keeps no device state. */
size_t gga_loss_scratch_bytes(void);

int gga_box_project_loss(const gga_box_loss_args* args, void* stream);

/* Backward of the projection alone: grad_boxes[n,7] from grad_box2d[n,4] and the saved
 * argidx (torch routes the gradient of min/max(dim) to the single returned index). */
int gga_box_project_backward(const float* boxes, const float* proj, int proj_stride,
                             const float* rt, int rt_stride, const int32_t* frame_of_box,
                             const uint8_t* argidx, const float* grad_box2d, float* grad_boxes,
                             int n, int mode, float depth_clamp, void* stream);

/* 2D loss on given boxes (no projection): per-box loss, weighted sum, gradients. */
int gga_box2d_loss(const float* pred, const float* target, const float* weight, int weight_cols,
                   const float* grad_loss, int n, int loss_kind, float eps, float grad_scale,
                   float* loss, float* loss_sum, float* grad_pred, float* grad_target,
                   void* scratch, size_t scratch_bytes, void* stream);

/* Axis-aligned 3-D IoU (giou = 0) / GIoU (giou = 1) loss of aligned pairs, boxes [n, 6] =
 * (x1, y1, z1, x2, y2, z2): AxisAlignedIoULoss, /root/reference/mmdet3d/models/losses/
 * axis_aligned_iou_loss.py:10-82, over axis_aligned_bbox_overlaps_3d (is_aligned),
 * mmdet3d/core/bbox/iou_calculators/iou3d_calculator.py:281-329 — the FCAF3D box loss
 * (fcaf3d_head.py:59,313-318).  loss = 1 - iou; weight [n] or NULL; outputs as gga_box2d_loss. */
int gga_box3d_aa_loss(const float* pred, const float* target, const float* weight,
                      const float* grad_loss, int n, int giou, float eps, float grad_scale,
                      float* loss, float* loss_sum, float* grad_pred, float* grad_target,
                      void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Matching — block-diagonal pairwise 2D IoU + argmax (pseudo-label matching).
 * Mirrors image_box_overlap (/root/reference/mmdet3d/core/evaluation/kitti_utils/eval.py:85-114,
 * criterion -1) as called with dt first by tools/utils_pseudo_labels_gga.py:45, followed by
 * np.argmax(axis=-1) (:60, first maximum wins).  Only the per-frame diagonal blocks that
 * calculate_iou_partly (eval.py:402-416) keeps are computed.
 *   dt : float32 [sum_dt, 4] (projected boxes), dt_offsets int32 [F+1]
 *   gt : float64 [sum_gt, 4] (annotation boxes), gt_offsets int32 [F+1]
 *   match : int32 [sum_dt] index into the frame's gt list (-1 if the frame has no gt)
 *   best_iou : float32 [sum_dt] (the float32 the reference's overlaps array holds)
 *   overlaps : optional float32, the concatenated row-major blocks [n_dt_f, n_gt_f]
 *              at ov_offsets int64 [F+1]
 * ---------------------------------------------------------------------------------- */
int gga_match_dt_gt(const float* dt, const int32_t* dt_offsets, const double* gt,
                    const int32_t* gt_offsets, int num_frames, int32_t* match, float* best_iou,
                    float* overlaps, const int64_t* ov_offsets, void* stream);

/* Dense pairwise IoU of the same function in float64: out [N, K]. */
int gga_image_box_overlap_f64(const double* boxes, int N, const double* query, int K,
                              int criterion, double* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Convex-polygon membership: _points_in_convex_polygon_3d_jit,
 * /root/reference/mmdet3d/core/bbox/box_np_ops.py:641-675 — the test under points_in_rbbox
 * (:353-376, the numpy/numba membership contract: ALL faces open, unlike the mmcv op) and under
 * the frustum membership of tools/data_converter/utils_gga.py:88-101.
 *   points : [N, pts_stride] float32 (points_f64 = 0) or float64, xyz first
 *   normal : [M, S, 3], d : [M, S] plane coefficients from surface_equ_3d (:617-638), float32 or
 *            float64 (planes_f64); arithmetic in the promoted type, left to right, unfused
 *   num_surfaces : optional int64 [M] (the reference's quirk `k > num_surfaces[j]` is kept)
 *   out    : uint8 [N, M], 1 = inside (n.p + d < 0 on every surface; NaN points are inside)
 * ---------------------------------------------------------------------------------- */
int gga_points_in_convex_polygons(const void* points, int pts_stride, int points_f64, const void* normal,
                                  const void* d, int planes_f64, const int64_t* num_surfaces, int N, int M,
                                  int S, uint8_t* out, void* stream);

/* FCAF3DHead._get_face_distances, /root/reference/mmdet3d/models/dense_heads/fcaf3d_head.py:495-520,
 * and its inside-box condition `min > 0` (:566-572).
 *   points [N, 3], boxes [M, 7] = (gravity centre, dims, yaw) -> dist [N, M, 6] =
 *   (dx_min, dx_max, dy_min, dy_max, dz_min, dz_max) and / or inside uint8 [N, M]. */
int gga_face_distances(const float* points, const float* boxes, int N, int M, float* dist, uint8_t* inside,
                       void* stream);

/* ------------------------------------------------------------------------------------
 * Point-to-Box Alignment distances (the PAL loss inputs of GGA training) with their Jacobian.
 * Mirrors CenterHead_GGA.get_distance_single / get_distance_bev,
 * /root/reference/mmdet3d/models/dense_heads/centerpoint_head_gga.py:184-248 (called :692):
 * a Python loop over objects with ~20 torch launches each in the reference, one launch here.
 *   points_xy : float32 [P, 2], the in-box points of all objects, object-major (the
 *               GGA_in_box_points lists, `clt[..., :2].float()` of :201, packed)
 *   offsets   : int32 [n_obj + 1] (CSR)
 *   box_bev   : float32 [n_obj, 5] = (cx, cy, w, h, rot)   (pred_box_bev, :277-286)
 *   dist      : float32 [n_obj, 3] = (min_dis, x_dis, y_dis)  (p2c_min, p2c_x, p2c_y)
 *   jac       : optional float32 [n_obj, 3, 5] = d dist / d box_bev
 *   max_points_per_object : upper bound of the list lengths (a host-side number known when the
 *               lists are packed; the reference caps them at 6000, kitti_converter_gga.py:410-413).
 *               Objects are split into 512-point chunks served by one warp each; an object with
 *               more points than the bound is an error of the caller (points beyond are ignored).
 *   workspace : caller-owned scratch for the per-chunk partial sums, gga_pal_workspace_bytes()
 *               bytes, any contents; results are summed in chunk order (deterministic).
 * ---------------------------------------------------------------------------------- */
size_t gga_pal_workspace_bytes(int n_obj, int max_points_per_object);
int gga_point_box_alignment(const float* points_xy, const int32_t* offsets, const float* box_bev,
                            int n_obj, int max_points_per_object, float* dist, float* jac,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * GGA training-target packing — CenterHead_GGA.get_targets / get_targets_single,
 * /root/reference/mmdet3d/models/dense_heads/centerpoint_head_gga.py:343-627, with
 * gaussian_radius / draw_heatmap_gaussian of /root/reference/mmdet3d/core/utils/gaussian.py:6-86.
 * One launch packs every frame and task (the reference: a Python loop over tasks x <= 500
 * objects x frames of 0-dim tensor ops).  Objects of all frames are concatenated; frame f owns
 * [frame_offsets[f], frame_offsets[f+1]).  Per task t the outputs are contiguous [F, max_objs, ...]
 * blocks (task major), the layout get_targets builds with torch.stack (:377-397).
 *   labels          int32 [n]       class index (gt_labels_3d); values outside [0, n_classes) match no task
 *   boxes_img       float [n, 4]    2D labels (GGA_boxes_img)             -> anno_box[:, 0:4]
 *   lidar2img       float [n, 16]   one 4x4 per object (GGA_lidar2img)    -> anno_lidar2img
 *   pseudo          fp32 or fp64 [n, 7]  initial pseudo 3D boxes (GGA_init_pseudo_labels); the
 *                   radius / centre arithmetic runs in this dtype, like torch's promotion of the
 *                   0-dim operands in the reference (:546-572)
 *   bdry            uint8 [n, 4]    GGA_bdry_masks (bool)                 -> boundary_mask = !bdry
 *   base_lidar2img  float [F, 16]   img_meta['lidar2img'] (fills unused slots, :509-511)
 *   srl             float [F, n_tasks]  the Semantic-Ratio sample of each (frame, task) (:515-527;
 *                   the reference draws it with torch.normal — randomness stays with the caller)
 *   class_task / class_cls  int32 [n_classes]  task of each class index and its index inside the task
 *   task_channel0   int32 [n_tasks] first heatmap channel of each task (channels = classes, task order)
 * outputs
 *   heatmap         float [F, n_channels, fm_h, fm_w]   (zeroed by the call)
 *   anno_box        float [n_tasks, F, max_objs, 5] = (x1, y1, x2, y2, srl)
 *   ind             int64 [n_tasks, F, max_objs]   cy * fm_w + cx
 *   mask            uint8 [n_tasks, F, max_objs]
 *   anno_lidar2img  float [n_tasks, F, max_objs, 16]
 *   boundary_mask   uint8 [n_tasks, F, max_objs, 4]  (4-byte aligned)
 *   src_index       int32 [n_tasks, F, max_objs]   object (global index) placed in each slot, -1 = none:
 *                   the order task_GGA_in_box_points lists the in-box point clusters in (:463-479)
 * ---------------------------------------------------------------------------------- */
#define GGA_F32 0
#define GGA_F64 1
typedef struct gga_target_args {
  const int32_t* labels;
  const int32_t* frame_offsets; /* [num_frames + 1] */
  const float* boxes_img;
  const float* lidar2img;
  const void* pseudo;
  const uint8_t* bdry;
  const float* base_lidar2img;
  const float* srl;
  const int32_t* class_task;
  const int32_t* class_cls;
  const int32_t* task_channel0;
  int32_t pseudo_dtype; /* GGA_F32 / GGA_F64 */
  int32_t num_frames, n_tasks, n_classes, n_channels;
  int32_t max_frame_objs; /* upper bound of objects per frame (<= 2048) */
  int32_t max_objs;       /* train_cfg max_objs * dense_reg */
  int32_t fm_w, fm_h;     /* grid_size[:2] // out_size_factor */
  int32_t out_size_factor, min_radius;
  float pc_x0, pc_y0, voxel_x, voxel_y; /* fp32, as the reference's torch.tensor(...) of the config lists */
  double gaussian_overlap;
  float* heatmap;
  float* anno_box;
  int64_t* ind;
  uint8_t* mask;
  float* anno_lidar2img;
  uint8_t* boundary_mask;
  int32_t* src_index;
} gga_target_args;
int gga_pack_targets(const gga_target_args* args, void* stream);

/* ------------------------------------------------------------------------------------
 * One training-shaped step from HOST buffers (the `points_in_boxes_cpu`-style contract of
 * /root/reference/mmdet3d/ops/__init__.py:12,38 extended to the whole loss step of
 * mmdet3d/models/dense_heads/centerpoint_head_gga.py:629-723): H2D copies, membership masks,
 * projection + loss forward/backward, D2H copies — pipelined frame by frame over `n_streams`
 * streams (the PCIe link is the bound); n_streams == 1 means ONE copy each way and one membership
 * launch over all frames (for callers that overlap whole steps with gga_step_submit_host).  A context owns every device buffer, stream and
 * event of one (frames, points, boxes) shape on the current device; run calls allocate nothing
 * and are synchronous.  Host buffers should be page-locked for the copies to overlap.
 *   points [F, N, pts_stride], boxes [F, M, 7], lidar2img [F, M, 16] (one 4x4 per object),
 *   target [F, M, 4], weight [F, M] or NULL
 *   bits uint32 [F, N, gga_pib_row_words(M)], loss_sum [1] (weighted SUM), grad_boxes [F*M, 7]
 *   (already scaled by loss_weight / avg_factor).
 * ---------------------------------------------------------------------------------- */
int gga_step_create(int num_frames, int num_points, int num_boxes, int pts_stride, int n_streams,
                    void** ctx_out);
int gga_step_destroy(void* ctx);
int gga_step_run_host(void* ctx, const float* points, const float* boxes, const float* lidar2img,
                      const float* target, const float* weight, int proj_mode, int loss_kind,
                      float loss_weight, float avg_factor, float eps, float depth_clamp,
                      uint32_t* bits, float* loss_sum, float* grad_boxes);
/* Same step, the masks returned as the hit list of gga_pib_hit_list instead of dense rows:
 * hits int32 [hit_capacity, 2] = (frame * N + point, box) in HOST memory, *n_hits = their number
 * (GGA_ERR_UNSUPPORTED if they did not fit).  At KITTI densities 1.2 MB instead of 30.7 MB cross PCIe. */
int gga_step_run_host_hits(void* ctx, const float* points, const float* boxes, const float* lidar2img,
                           const float* target, const float* weight, int proj_mode, int loss_kind,
                           float loss_weight, float avg_factor, float eps, float depth_clamp,
                           int32_t* hits, int hit_capacity, int32_t* n_hits, float* loss_sum, float* grad_boxes);
/* Asynchronous form.  submit enqueues the whole step and returns; EVERY host buffer passed to it
 * (inputs and results) must stay valid and untouched until gga_step_wait_host(ctx, ...) has returned.
 * A context holds one step in flight (GGA_ERR_INVALID otherwise); two contexts used alternately keep
 * both directions of the PCIe link busy: the H2D copies of step k+1 overlap the D2H copies of step k.
 * hit_capacity > 0 asks for the hit list (normally with `bits` NULL, i.e. instead of dense rows); its pairs
 * are delivered by wait (hits [hit_capacity, 2], *n_hits), which may be NULL otherwise.
 * gga_step_run_host / _run_host_hits are submit + wait. */
int gga_step_submit_host(void* ctx, const float* points, const float* boxes, const float* lidar2img,
                         const float* target, const float* weight, int proj_mode, int loss_kind,
                         float loss_weight, float avg_factor, float eps, float depth_clamp,
                         uint32_t* bits, int hit_capacity, float* loss_sum, float* grad_boxes);
int gga_step_wait_host(void* ctx, int32_t* hits, int32_t* n_hits);
/* `bits` may be NULL: the masks then stay on the device (the training use — the reference's GPU op
 * `points_in_boxes_all` leaves them there too, base_box3d.py:566) and only loss and gradients come
 * back.  gga_step_device_bits() gives the device buffer [F, N, W] they are in, valid until the
 * next run or destroy. */
int gga_step_device_bits(void* ctx, uint32_t** bits_device);

/* ------------------------------------------------------------------------------------
 * Test hooks (device build of include/gga_detmath.h and of the per-box preparation).
 * ---------------------------------------------------------------------------------- */
int gga_test_sincos(const float* x, int64_t n, float* sn, float* cs, void* stream);
/* prep : float [num_boxes, 8] = (cx, cy, cz_centre, hz, cosa, sina, hx, hy) */
int gga_test_box_prep(const float* boxes, int num_boxes, float* prep, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GGA_B200_H_ */
