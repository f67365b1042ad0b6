"""gga_b200 — B200 (sm_100a) implementation of GGA's geometry hot path:
membership masking, 3D-box -> 2D-box projection, IoU/GIoU/L1 consistency loss + backward,
pseudo-label matching.  Host side mirrors the reference's function / loss-module API; the
compute is hand-written CUDA behind the C ABI of include/gga_b200.h.  No CPU fallback."""
from . import _lib
from .ops import (hit_list, points_in_boxes_all, points_in_boxes_bits, points_in_boxes_cpu, points_in_boxes_part,
                  row_words, unpack_bits)
from .project import box3d_project, pad_proj
from .losses import (AxisAlignedIoULoss, GIoULoss, IoULoss, L1Loss, ProjectedGIoULoss, ProjectedIoULoss,
                     ProjectedL1Loss, axis_aligned_iou_loss, box2d_loss, projected_box_loss)
from . import np_ops
from .np_ops import (box3d_to_bbox, face_distances, iou_jit, points_in_convex_polygon_3d_jit, points_in_frustm_indices,
                     points_in_rbbox)
from .matching import convert_valid_bboxes_batch, image_box_overlap, match_dt_to_gt
from .head import (boundary_projection_loss, get_distance_bev, get_prediction_single, gga_calculate_rotation,
                   pack_in_box_points, point_alignment_losses, point_box_distances)
from .targets import get_targets, pack_targets, semantic_ratio_samples
from .kitti_format import bbox2result_kitti, kitti_lines, pseudo_label_matching_kitti

__all__ = [
    'points_in_boxes_all', 'points_in_boxes_part', 'points_in_boxes_cpu', 'points_in_boxes_bits',
    'row_words', 'unpack_bits', 'box3d_project', 'pad_proj', 'box2d_loss', 'projected_box_loss',
    'ProjectedGIoULoss', 'ProjectedIoULoss', 'ProjectedL1Loss', 'GIoULoss', 'IoULoss', 'L1Loss',
    'AxisAlignedIoULoss', 'axis_aligned_iou_loss', 'points_in_rbbox', 'points_in_convex_polygon_3d_jit',
    'points_in_frustm_indices', 'face_distances', 'np_ops',
    'convert_valid_bboxes_batch', 'image_box_overlap', 'match_dt_to_gt', 'get_prediction_single',
    'gga_calculate_rotation', 'boundary_projection_loss', 'get_distance_bev', 'pack_in_box_points',
    'point_box_distances', 'point_alignment_losses', 'get_targets', 'pack_targets', 'semantic_ratio_samples',
    'bbox2result_kitti', 'kitti_lines', 'pseudo_label_matching_kitti', 'box3d_to_bbox', 'iou_jit',
]
__version__ = '0.1.0'
