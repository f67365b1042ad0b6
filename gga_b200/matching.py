"""Pseudo-label matching on the GPU: batched ``convert_valid_bboxes`` + block-diagonal
2D IoU + argmax over all frames of a shard.

Reference functions mirrored:
* ``KittiDataset_GGA_match.convert_valid_bboxes`` (``/root/reference/mmdet3d/datasets/
  kitti_dataset_GGA_match.py:685-765``) and the clamp of ``bbox2result_kitti`` (``:508-512``)
  -> :func:`convert_valid_bboxes_batch` (per-frame Python loop on CPU tensors in the reference);
* ``image_box_overlap`` (``/root/reference/mmdet3d/core/evaluation/kitti_utils/eval.py:85-114``)
  -> :func:`image_box_overlap`;
* ``calculate_iou_partly(dt, gt, metric=0)`` + ``np.argmax(axis=-1)``
  (``eval.py:343-418``; ``/root/reference/tools/utils_pseudo_labels_gga.py:45,60``)
  -> :func:`match_dt_to_gt`.
"""
import numpy as np
import torch

from . import _lib
from .project import box3d_project


def convert_valid_bboxes_batch(boxes_lidar, frame_of_box, rect, Trv2c, P2, img_hw, pcd_range):
    """All frames of a shard in one launch.

    Args:
        boxes_lidar (Tensor): [n, 7] LiDAR boxes of all frames, concatenated.
        frame_of_box (Tensor): [n] int frame index of each box.
        rect, Trv2c, P2 (Tensor): [F, 4, 4] (P2 may be [F, 3, 4]) calibration per frame
            (``info['calib']`` in the reference, cast to float32 like ``:724-726``).
        img_hw (Tensor): [F, 2] (H, W) = ``info['image']['image_shape']``.
        pcd_range: 6 floats, ``self.pcd_limit_range``.
    Returns:
        dict(bbox [n,4] clamped to the image, valid [n] bool) — the reference keeps the rows
        where ``valid`` and drops the rest (``:750-757``).
    """
    rt = torch.matmul(rect.float(), Trv2c.float())  # rect @ Trv2c, fp32 like :730
    bbox, valid = box3d_project(boxes_lidar, P2, mode='kitti_cam', rt=rt, img_hw=img_hw,
                                pcd_range=pcd_range, clamp=True, frame_of_box=frame_of_box)
    return dict(bbox=bbox, valid=valid)


def image_box_overlap(boxes, query_boxes, criterion=-1):
    """float64 pairwise IoU [N, K] with the reference's numba semantics (no +1, no eps,
    zero unless ``iw > 0 and ih > 0``).  CUDA float64 tensors in, CUDA float64 out."""
    assert boxes.is_cuda and query_boxes.is_cuda, 'CUDA tensors required (no CPU fallback)'
    b = boxes.detach().double().contiguous()
    q = query_boxes.detach().double().contiguous()
    N, K = b.shape[0], q.shape[0]
    out = torch.zeros((N, K), dtype=torch.float64, device=b.device)
    with torch.cuda.device(b.device):
        _lib.check(_lib.load().gga_image_box_overlap_f64(_lib.ptr(b), N, _lib.ptr(q), K, int(criterion),
                                                         _lib.ptr(out), _lib.current_stream(b.device)),
                   'image_box_overlap')
    return out


def match_dt_to_gt(dt_boxes, dt_offsets, gt_boxes, gt_offsets, return_overlaps=False):
    """Block-diagonal IoU + argmax.

    Args:
        dt_boxes (Tensor): float32 [sum_dt, 4] projected detections (all frames).
        dt_offsets (Tensor): int32 [F + 1] CSR offsets of the frames into ``dt_boxes``.
        gt_boxes (Tensor): float64 [sum_gt, 4] annotation boxes.
        gt_offsets (Tensor): int32 [F + 1].
    Returns:
        match int32 [sum_dt] (index inside the frame's gt list, -1 if it has none),
        best_iou float32 [sum_dt], and optionally the float32 overlaps of every block
        (flattened, with int64 offsets).
    """
    assert dt_boxes.is_cuda, 'CUDA tensors required (no CPU fallback)'
    dev = dt_boxes.device
    dt = dt_boxes.detach().float().contiguous()
    gt = gt_boxes.detach().to(dev).double().contiguous()
    do = dt_offsets.to(device=dev, dtype=torch.int32).contiguous()
    go = gt_offsets.to(device=dev, dtype=torch.int32).contiguous()
    F = do.numel() - 1
    assert go.numel() == F + 1
    match = torch.empty((dt.shape[0],), dtype=torch.int32, device=dev)
    best = torch.empty((dt.shape[0],), dtype=torch.float32, device=dev)
    ov, oo = None, None
    if return_overlaps:
        sizes = (do[1:] - do[:-1]).long() * (go[1:] - go[:-1]).long()
        oo = torch.zeros((F + 1,), dtype=torch.int64, device=dev)
        oo[1:] = torch.cumsum(sizes, 0)
        ov = torch.zeros((int(oo[-1].item()),), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().gga_match_dt_gt(_lib.ptr(dt), _lib.ptr(do), _lib.ptr(gt), _lib.ptr(go), F,
                                               _lib.ptr(match), _lib.ptr(best), _lib.ptr(ov), _lib.ptr(oo),
                                               _lib.current_stream(dev)), 'match_dt_gt')
    if return_overlaps:
        return match, best, ov, oo
    return match, best


def fix_matched_dims(dimensions, rotation_y):
    """``utils_pseudo_labels_gga.py:74-78``: swap l/w and rotate by pi/2 where ``dim[2] > dim[0]``
    (host-side numpy, part of the annos rewrite)."""
    dimensions = np.array(dimensions, copy=True)
    rotation_y = np.array(rotation_y, copy=True)
    swap = dimensions[:, 2] > dimensions[:, 0]
    dimensions[swap] = dimensions[swap][:, [2, 1, 0]]
    rotation_y[swap] = rotation_y[swap] + np.pi / 2.0
    return dimensions, rotation_y
