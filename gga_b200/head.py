"""Head-level entry points of GGA training, same signatures as ``CenterHead_GGA``.

* :func:`gga_calculate_rotation` — ``centerpoint_head_gga.py:167-182``
* :func:`get_prediction_single` — ``centerpoint_head_gga.py:250-341``: decode BEV centre
  from the voxel index, ``dims = exp``, bottom centre, then corners -> per-object
  ``lidar2img`` -> depth clamp -> divide -> min/max.  The decode is a handful of
  element-wise torch ops (autograd-transparent); the corner/projection/min-max chain (~35
  torch kernels in the reference) is the single CUDA launch of :func:`box3d_project`.
* :func:`boundary_projection_loss` — the BPL of ``centerpoint_head_gga.py:682-687,714-720``.
"""
import torch

from .losses import box2d_loss
from .project import box3d_project


def gga_calculate_rotation(pred):
    rot_sine = pred[..., 0]
    rot_cosine = pred[..., 1]
    rot = torch.atan2(rot_sine, rot_cosine).squeeze()
    ones = torch.ones_like(rot_cosine)
    zeros = torch.zeros_like(rot_cosine)
    rmat_T = torch.stack([
        torch.stack([rot_cosine, rot_sine, zeros], dim=-1),
        torch.stack([-rot_sine, rot_cosine, zeros], dim=-1),
        torch.stack([zeros, zeros, ones], dim=-1),
    ], dim=-1)
    return rot, rmat_T


def get_prediction_single(pred_all, ind, ann_lidar2img, rot, train_cfg, norm_bbox=True):
    """Args as in the reference (``self.train_cfg`` / ``self.norm_bbox`` passed explicitly):
    pred_all [B, K, 8] = (reg_x, reg_y, height, dim_x, dim_y, dim_z, rot_sin, rot_cos),
    ind [B, K] int64, ann_lidar2img [B, K, 4, 4], rot [B, K].
    Returns (pred_ratio [B, K, 2], pred_iou [B, K, 4], pred_box_bev [B, K, 5])."""
    dev = pred_all.device
    osf = train_cfg['out_size_factor']
    fmap_x = int(train_cfg['grid_size'][0]) // int(osf)
    vs = train_cfg['voxel_size']
    pr = train_cfg['point_cloud_range']
    voxel_y = (torch.div(ind, fmap_x, rounding_mode='trunc') + pred_all[..., 1]) * vs[1] * osf + pr[1]
    voxel_x = ((ind % fmap_x) + pred_all[..., 0]) * vs[0] * osf + pr[0]
    b, k, _ = pred_all.shape
    if norm_bbox:
        dims = torch.exp(pred_all[..., 3:6])
    else:
        dims = pred_all[..., 3:6]
    z_bottom = pred_all[..., 2] - dims[..., 2] * 0.5
    boxes = torch.cat([voxel_x[..., None], voxel_y[..., None], z_bottom[..., None], dims,
                       rot.reshape(b, k, 1)], dim=-1)
    pred_iou, _ = box3d_project(boxes.reshape(-1, 7), ann_lidar2img.reshape(-1, 4, 4).to(dev),
                                mode='lidar_direct', depth_clamp=0.1)
    pred_iou = pred_iou.reshape(b, k, 4)
    w = torch.exp(pred_all[..., 3, None])
    h = torch.exp(pred_all[..., 4, None])
    pred_ratio = torch.cat([w, h], dim=-1)
    pred_box_bev = torch.cat([voxel_x[..., None], voxel_y[..., None], w, h, rot.reshape(b, k, 1)], dim=-1)
    return pred_ratio, pred_iou, pred_box_bev


def boundary_projection_loss(pred_iou, target_box, mask, boundary_mask, code_weight=0.5,
                             loss_weight=0.25, scale=0.3):
    """``loss_bpl * 0.3`` of ``centerpoint_head_gga.py:682-687,714-720`` with the config's
    ``L1Loss(reduction='mean', loss_weight=0.25)`` and ``code_weights`` 0.5."""
    num = mask.float().sum()
    m = mask.unsqueeze(2).expand_as(target_box).float()
    m = m * (~torch.isnan(target_box)).float()
    w = (m * code_weight)[..., :4] * boundary_mask.float()
    loss = box2d_loss(pred_iou, target_box[..., :4], w, avg_factor=(num + 1e-4), kind='l1',
                      reduction='mean', loss_weight=loss_weight)
    return loss * scale
