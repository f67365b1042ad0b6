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


# ----------------------------------------------------------------------------------------------
# Point-to-Box Alignment (PAL): centerpoint_head_gga.py:184-248 (distances), :690-699 (losses)
# ----------------------------------------------------------------------------------------------
def pack_in_box_points(ibp_points, device, max_objs=None):
    """``GGA_in_box_points`` — per frame a list (one entry per object) of ``[n_i, >=2]`` tensors
    (float64 ``(x, y, z, 1)`` in the reference, kitti_converter_gga.py:245-247) — packed ONCE
    into the CSR layout the kernel reads: (points_xy float32 [P, 2], offsets int32 [n_obj + 1],
    max_points int), objects in frame-major order.  The reference moves every list entry to the device
    separately (:469-470) and uses only ``[:, :2].float()`` (:201).

    ``max_objs`` = K of the head's ``[B, K, 5]`` boxes: the lists are RAGGED in the reference (a frame
    has n_obj <= max_objs clusters, centerpoint_head_gga.py:463-479,693) and the rows of the missing
    objects stay zero (:190-199); every frame is padded here to ``max_objs`` entries with empty objects
    so that object ``b * K + k`` of the packed layout is row ``[b, k]``."""
    if max_objs is not None:
        empty = torch.zeros((0, 2), dtype=torch.float32)
        padded = []
        for frame in ibp_points:
            frame = list(frame)
            assert len(frame) <= max_objs, f'{len(frame)} point clusters in a frame, but only {max_objs} objects'
            padded.append(frame + [empty] * (max_objs - len(frame)))
        ibp_points = padded
    flat = [t for frame in ibp_points for t in frame]
    counts = torch.tensor([0] + [int(t.shape[0]) for t in flat], dtype=torch.int64)
    offsets = torch.cumsum(counts, 0).to(torch.int32)
    if len(flat) and int(offsets[-1]) > 0:
        xy = torch.cat([t[:, :2].to(torch.float32).cpu() if not t.is_cuda else t[:, :2].to(torch.float32)
                        for t in flat if t.shape[0] > 0], 0)
    else:
        xy = torch.zeros((0, 2), dtype=torch.float32)
    return xy.to(device).contiguous(), offsets.to(device), int(counts.max())


class _PointBoxDistances(torch.autograd.Function):

    @staticmethod
    def forward(ctx, points_xy, offsets, box_bev, max_points):
        from . import _lib
        assert box_bev.is_cuda and points_xy.is_cuda and offsets.is_cuda, 'CUDA tensors required (no CPU fallback)'
        n = box_bev.shape[0]
        bev = box_bev.detach().float().contiguous()
        xy = points_xy.detach().float().contiguous()
        off = offsets.to(torch.int32).contiguous()
        assert off.numel() == n + 1 and bev.shape[1] == 5
        dist = torch.empty((n, 3), dtype=torch.float32, device=bev.device)
        jac = torch.empty((n, 3, 5), dtype=torch.float32, device=bev.device)
        L = _lib.load()
        ws = torch.empty((int(L.gga_pal_workspace_bytes(n, int(max_points))),), dtype=torch.uint8, device=bev.device)
        with torch.cuda.device(bev.device):
            _lib.check(L.gga_point_box_alignment(_lib.ptr(xy), _lib.ptr(off), _lib.ptr(bev), n, int(max_points),
                                                 _lib.ptr(dist), _lib.ptr(jac), _lib.ptr(ws), ws.numel(),
                                                 _lib.current_stream(bev.device)), 'point_box_alignment')
        ctx.save_for_backward(jac)
        ctx.in_dtype = box_bev.dtype
        return dist

    @staticmethod
    def backward(ctx, g):
        jac, = ctx.saved_tensors
        return None, None, torch.einsum('nk,nkj->nj', g.float(), jac).to(ctx.in_dtype), None


def point_box_distances(points_xy, offsets, box_bev, max_points=None):
    """Packed form: returns ``dist [n_obj, 3] = (min_dis, x_dis, y_dis)``, differentiable w.r.t.
    ``box_bev [n_obj, 5] = (cx, cy, w, h, rot)``; two CUDA launches for all objects.
    ``max_points`` = upper bound of the per-object list length (from ``pack_in_box_points``;
    computed here with a device sync when omitted)."""
    if max_points is None:
        o = offsets.to(torch.int64)
        max_points = int((o[1:] - o[:-1]).max().item()) if o.numel() > 1 else 0
    return _PointBoxDistances.apply(points_xy, offsets, box_bev, int(max_points))


def get_distance_bev(ibp_points, pred_box_bev, packed=None):
    """Same signature and return layout as ``CenterHead_GGA.get_distance_bev``
    (centerpoint_head_gga.py:241-248): ``ibp_points`` = per-frame lists of per-object point
    tensors, ``pred_box_bev [B, K, 5]``; returns ``(pts_min_dis, pts_x_dis, pts_y_dis)``, each
    ``[B, K, 1]``; a frame may list fewer than K clusters (the rows of the others are zero, like
    the reference's ``new_zeros`` rows).  Pass ``packed=pack_in_box_points(..., max_objs=K)`` to skip
    the per-call packing."""
    b, k, _ = pred_box_bev.shape
    xy, off, mx = packed if packed is not None else pack_in_box_points(ibp_points, pred_box_bev.device, max_objs=k)
    assert off.numel() == b * k + 1, f'expected {b * k} objects, got {off.numel() - 1} (pack with max_objs=K)'
    d = point_box_distances(xy, off, pred_box_bev.reshape(-1, 5), mx).reshape(b, k, 3)
    return d[..., 0:1], d[..., 1:2], d[..., 2:3]


def point_alignment_losses(p2c_min, p2c_x, p2c_y, mask, code_weight=0.5, loss_weight=0.25, scale=0.1):
    """``distancemin / distancex / distancey`` of centerpoint_head_gga.py:690-699: mmdet L1Loss
    against a zero target, weight ``mask * code_weights[0]``, ``avg_factor = mask.sum() + 1e-4``,
    ``loss_weight`` 0.25 (gga_kitti_config.py:60), then ``* 0.1``.  Inputs ``[B, K, 1]``, mask ``[B, K]``.
    The distances are sums of non-negative terms, so L1 against zero is a weighted sum (plain
    torch ops on [B, K] tensors; the heavy part is the distance kernel)."""
    num = mask.float().sum()
    w = (mask.float() * code_weight).unsqueeze(-1)
    out = []
    for d in (p2c_min, p2c_x, p2c_y):
        out.append((d.abs() * w).sum() / (num + 1e-4) * loss_weight * scale)
    return tuple(out)
