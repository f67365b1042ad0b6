"""TEST INFRASTRUCTURE — CPU oracle for part 3 (2D IoU / GIoU / L1 consistency losses,
pairwise IoU and pseudo-label matching).  Never imported by ``gga_b200/``.

The 2D loss modules the reference selects by config (``pgd_head.py:72,112,744-748``:
``loss_consistency=dict(type='GIoULoss')``; ``gga_kitti_config.py:60``: ``L1Loss``) live in
the un-vendored ``mmdet`` (pinned 2.24.0 by ``docker/Dockerfile:4-6``; range
``>=2.24.0,<=3.0.0`` at ``mmdet3d/__init__.py:31-32``).  Their published algorithm is
restated here following the structurally identical *vendored* code:
``axis_aligned_bbox_overlaps_3d`` (``mmdet3d/core/bbox/iou_calculators/iou3d_calculator.py:
281-329``) with the z axis dropped, and ``AxisAlignedIoULoss`` (``mmdet3d/models/losses/
axis_aligned_iou_loss.py:10-82``) for the ``weighted_loss`` / early-out / ``loss_weight``
conventions.  The reference has no known-answer test for the 2D GIoU loss value
(``tests/test_models/test_heads/test_heads.py:1494`` only asserts ``>= 0``): **that part
of the parity is unpinned**; it is anchored on (a) the 3-axis golden of
``tests/test_metrics/test_losses.py:178-189`` reproduced by ``axis_aligned_iou_loss``
below, (b) the reference's own ``axis_aligned_bbox_overlaps_3d`` imported through
``oracle/ref_loader.py`` on boxes with a unit z extent, and (c) torchvision's
``generalized_box_iou_loss`` as an independent cross-check.
"""
import numpy as np
import torch


def bbox_overlaps_aligned(b1, b2, mode='iou', eps=1e-6):
    """2D twin of iou3d_calculator.py:281-329 (is_aligned=True branch); [..., 4] xyxy."""
    area1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    area2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    lt = torch.max(b1[..., :2], b2[..., :2])
    rb = torch.min(b1[..., 2:], b2[..., 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    union = area1 + area2 - overlap
    eps_t = union.new_tensor([eps])
    union = torch.max(union, eps_t)
    ious = overlap / union
    if mode == 'iou':
        return ious
    enclosed_lt = torch.min(b1[..., :2], b2[..., :2])
    enclosed_rb = torch.max(b1[..., 2:], b2[..., 2:])
    enclose_wh = (enclosed_rb - enclosed_lt).clamp(min=0)
    enclose_area = enclose_wh[..., 0] * enclose_wh[..., 1]
    enclose_area = torch.max(enclose_area, eps_t)
    return ious - (enclose_area - union) / enclose_area


def axis_aligned_overlaps_3d_aligned(b1, b2, mode='iou', eps=1e-6):
    """iou3d_calculator.py:281-329, is_aligned=True; [..., 6] = (x1,y1,z1,x2,y2,z2)."""
    area1 = (b1[..., 3] - b1[..., 0]) * (b1[..., 4] - b1[..., 1]) * (b1[..., 5] - b1[..., 2])
    area2 = (b2[..., 3] - b2[..., 0]) * (b2[..., 4] - b2[..., 1]) * (b2[..., 5] - b2[..., 2])
    lt = torch.max(b1[..., :3], b2[..., :3])
    rb = torch.min(b1[..., 3:], b2[..., 3:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1] * wh[..., 2]
    union = area1 + area2 - overlap
    eps_t = union.new_tensor([eps])
    union = torch.max(union, eps_t)
    ious = overlap / union
    if mode == 'iou':
        return ious
    elt = torch.min(b1[..., :3], b2[..., :3])
    erb = torch.max(b1[..., 3:], b2[..., 3:])
    ewh = (erb - elt).clamp(min=0)
    earea = torch.max(ewh[..., 0] * ewh[..., 1] * ewh[..., 2], eps_t)
    return ious - (earea - union) / earea


def weight_reduce_loss(loss, weight=None, reduction='mean', avg_factor=None):
    """mmdet `weight_reduce_loss` under `@weighted_loss` (axis_aligned_iou_loss.py:5,10).

    mmdet 2.24.0 (the Docker pin) divides by ``avg_factor``; later releases divide by
    ``avg_factor + finfo(float32).eps`` — a 1e-7 relative difference, inside the 1e-5
    tolerance of the contract (SURVEY.md §7 "mmdet version drift").
    """
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        if reduction == 'mean':
            return loss.mean()
        if reduction == 'sum':
            return loss.sum()
        return loss
    if reduction == 'mean':
        return loss.sum() / avg_factor
    if reduction == 'none':
        return loss
    raise ValueError('avg_factor can not be used with reduction="sum"')


def giou_loss_module(pred, target, weight=None, avg_factor=None, reduction='mean',
                     loss_weight=1.0, eps=1e-6):
    """mmdet ``GIoULoss.forward`` (call site pgd_head.py:744-748)."""
    if weight is not None and not torch.any(weight > 0):
        if pred.dim() == weight.dim() + 1:
            weight = weight.unsqueeze(1)
        return (pred * weight).sum()
    if weight is not None and weight.dim() > 1:
        assert weight.shape == pred.shape
        weight = weight.mean(-1)
    loss = 1 - bbox_overlaps_aligned(pred, target, 'giou', eps)
    return loss_weight * weight_reduce_loss(loss, weight, reduction, avg_factor)


def iou_loss_module(pred, target, weight=None, avg_factor=None, reduction='mean',
                    loss_weight=1.0, eps=1e-6, mode='log'):
    """mmdet ``IoULoss.forward`` (call site monoflex_head.py:90); mode in linear/square/log."""
    if weight is not None and not torch.any(weight > 0):
        if pred.dim() == weight.dim() + 1:
            weight = weight.unsqueeze(1)
        return (pred * weight).sum()
    if weight is not None and weight.dim() > 1:
        assert weight.shape == pred.shape
        weight = weight.mean(-1)
    ious = bbox_overlaps_aligned(pred, target, 'iou', eps).clamp(min=eps)
    if mode == 'linear':
        loss = 1 - ious
    elif mode == 'square':
        loss = 1 - ious ** 2
    else:
        loss = -ious.log()
    return loss_weight * weight_reduce_loss(loss, weight, reduction, avg_factor)


def l1_loss_module(pred, target, weight=None, avg_factor=None, reduction='mean', loss_weight=1.0):
    """mmdet ``L1Loss.forward`` as used for the GGA Boundary-Projection Loss
    (centerpoint_head_gga.py:714-720; config gga_kitti_config.py:60)."""
    if target.numel() == 0:
        return pred.sum() * 0
    loss = torch.abs(pred - target)
    return loss_weight * weight_reduce_loss(loss, weight, reduction, avg_factor)


def axis_aligned_iou_loss(pred, target, weight=None, avg_factor=None, reduction='mean',
                          loss_weight=1.0):
    """AxisAlignedIoULoss.forward, axis_aligned_iou_loss.py:47-82."""
    if (weight is not None) and (not torch.any(weight > 0)) and (reduction != 'none'):
        return (pred * weight).sum()
    loss = 1 - axis_aligned_overlaps_3d_aligned(pred, target, 'iou')
    return weight_reduce_loss(loss, weight, reduction, avg_factor) * loss_weight


def boundary_projection_loss(pred_iou, target_box, mask, boundary_mask, code_weight=0.5,
                             loss_weight=0.25, scale=0.3):
    """GGA BPL: centerpoint_head_gga.py:682-687, 714-720.

    pred_iou [B, K, 4]; target_box [B, K, 5] (first 4 = 2D box, NaN rows masked);
    mask [B, K] (0/1); boundary_mask [B, K, 4] (1 = supervise this side).
    """
    num = mask.float().sum()
    m = mask.unsqueeze(2).expand_as(target_box).float()
    m = m * (~torch.isnan(target_box)).float()
    w = (m * code_weight)[..., :4] * boundary_mask.float()
    tgt = target_box[..., :4]
    # NaN targets carry zero weight; abs(pred - NaN) * 0 would still be NaN in torch, the
    # reference relies on its targets being NaN-free where it matters.  Mirror it exactly.
    loss = l1_loss_module(pred_iou, tgt, w, avg_factor=(num + 1e-4), loss_weight=loss_weight)
    return loss * scale


def image_box_overlap(boxes, query_boxes, criterion=-1):
    """kitti_utils/eval.py:85-114 (numba in the reference): pairwise IoU, float64,
    no +1, no eps, zero unless iw > 0 and ih > 0."""
    boxes = np.asarray(boxes)
    q = np.asarray(query_boxes)
    N, K = boxes.shape[0], q.shape[0]
    out = np.zeros((N, K), dtype=boxes.dtype)
    if N == 0 or K == 0:
        return out
    qa = (q[:, 2] - q[:, 0]) * (q[:, 3] - q[:, 1])
    iw = np.minimum(boxes[:, None, 2], q[None, :, 2]) - np.maximum(boxes[:, None, 0], q[None, :, 0])
    ih = np.minimum(boxes[:, None, 3], q[None, :, 3]) - np.maximum(boxes[:, None, 1], q[None, :, 1])
    ba = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    ok = (iw > 0) & (ih > 0)
    inter = iw * ih
    if criterion == -1:
        ua = ba[:, None] + qa[None, :] - inter
    elif criterion == 0:
        ua = np.broadcast_to(ba[:, None], inter.shape)
    elif criterion == 1:
        ua = np.broadcast_to(qa[None, :], inter.shape)
    else:
        ua = np.ones_like(inter)
    with np.errstate(divide='ignore', invalid='ignore'):
        out[ok] = (inter / ua)[ok]
    return out


def match_dt_to_gt(dt_boxes, gt_boxes):
    """tools/utils_pseudo_labels_gga.py:45,60: overlaps [n_dt, n_gt] (dt first), argmax over
    the gt axis, numpy tie rule (first maximum).  Returns (match int64 [n_dt], best_iou)."""
    ov = image_box_overlap(np.asarray(dt_boxes, np.float64), np.asarray(gt_boxes, np.float64))
    if ov.shape[1] == 0:
        return np.full((ov.shape[0],), -1, np.int64), np.zeros((ov.shape[0],))
    m = np.argmax(ov, axis=-1)
    return m, ov[np.arange(ov.shape[0]), m]


def point_box_distances(points_xy, offsets, box_bev):
    """Point-to-Box Alignment distances: ``CenterHead_GGA.get_distance_single`` /
    ``get_distance_bev`` (centerpoint_head_gga.py:184-248), restated on a CSR layout.

    points_xy [P, 2] float32 (the ``clt[..., :2].float()`` of :201), object-major;
    offsets int [n_obj + 1]; box_bev [n_obj, 5] = (cx, cy, w, h, rot) (:277-286).
    Returns (min_dis, x_dis, y_dis), each [n_obj] (the reference's ``[B, K, 1]`` flattened):
      rotate points and centre clockwise by rot (utils.py:28-117, 2-D branch, clockwise):
        x' = x cos + y sin,  y' = -x sin + y cos
      min_dis = sum_p min(|x' - (cx' -+ w/2)|, |y' - (cy' -+ h/2)|)          (:209-227)
      x_dis   = sum_p relu(|x' - cx'| - 2 (w/2)),  y_dis likewise with h     (:215-219,228-229)
    Differentiable w.r.t. box_bev (torch autograd is the gradient oracle).
    """
    n = box_bev.shape[0]
    mins, xs, ys = [], [], []
    for i in range(n):
        clt = points_xy[int(offsets[i]):int(offsets[i + 1])]
        cx, cy, w, h, rot = box_bev[i, 0], box_bev[i, 1], box_bev[i, 2], box_bev[i, 3], box_bev[i, 4]
        c, s = torch.cos(rot), torch.sin(rot)
        px = clt[:, 0] * c + clt[:, 1] * s
        py = clt[:, 0] * (-s) + clt[:, 1] * c
        cxr = cx * c + cy * s
        cyr = cx * (-s) + cy * c
        half_l, half_h = w / 2.0, h / 2.0
        dx1, dx2 = px - (cxr - half_l), px - (cxr + half_l)
        dy1, dy2 = py - (cyr - half_h), py - (cyr + half_h)
        dx = torch.relu(torch.abs(px - cxr) - 2 * half_l)
        dy = torch.relu(torch.abs(py - cyr) - 2 * half_h)
        dis = torch.abs(torch.stack([dx1, dx2, dy1, dy2]).transpose(1, 0))
        all_dis = torch.min(dis, dim=-1)[0] if clt.shape[0] else dis.new_zeros((0,))
        mins.append(all_dis.sum())
        xs.append(dx.sum())
        ys.append(dy.sum())
    if n == 0:
        z = box_bev.new_zeros((0,))
        return z, z, z
    return torch.stack(mins), torch.stack(xs), torch.stack(ys)


def point_alignment_losses(min_dis, x_dis, y_dis, mask, code_weight=0.5, loss_weight=0.25, scale=0.1):
    """The three PAL terms of centerpoint_head_gga.py:690-699: mmdet ``L1Loss`` against a zero
    target, weight = mask * code_weights[0], avg_factor = mask.sum() + 1e-4, then * 0.1.
    min_dis / x_dis / y_dis / mask: [B, K]."""
    num = mask.float().sum()
    w = mask.float() * code_weight
    out = []
    for d in (min_dis, x_dis, y_dis):
        out.append(l1_loss_module(d, torch.zeros_like(d), w, avg_factor=(num + 1e-4), loss_weight=loss_weight) * scale)
    return tuple(out)
