"""Projected-box vs 2D-target consistency losses (IoU / GIoU / L1) with the mmdet
loss-module signature, backed by CUDA kernels that produce the loss and its gradients in
one launch.

Signature mirrored: ``forward(pred, target, weight=None, avg_factor=None,
reduction_override=None, **kwargs)`` and ``__init__(reduction='mean', loss_weight=1.0, ...)``
of the vendored ``AxisAlignedIoULoss`` (``/root/reference/mmdet3d/models/losses/
axis_aligned_iou_loss.py:30-82``) and of the mmdet modules the reference selects by config:
``loss_consistency=dict(type='GIoULoss', loss_weight=1.0)`` (``pgd_head.py:72``, called at
``:744-748``), ``loss_bbox=dict(type='L1Loss', reduction='mean', loss_weight=0.25)``
(``configs/gga/gga_kitti_config.py:60``, Boundary-Projection Loss call at
``centerpoint_head_gga.py:714-720``), ``IoULoss`` (``monoflex_head.py:90``).

``projected_box_loss`` fuses projection + loss + backward to the 3D box parameters into a
single launch (the reference's path is ~100 tiny torch kernels).
"""
import torch
from torch import nn

from . import _lib
from .project import _fill, _prepare

KINDS = {'giou': _lib.LOSS_GIOU, 'iou_linear': _lib.LOSS_IOU_LINEAR, 'iou_square': _lib.LOSS_IOU_SQUARE,
         'iou_log': _lib.LOSS_IOU_LOG, 'l1': _lib.LOSS_L1}


def _weight_arg(weight, n, kind, device):
    """mmdet semantics: IoU-family losses average an [n, 4] weight over the last dim;
    L1 applies it element-wise ([n] broadcasts)."""
    if weight is None:
        return None, 0
    w = weight.detach().to(device=device, dtype=torch.float32)
    if w.dim() >= 2 and w.shape[-1] == 4 and w.numel() == n * 4:
        return w.reshape(n, 4).contiguous(), 4
    assert w.numel() == n, f'weight must have {n} or {n}x4 elements, got {tuple(weight.shape)}'
    return w.reshape(n).contiguous(), 1


def _reduce(per_sum, n_elems, reduction, avg_factor, loss_weight):
    """mmdet weight_reduce_loss on the already weighted sum (mmdet 2.24.0: / avg_factor)."""
    if avg_factor is None:
        if reduction == 'mean':
            return loss_weight / max(n_elems, 1)
        return loss_weight
    if reduction == 'mean':
        return loss_weight / avg_factor
    raise ValueError('avg_factor can not be used with reduction="sum"')


class _Box2DLoss(torch.autograd.Function):
    """Σ_i w_i · loss_i (or the per-box vector) and gradients to pred and target."""

    @staticmethod
    def forward(ctx, pred, target, weight, kind, eps, per_box):
        dev = pred.device
        lead = pred.shape[:-1]
        p = pred.detach().reshape(-1, 4).float().contiguous()
        t = target.detach().reshape(-1, 4).float().contiguous()
        n = p.shape[0]
        w, wc = _weight_arg(weight, n, kind, dev)
        ncol = 4 if kind == _lib.LOSS_L1 else 1
        loss = torch.empty((n, ncol), dtype=torch.float32, device=dev)
        loss_sum = torch.empty((1,), dtype=torch.float32, device=dev)
        gp = torch.empty((n, 4), dtype=torch.float32, device=dev)
        gt = torch.empty((n, 4), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            sc = _lib.loss_scratch(dev)
            _lib.check(_lib.load().gga_box2d_loss(
                _lib.ptr(p), _lib.ptr(t), _lib.ptr(w), wc, None, n, kind, float(eps), 1.0,
                _lib.ptr(loss), _lib.ptr(loss_sum), _lib.ptr(gp), _lib.ptr(gt), sc.data_ptr(), sc.numel(),
                _lib.current_stream(dev)), 'box2d_loss')
        ctx.save_for_backward(gp, gt)
        ctx.per_box = per_box
        ctx.shape = pred.shape
        ctx.ncol = ncol
        if per_box:
            if w is not None:
                loss = loss * (w.reshape(n, -1) if ncol == 4 or wc == 1 else w.mean(-1, keepdim=True))
            return loss.reshape(*lead, 4) if ncol == 4 else loss.reshape(lead)
        return loss_sum.reshape(())

    @staticmethod
    def backward(ctx, g):
        gp, gt = ctx.saved_tensors
        if ctx.per_box:
            g = g.reshape(-1, ctx.ncol).float()
            gp, gt = gp * g, gt * g
        else:
            gp, gt = gp * g, gt * g
        return gp.reshape(ctx.shape), gt.reshape(ctx.shape), None, None, None, None


def box2d_loss(pred, target, weight=None, avg_factor=None, kind='giou', reduction='mean',
               loss_weight=1.0, eps=1e-6):
    """Functional form; ``pred``/``target`` [..., 4] CUDA tensors."""
    assert pred.is_cuda, 'box2d_loss needs CUDA tensors (no CPU fallback)'
    k = KINDS[kind] if isinstance(kind, str) else int(kind)
    n_elems = pred.numel() // (1 if k == _lib.LOSS_L1 else 4)
    if reduction == 'none':
        if avg_factor is not None:
            pass  # mmdet: reduction 'none' ignores avg_factor
        return loss_weight * _Box2DLoss.apply(pred, target, weight, k, eps, True)
    s = _Box2DLoss.apply(pred, target, weight, k, eps, False)
    return s * _reduce(s, n_elems, reduction, avg_factor, loss_weight)


class _LossBase(nn.Module):
    kind = 'giou'

    def __init__(self, reduction='mean', loss_weight=1.0, eps=1e-6):
        super().__init__()
        assert reduction in ['none', 'sum', 'mean']
        self.reduction = reduction
        self.loss_weight = loss_weight
        self.eps = eps

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if weight is not None and not torch.any(weight > 0):
            # mmdet GIoULoss / IoULoss early-out: keeps the graph, returns 0
            if pred.dim() == weight.dim() + 1:
                weight = weight.unsqueeze(1)
            return (pred * weight).sum()
        return box2d_loss(pred, target, weight, avg_factor, self.kind, reduction, self.loss_weight, self.eps)


class ProjectedGIoULoss(_LossBase):
    """mmdet ``GIoULoss(eps=1e-6, reduction='mean', loss_weight=1.0)``: ``1 - giou``."""
    kind = 'giou'

    def __init__(self, eps=1e-6, reduction='mean', loss_weight=1.0):
        super().__init__(reduction, loss_weight, eps)


class ProjectedIoULoss(_LossBase):
    """mmdet ``IoULoss(linear=False, eps=1e-6, reduction='mean', loss_weight=1.0, mode='log')``."""

    def __init__(self, linear=False, eps=1e-6, reduction='mean', loss_weight=1.0, mode='log'):
        super().__init__(reduction, loss_weight, eps)
        assert mode in ['linear', 'square', 'log']
        if linear:
            mode = 'linear'
        self.mode = mode
        self.kind = 'iou_' + mode


class ProjectedL1Loss(nn.Module):
    """mmdet ``L1Loss(reduction='mean', loss_weight=1.0)`` on [..., 4] boxes (GGA BPL)."""

    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if target.numel() == 0:
            return pred.sum() * 0
        return box2d_loss(pred, target, weight, avg_factor, 'l1', reduction, self.loss_weight)


GIoULoss = ProjectedGIoULoss
IoULoss = ProjectedIoULoss
L1Loss = ProjectedL1Loss


class _ProjectedBoxLoss(torch.autograd.Function):
    """Projection + loss + backward in ONE launch: returns Σ_i w_i · loss_i, box2d, valid."""

    @staticmethod
    def forward(ctx, boxes, proj, target, weight, rt, mode, kind, eps, depth_clamp, frame_of_box,
                img_hw, pcd_range):
        dev = boxes.device
        c = _prepare(boxes, proj, rt, mode, frame_of_box, img_hw, pcd_range)
        t = target.detach().reshape(-1, 4).float().contiguous()
        assert t.shape[0] == c.n
        w, wc = _weight_arg(weight, c.n, kind, dev)
        ncol = 4 if kind == _lib.LOSS_L1 else 1
        box2d = torch.empty((c.n, 4), dtype=torch.float32, device=dev)
        valid = torch.empty((c.n,), dtype=torch.uint8, device=dev)
        loss = torch.empty((c.n, ncol), dtype=torch.float32, device=dev)
        loss_sum = torch.empty((1,), dtype=torch.float32, device=dev)
        gb = torch.empty((c.n, 7), dtype=torch.float32, device=dev)
        gt = torch.empty((c.n, 4), dtype=torch.float32, device=dev)
        a = _lib.BoxLossArgs()
        _fill(a, c)
        a.target, a.weight, a.weight_cols = _lib.ptr(t), _lib.ptr(w), wc
        a.loss_kind = kind
        a.clamp_to_image = 0
        a.depth_clamp, a.eps, a.grad_scale = float(depth_clamp), float(eps), 1.0
        a.box2d, a.valid, a.loss, a.loss_sum = _lib.ptr(box2d), _lib.ptr(valid), _lib.ptr(loss), _lib.ptr(loss_sum)
        a.grad_boxes, a.grad_target = _lib.ptr(gb), _lib.ptr(gt)
        with torch.cuda.device(dev):
            sc = _lib.loss_scratch(dev)
            a.scratch, a.scratch_bytes = sc.data_ptr(), sc.numel()
            _lib.check(_lib.load().gga_box_project_loss(a, _lib.current_stream(dev)), 'box_project_loss')
        ctx.save_for_backward(gb, gt)
        ctx.bshape, ctx.tshape, ctx.bdtype = boxes.shape, target.shape, boxes.dtype
        box2d = box2d.reshape(*c.lead, 4)
        valid = valid.bool().reshape(c.lead)
        loss = loss.reshape(*c.lead, 4) if ncol == 4 else loss.reshape(c.lead)
        ctx.mark_non_differentiable(box2d, valid, loss)
        return loss_sum.reshape(()), box2d, valid, loss

    @staticmethod
    def backward(ctx, g, _gb, _gv, _gl):
        gb, gt = ctx.saved_tensors
        return ((gb * g).reshape(ctx.bshape).to(ctx.bdtype), None, (gt * g).reshape(ctx.tshape), None, None,
                None, None, None, None, None, None, None)


def projected_box_loss(boxes, proj, target, weight=None, avg_factor=None, kind='giou', reduction='mean',
                       loss_weight=1.0, eps=1e-6, mode='lidar_direct', rt=None, depth_clamp=0.1,
                       frame_of_box=None, img_hw=None, pcd_range=None, return_box2d=False):
    """Fused ``box3d_project`` + 2D loss + backward (one kernel launch).

    Equivalent to ``Loss(kind)(box3d_project(boxes, proj, mode)[0], target, weight, avg_factor)``
    with the reference conventions documented above.  Returns the reduced loss (and, with
    ``return_box2d``, also the detached projected boxes and validity).
    """
    assert boxes.is_cuda, 'projected_box_loss needs CUDA tensors (no CPU fallback)'
    assert reduction in ('mean', 'sum')
    k = KINDS[kind] if isinstance(kind, str) else int(kind)
    s, box2d, valid, _ = _ProjectedBoxLoss.apply(boxes, proj, target, weight, rt, mode, k, eps, depth_clamp,
                                                 frame_of_box, img_hw, pcd_range)
    n_elems = boxes.numel() // 7 * (4 if k == _lib.LOSS_L1 else 1)
    out = s * _reduce(s, n_elems, reduction, avg_factor, loss_weight)
    return (out, box2d, valid) if return_box2d else out


# ----------------------------------------------------------------------------------------------
# AxisAlignedIoULoss (3-D, FCAF3D): mmdet3d/models/losses/axis_aligned_iou_loss.py:10-82
# ----------------------------------------------------------------------------------------------
class _Box3DAALoss(torch.autograd.Function):
    """Σ_i w_i (1 - iou_i) (or the per-pair vector) over aligned [.., 6] boxes + gradients."""

    @staticmethod
    def forward(ctx, pred, target, weight, giou, eps, per_box):
        dev = pred.device
        lead = pred.shape[:-1]
        p = pred.detach().reshape(-1, 6).float().contiguous()
        t = target.detach().reshape(-1, 6).float().contiguous()
        n = p.shape[0]
        w = None if weight is None else weight.detach().to(device=dev, dtype=torch.float32).reshape(n).contiguous()
        loss = torch.empty((n,), dtype=torch.float32, device=dev)
        loss_sum = torch.empty((1,), dtype=torch.float32, device=dev)
        gp = torch.empty((n, 6), dtype=torch.float32, device=dev)
        gt = torch.empty((n, 6), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            sc = _lib.loss_scratch(dev)
            _lib.check(_lib.load().gga_box3d_aa_loss(_lib.ptr(p), _lib.ptr(t), _lib.ptr(w), None, n, int(giou),
                                                     float(eps), 1.0, _lib.ptr(loss), _lib.ptr(loss_sum),
                                                     _lib.ptr(gp), _lib.ptr(gt), sc.data_ptr(), sc.numel(),
                                                     _lib.current_stream(dev)),
                       'box3d_aa_loss')
        ctx.save_for_backward(gp, gt)
        ctx.per_box, ctx.shape = per_box, pred.shape
        if per_box:
            return (loss * w if w is not None else loss).reshape(lead)
        return loss_sum.reshape(())

    @staticmethod
    def backward(ctx, g):
        gp, gt = ctx.saved_tensors
        if ctx.per_box:
            g = g.reshape(-1, 1).float()
        return (gp * g).reshape(ctx.shape), (gt * g).reshape(ctx.shape), None, None, None, None


def axis_aligned_iou_loss(pred, target, weight=None, avg_factor=None, reduction='mean', loss_weight=1.0,
                          mode='iou', eps=1e-6):
    """Functional form of ``AxisAlignedIoULoss`` on [..., 6] CUDA boxes (x1, y1, z1, x2, y2, z2)."""
    assert pred.is_cuda, 'axis_aligned_iou_loss needs CUDA tensors (no CPU fallback)'
    assert mode in ('iou', 'giou')
    n = pred.numel() // 6
    if reduction == 'none':
        return loss_weight * _Box3DAALoss.apply(pred, target, weight, mode == 'giou', eps, True)
    s = _Box3DAALoss.apply(pred, target, weight, mode == 'giou', eps, False)
    return s * _reduce(s, n, reduction, avg_factor, loss_weight)


class AxisAlignedIoULoss(nn.Module):
    """Same constructor / forward signature as the reference module
    (``axis_aligned_iou_loss.py:30-82``): ``loss = 1 - iou`` of aligned axis-aligned 3-D boxes,
    mmdet ``weighted_loss`` reduction, early-out ``(pred * weight).sum()`` when no weight is
    positive (:74-76)."""

    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        assert reduction in ['none', 'sum', 'mean']
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if (weight is not None) and (not torch.any(weight > 0)) and (reduction != 'none'):
            if weight.dim() == pred.dim() - 1:
                weight = weight.unsqueeze(-1)
            return (pred * weight).sum()
        return axis_aligned_iou_loss(pred, target, weight, avg_factor, reduction, self.loss_weight)
