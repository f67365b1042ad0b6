"""3D box -> projected 2D box (corners -> calib matrix -> 8-corner min/max -> clamp), with
autograd, backed by ONE CUDA launch per direction.

Mirrors, selected by ``mode``:

* ``'lidar_direct'`` — ``CenterHead_GGA.get_prediction_single``
  (``/root/reference/mmdet3d/models/dense_heads/centerpoint_head_gga.py:252-275,317-338``):
  LiDAR corners (``lidar_box3d.py:49-89``) x per-object ``lidar2img``, ``depth = max(d, 0.1)``.
* ``'kitti_cam'`` — ``KittiDataset_GGA*.convert_valid_bboxes``
  (``/root/reference/mmdet3d/datasets/kitti_dataset_GGA_match.py:713-748``, clamp ``:511-512``):
  ``limit_yaw`` -> ``convert_to(CAM, rect @ Trv2c)`` -> ``.corners`` -> ``points_cam2img(P2)``.
* ``'cam_center'`` — ``PGDHead.get_proj_bbox2d`` core (``pgd_head.py:413-427``): CAM boxes with
  origin (0.5, 0.5, 0.5), ``cam2img``.
* ``'cam_bottom'`` — ``CameraInstance3DBoxes.corners`` + ``points_cam2img``
  (``cam_box3d.py:116-157``, ``structures/utils.py:175-214``).
"""
import torch

from . import _lib

MODES = {'lidar_direct': _lib.PROJ_LIDAR_DIRECT, 'kitti_cam': _lib.PROJ_KITTI_CAM,
         'cam_center': _lib.PROJ_CAM_CENTER, 'cam_bottom': _lib.PROJ_CAM_BOTTOM,
         'depth_direct': _lib.PROJ_LIDAR_DIRECT}


def pad_proj(mat):
    """3x3 / 3x4 / 4x4 (optionally batched) -> 4x4, the eye-padding of utils.py:199-203."""
    d1, d2 = mat.shape[-2:]
    assert (d1, d2) in ((3, 3), (3, 4), (4, 4)), \
        f'The shape of the projection matrix ({d1}*{d2}) is not supported.'
    if (d1, d2) == (4, 4):
        return mat
    out = torch.eye(4, dtype=mat.dtype, device=mat.device).expand(*mat.shape[:-2], 4, 4).clone()
    out[..., :d1, :d2] = mat
    return out


def _mat_arg(mat, n, device, frame_of_box):
    """Returns (contiguous fp32 tensor, element stride between boxes)."""
    mat = pad_proj(mat.to(device=device, dtype=torch.float32))
    if mat.dim() == 2:
        return mat.contiguous(), 0
    mat = mat.reshape(-1, 4, 4).contiguous()
    if frame_of_box is None:
        assert mat.shape[0] == n, f'expected one matrix per box ({n}), got {mat.shape[0]}'
    return mat, 16


class _Ctx:
    pass


def _prepare(boxes, proj, rt, mode, frame_of_box, img_hw, pcd_range):
    assert boxes.is_cuda, 'box3d_project needs CUDA tensors (no CPU fallback)'
    assert boxes.shape[-1] == 7, f'boxes must be [..., 7], got {tuple(boxes.shape)}'
    dev = boxes.device
    c = _Ctx()
    c.lead = boxes.shape[:-1]
    c.boxes = boxes.detach().reshape(-1, 7).float().contiguous()
    c.n = c.boxes.shape[0]
    c.mode = MODES[mode] if isinstance(mode, str) else int(mode)
    c.fob = None if frame_of_box is None else \
        frame_of_box.to(device=dev, dtype=torch.int32).reshape(-1).contiguous()
    if c.fob is not None:
        assert c.fob.numel() == c.n
    c.proj, c.proj_stride = _mat_arg(proj.detach(), c.n, dev, c.fob)
    c.rt, c.rt_stride = (None, 0)
    if c.mode == _lib.PROJ_KITTI_CAM:
        assert rt is not None, "mode 'kitti_cam' needs rt = rect @ Trv2c"
        c.rt, c.rt_stride = _mat_arg(rt.detach(), c.n, dev, c.fob)
    c.img_hw = None if img_hw is None else \
        torch.as_tensor(img_hw, dtype=torch.float32).to(dev).reshape(-1, 2).contiguous()
    c.pcd = None if pcd_range is None else \
        torch.as_tensor(pcd_range, dtype=torch.float32).to(dev).reshape(6).contiguous()
    return c


def _fill(args, c):
    args.boxes = _lib.ptr(c.boxes)
    args.proj = _lib.ptr(c.proj)
    args.proj_stride = c.proj_stride
    args.rt = _lib.ptr(c.rt)
    args.rt_stride = c.rt_stride
    args.frame_of_box = _lib.ptr(c.fob)
    args.img_hw = _lib.ptr(c.img_hw)
    args.pcd_range = _lib.ptr(c.pcd)
    args.n = c.n
    args.mode = c.mode


class _BoxProject(torch.autograd.Function):

    @staticmethod
    def forward(ctx, boxes, proj, rt, mode, depth_clamp, img_hw, pcd_range, clamp, frame_of_box):
        c = _prepare(boxes, proj, rt, mode, frame_of_box, img_hw, pcd_range)
        dev = boxes.device
        box2d = torch.empty((c.n, 4), dtype=torch.float32, device=dev)
        valid = torch.empty((c.n,), dtype=torch.uint8, device=dev)
        argidx = torch.empty((c.n, 4), dtype=torch.uint8, device=dev)
        a = _lib.BoxLossArgs()
        _fill(a, c)
        a.loss_kind = _lib.LOSS_NONE
        a.clamp_to_image = 1 if clamp else 0
        a.depth_clamp = float(depth_clamp)
        a.eps = 1e-6
        a.grad_scale = 1.0
        a.box2d, a.valid, a.argidx = _lib.ptr(box2d), _lib.ptr(valid), _lib.ptr(argidx)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().gga_box_project_loss(a, _lib.current_stream(dev)), 'box_project')
        ctx.c = c
        ctx.argidx = argidx
        ctx.depth_clamp = float(depth_clamp)
        ctx.in_dtype = boxes.dtype
        valid = valid.bool().reshape(c.lead)
        ctx.mark_non_differentiable(valid)
        return box2d.reshape(*c.lead, 4), valid

    @staticmethod
    def backward(ctx, g_box2d, _g_valid):
        c = ctx.c
        g = g_box2d.reshape(-1, 4).float().contiguous()
        gb = torch.empty((c.n, 7), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().gga_box_project_backward(
                _lib.ptr(c.boxes), _lib.ptr(c.proj), c.proj_stride, _lib.ptr(c.rt), c.rt_stride,
                _lib.ptr(c.fob), _lib.ptr(ctx.argidx), _lib.ptr(g), _lib.ptr(gb), c.n, c.mode,
                ctx.depth_clamp, _lib.current_stream(g.device)), 'box_project_backward')
        return gb.reshape(*c.lead, 7).to(ctx.in_dtype), None, None, None, None, None, None, None, None


def box3d_project(boxes, proj, mode='lidar_direct', rt=None, depth_clamp=0.1, img_hw=None,
                  pcd_range=None, clamp=False, frame_of_box=None):
    """Projects 3D boxes to axis-aligned 2D boxes.

    Args:
        boxes (Tensor): [..., 7] CUDA fp32.
        proj (Tensor): [4,4] shared, [n,4,4] one per box, or [F,4,4] with ``frame_of_box``;
            3x3 / 3x4 are eye-padded like ``points_cam2img``.
        mode (str): see module docstring.
        rt (Tensor): ``rect @ Trv2c`` for ``'kitti_cam'`` (same shape rules as ``proj``).
        depth_clamp (float): ``'lidar_direct'`` only; 0.1 in the reference, <= 0 disables.
        img_hw: (H, W) or [F, 2]; enables the image validity test and, with ``clamp``, the
            clamp of ``bbox2result_kitti`` (the gradient is that of the unclamped box).
        pcd_range: 6 floats; centre-in-range validity (``kitti_dataset_GGA_match.py:745-748``).
    Returns:
        (box2d [..., 4] = (x1, y1, x2, y2), valid [...] bool)
    """
    return _BoxProject.apply(boxes, proj, rt, mode, depth_clamp, img_hw, pcd_range, clamp, frame_of_box)


def points_cam2img_boxes(boxes_cam, cam2img):
    """Convenience: ``CameraInstance3DBoxes(boxes).corners`` -> ``points_cam2img`` -> min/max."""
    return box3d_project(boxes_cam, cam2img, mode='cam_bottom')[0]
