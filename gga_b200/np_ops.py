"""The numpy-side membership family of the reference, computed on the GPU.

Same names, arguments and numpy in / numpy out conventions as
``/root/reference/mmdet3d/core/bbox/box_np_ops.py`` (``points_in_rbbox`` :353-376,
``points_in_convex_polygon_3d_jit`` :677-705) and
``/root/reference/tools/data_converter/utils_gga.py`` (``points_in_frustm_indices`` :88-101),
which the reference calls from its data pipelines and converters (``transforms_3d.py:446``,
``gga_processing.py:54``, ``kitti_converter_gga.py:186,341,376``, ``create_gt_database_gga.py:296``),
plus FCAF3D's ``_get_face_distances`` (``fcaf3d_head.py:495-520``) on CUDA tensors.

The per-box / per-frustum geometry (corners, surfaces, plane equations: M items) is format work
done on the host exactly as the reference does it (including its float32 torch rotation under
``array_converter``); the N x M x 6 plane tests run in ``gga_points_in_convex_polygons``.
There is no CPU fallback: the library and a CUDA device are required.
"""
import numpy as np
import torch

from . import _lib

_SURF_IDX = np.array([0, 1, 2, 3, 7, 6, 5, 4, 0, 3, 7, 4, 1, 5, 6, 2, 0, 4, 5, 1, 3, 2, 6, 7]).reshape(6, 4)


# ------------------------------------------------------------------ host-side geometry (M items)
def corners_nd(dims, origin=0.5):
    """box_np_ops.py:62-93."""
    ndim = int(dims.shape[1])
    cn = np.stack(np.unravel_index(np.arange(2 ** ndim), [2] * ndim), axis=1).astype(dims.dtype)
    if ndim == 2:
        cn = cn[[0, 1, 3, 2]]
    elif ndim == 3:
        cn = cn[[0, 1, 3, 2, 4, 5, 7, 6]]
    cn = cn - np.array(origin, dtype=dims.dtype)
    return dims.reshape([-1, 1, ndim]) * cn.reshape([1, 2 ** ndim, ndim])


def _rotate(points, angles, axis):
    """``rotation_3d_in_axis`` as numpy callers get it: float32 torch inside
    (core/utils/array_converter.py:296-299), cast back to the input dtype."""
    p = torch.from_numpy(np.ascontiguousarray(points)).float()
    a = torch.from_numpy(np.ascontiguousarray(angles)).float()
    s, c = torch.sin(a), torch.cos(a)
    one, zero = torch.ones_like(c), torch.zeros_like(c)
    if axis in (1, -2):
        m = torch.stack([torch.stack([c, zero, -s]), torch.stack([zero, one, zero]), torch.stack([s, zero, c])])
    elif axis in (2, -1):
        m = torch.stack([torch.stack([c, s, zero]), torch.stack([-s, c, zero]), torch.stack([zero, zero, one])])
    elif axis in (0, -3):
        m = torch.stack([torch.stack([one, zero, zero]), torch.stack([zero, c, s]), torch.stack([zero, -s, c])])
    else:
        raise ValueError(f'axis should in range [-3, -2, -1, 0, 1, 2], got {axis}')
    out = torch.einsum('aij,jka->aik', p, m) if p.shape[0] else p
    return out.numpy().astype(points.dtype)


def center_to_corner_box3d(centers, dims, angles=None, origin=(0.5, 1.0, 0.5), axis=1):
    """box_np_ops.py:171-200."""
    corners = corners_nd(dims, origin=origin)
    if angles is not None:
        corners = _rotate(corners, angles, axis)
    corners += centers.reshape([-1, 1, 3])
    return corners


def corner_to_surfaces_3d(corners):
    """box_np_ops.py:331-350 (and the identical ``_jit`` twin :256-278): [N, 8, 3] -> [N, 6, 4, 3]."""
    return corners[:, _SURF_IDX]


def surface_equ_3d(polygon_surfaces):
    """box_np_ops.py:617-638: (normal_vec [M, S, 3], d [M, S]) of a x + b y + c z + d = 0."""
    sv = polygon_surfaces[:, :, :2, :] - polygon_surfaces[:, :, 1:3, :]
    nv = np.cross(sv[:, :, 0, :], sv[:, :, 1, :])
    d = np.einsum('aij, aij->ai', nv, polygon_surfaces[:, :, 0, :])
    return nv, -d


def projection_matrix_to_CRT_kitti(proj):
    """box_np_ops.py:526-549."""
    CR, CT = proj[0:3, 0:3], proj[0:3, 3]
    Rinv, Cinv = np.linalg.qr(np.linalg.inv(CR))
    return np.linalg.inv(Cinv), np.linalg.inv(Rinv), Cinv @ CT


def get_frustum(bbox_image, C, near_clip=0.001, far_clip=100):
    """box_np_ops.py:584-614."""
    fku, fkv = C[0, 0], -C[1, 1]
    u0v0 = C[0:2, 2]
    z = np.array([near_clip] * 4 + [far_clip] * 4, dtype=C.dtype)[:, np.newaxis]
    b = bbox_image
    bc = np.array([[b[0], b[1]], [b[0], b[3]], [b[2], b[3]], [b[2], b[1]]], dtype=C.dtype)
    near = (bc - u0v0) / np.array([fku / near_clip, -fkv / near_clip], dtype=C.dtype)
    far = (bc - u0v0) / np.array([fku / far_clip, -fkv / far_clip], dtype=C.dtype)
    return np.concatenate([np.concatenate([near, far], axis=0), z], axis=1)


def camera_to_lidar(points, r_rect, velo2cam):
    """box_np_ops.py:13-32."""
    shp = list(points.shape[0:-1])
    if points.shape[-1] == 3:
        points = np.concatenate([points, np.ones(shp + [1])], axis=-1)
    return (points @ np.linalg.inv((r_rect @ velo2cam).T))[..., :3]


# ------------------------------------------------------------------ device part (N x M x S tests)
def _convex_device(points_t, stride, points_f64, nv, d, num_surfaces, N, device):
    L = _lib.load()
    M, S = nv.shape[0], nv.shape[1]
    planes_f64 = nv.dtype == np.float64
    if points_f64 and not planes_f64:
        nv, d, planes_f64 = nv.astype(np.float64), d.astype(np.float64), True
    tn = torch.from_numpy(np.ascontiguousarray(nv)).to(device)
    td = torch.from_numpy(np.ascontiguousarray(d)).to(device)
    tns = None if num_surfaces is None else torch.from_numpy(np.ascontiguousarray(num_surfaces, dtype=np.int64)).to(device)
    out = torch.empty((N, M), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _lib.check(L.gga_points_in_convex_polygons(_lib.ptr(points_t), stride, int(points_f64), _lib.ptr(tn), _lib.ptr(td),
                                                   int(planes_f64), _lib.ptr(tns), N, M, S, _lib.ptr(out),
                                                   _lib.current_stream(device)), 'points_in_convex_polygons')
    return out


def points_in_convex_polygon_3d_jit(points, polygon_surfaces, num_surfaces=None, device='cuda'):
    """box_np_ops.py:677-705: points [N, 3] numpy (or a CUDA tensor), surfaces
    [M, S, >=3, 3] numpy -> bool [N, M] (numpy for numpy points, CUDA bool tensor otherwise)."""
    nv, d = surface_equ_3d(np.asarray(polygon_surfaces)[:, :, :3, :])
    if isinstance(points, torch.Tensor):
        assert points.is_cuda and points.dtype in (torch.float32, torch.float64) and points.stride(-1) == 1
        pts = points
        out = _convex_device(pts, pts.stride(0), pts.dtype == torch.float64, nv, d, num_surfaces, pts.shape[0], pts.device)
        return out.bool()
    p = np.asarray(points)
    f64 = p.dtype == np.float64
    p = np.ascontiguousarray(p, dtype=np.float64 if f64 else np.float32)
    dev = torch.device(device)
    pts = torch.from_numpy(p).to(dev)
    out = _convex_device(pts, p.shape[1], f64, nv, d, num_surfaces, p.shape[0], dev)
    return out.cpu().numpy().astype(np.bool_)


def points_in_rbbox(points, rbbox, z_axis=2, origin=(0.5, 0.5, 0), device='cuda'):
    """box_np_ops.py:353-376 (all faces open; counter-clockwise boxes): points [N, 3+] and
    rbbox [M, 7] numpy -> bool [N, M]."""
    rbbox = np.asarray(rbbox)
    corners = center_to_corner_box3d(rbbox[:, :3], rbbox[:, 3:6], rbbox[:, 6], origin=origin, axis=z_axis)
    surfaces = corner_to_surfaces_3d(corners)
    pts = points if isinstance(points, torch.Tensor) else np.asarray(points)
    return points_in_convex_polygon_3d_jit(pts, surfaces, device=device)


def frustum_surfaces(rect, Trv2c, P2, bbox_shape):
    """The frustum construction of utils_gga.py:90-96 / remove_outside_points (box_np_ops.py:569-576)."""
    C, R, T = projection_matrix_to_CRT_kitti(P2)
    fr = get_frustum(np.asarray(bbox_shape).tolist(), C)
    fr -= T
    fr = np.linalg.inv(R) @ fr.T
    fr = camera_to_lidar(fr.T, rect, Trv2c)
    return corner_to_surfaces_3d(fr[np.newaxis, ...])


def points_in_frustm_indices(points, rect, Trv2c, P2, bbox_shape, device='cuda'):
    """tools/data_converter/utils_gga.py:88-101 (name as in the reference): bool [N, 1]."""
    return points_in_convex_polygon_3d_jit(points, frustum_surfaces(rect, Trv2c, P2, bbox_shape), device=device)


def face_distances(points, boxes, return_inside=False):
    """``FCAF3DHead._get_face_distances`` (fcaf3d_head.py:495-520) for points [N, 3] and gravity-
    centre boxes [M, 7] (CUDA fp32): [N, M, 6]; with ``return_inside`` also ``min > 0`` (:566-572).
    The reference expands both to [N, M, *] first; here nothing but the result is materialised."""
    assert points.is_cuda and boxes.is_cuda, 'CUDA tensors required (no CPU fallback)'
    p = points.detach().float().contiguous()
    b = boxes.detach().float().contiguous()
    n, m = p.shape[0], b.shape[0]
    dist = torch.empty((n, m, 6), dtype=torch.float32, device=p.device)
    inside = torch.empty((n, m), dtype=torch.uint8, device=p.device) if return_inside else None
    with torch.cuda.device(p.device):
        _lib.check(_lib.load().gga_face_distances(_lib.ptr(p), _lib.ptr(b), n, m, _lib.ptr(dist), _lib.ptr(inside),
                                                  _lib.current_stream(p.device)), 'face_distances')
    return (dist, inside.bool()) if return_inside else dist


def box3d_to_bbox(box3d, P2, device='cuda'):
    """``box_np_ops.box3d_to_bbox`` (box_np_ops.py:311-328): camera boxes [N, 7] numpy and P2 ->
    [N, 4] = (min xy, max xy) of the 8 projected corners; corners with origin (0.5, 1.0, 0.5),
    rotation about y (:171-200), ``points_cam2img`` (structures/utils.py:175-214).  One
    ``box3d_project(mode='cam_bottom')`` launch; float32 arithmetic like the reference's
    array_converter path, result in the input dtype."""
    from .project import box3d_project
    b = np.asarray(box3d)
    if b.shape[0] == 0:
        return np.zeros((0, 4), dtype=b.dtype)
    dev = torch.device(device)
    tb = torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)).to(dev)
    tp = torch.from_numpy(np.ascontiguousarray(P2, dtype=np.float32)).to(dev)
    out, _ = box3d_project(tb, tp, mode='cam_bottom')
    return out.cpu().numpy().astype(b.dtype)


def iou_jit(boxes, query_boxes, mode='iou', eps=0.0, device='cuda'):
    """``box_np_ops.iou_jit`` (box_np_ops.py:482-523): pairwise 2D IoU [N, K] of numpy boxes with
    ``eps`` added to every width / height / intersection side; ``mode != 'iou'`` divides by the
    area of ``boxes[n]`` only.  Adding eps to (x2, y2) of both sets turns this into the eps-free
    ``image_box_overlap`` kernel (criterion -1 / 0), evaluated in float64 and cast to
    ``boxes.dtype`` (the reference's numba loop mixes float32 differences with the float64 eps)."""
    from .matching import image_box_overlap
    b, q = np.asarray(boxes), np.asarray(query_boxes)
    if b.shape[0] == 0 or q.shape[0] == 0:
        return np.zeros((b.shape[0], q.shape[0]), dtype=b.dtype)
    dev = torch.device(device)
    grow = np.array([0.0, 0.0, eps, eps])
    tb = torch.from_numpy(b[:, :4].astype(np.float64) + grow).to(dev)
    tq = torch.from_numpy(q[:, :4].astype(np.float64) + grow).to(dev)
    ov = image_box_overlap(tb, tq, criterion=-1 if mode == 'iou' else 0)
    return ov.cpu().numpy().astype(b.dtype)
