"""TEST INFRASTRUCTURE — CPU oracle of the convex-polygon membership family and FCAF3D's face
distances.  numpy / torch restatements of

* ``box_np_ops.corners_nd`` (:62-93), ``center_to_corner_box3d`` (:171-200; its rotation runs in
  float32 torch through ``array_converter``, core/utils/array_converter.py:296-299),
  ``corner_to_surfaces_3d`` (:331-350), ``surface_equ_3d`` (:617-638),
  ``_points_in_convex_polygon_3d_jit`` (:641-675), ``points_in_rbbox`` (:353-376);
* ``projection_matrix_to_CRT_kitti`` (:526-549), ``get_frustum`` (:584-614), ``camera_to_lidar``
  (:13-32), ``corner_to_surfaces_3d_jit`` (:256-278) and
  ``tools/data_converter/utils_gga.points_in_frustm_indices`` (:88-101);
* ``FCAF3DHead._get_face_distances`` (fcaf3d_head.py:495-520) + ``min > 0`` (:566-572).

Pinned against outputs of the reference itself (tests/golden/ref_convex.npz, ref_rbbox.npz) in
tests/test_oracle_convex.py.  Never imported by ``gga_b200/``.
"""
import numpy as np
import torch


def corners_nd(dims, origin=0.5):
    ndim = int(dims.shape[1])
    cn = np.stack(np.unravel_index(np.arange(2 ** ndim), [2] * ndim), axis=1).astype(dims.dtype)
    if ndim == 2:
        cn = cn[[0, 1, 3, 2]]
    elif ndim == 3:
        cn = cn[[0, 1, 3, 2, 4, 5, 7, 6]]
    cn = cn - np.array(origin, dtype=dims.dtype)
    return dims.reshape([-1, 1, ndim]) * cn.reshape([1, 2 ** ndim, ndim])


def rotation_3d_in_axis_np(points, angles, axis):
    """numpy in / out, float32 torch inside — what ``@array_converter`` makes of numpy inputs."""
    p = torch.from_numpy(np.ascontiguousarray(points)).float()
    a = torch.from_numpy(np.ascontiguousarray(angles)).float()
    s, c = torch.sin(a), torch.cos(a)
    one, zero = torch.ones_like(c), torch.zeros_like(c)
    if axis in (1, -2):
        m = torch.stack([torch.stack([c, zero, -s]), torch.stack([zero, one, zero]), torch.stack([s, zero, c])])
    elif axis in (2, -1):
        m = torch.stack([torch.stack([c, s, zero]), torch.stack([-s, c, zero]), torch.stack([zero, zero, one])])
    else:
        m = torch.stack([torch.stack([one, zero, zero]), torch.stack([zero, c, s]), torch.stack([zero, -s, c])])
    out = torch.einsum('aij,jka->aik', p, m) if p.shape[0] else p
    return out.numpy().astype(points.dtype)


def center_to_corner_box3d(centers, dims, angles=None, origin=(0.5, 1.0, 0.5), axis=1):
    corners = corners_nd(dims, origin=origin)
    if angles is not None:
        corners = rotation_3d_in_axis_np(corners, angles, axis)
    corners += centers.reshape([-1, 1, 3])
    return corners


_SURF_IDX = np.array([0, 1, 2, 3, 7, 6, 5, 4, 0, 3, 7, 4, 1, 5, 6, 2, 0, 4, 5, 1, 3, 2, 6, 7]).reshape(6, 4)


def corner_to_surfaces_3d(corners):
    return corners[:, _SURF_IDX]            # [N, 6, 4, 3], same layout as both reference variants


def surface_equ_3d(polygon_surfaces):
    sv = polygon_surfaces[:, :, :2, :] - polygon_surfaces[:, :, 1:3, :]
    nv = np.cross(sv[:, :, 0, :], sv[:, :, 1, :])
    d = np.einsum('aij, aij->ai', nv, polygon_surfaces[:, :, 0, :])
    return nv, -d


def points_in_convex_polygon_3d(points, polygon_surfaces, num_surfaces=None):
    """[N, 3+] points, [M, S, >=3, 3] surfaces -> bool [N, M]; inside iff n.p + d < 0 on every
    surface, evaluated left to right in the promoted dtype like the numba loop."""
    nv, d = surface_equ_3d(polygon_surfaces[:, :, :3, :])
    p = points[:, :3]
    dt = np.result_type(p.dtype, nv.dtype)
    p, nv, d = p.astype(dt), nv.astype(dt), d.astype(dt)
    M, S = nv.shape[:2]
    ret = np.ones((p.shape[0], M), dtype=bool)
    if num_surfaces is None:
        num_surfaces = np.full((M,), 9999999, dtype=np.int64)
    for j in range(M):
        alive = np.ones((p.shape[0],), dtype=bool)
        for k in range(S):
            if k > num_surfaces[j]:
                break
            sign = ((p[:, 0] * nv[j, k, 0] + p[:, 1] * nv[j, k, 1]) + p[:, 2] * nv[j, k, 2]) + d[j, k]
            alive &= ~(sign >= 0)
        ret[:, j] = alive
    return ret


def points_in_rbbox(points, rbbox, z_axis=2, origin=(0.5, 0.5, 0)):
    corners = center_to_corner_box3d(rbbox[:, :3], rbbox[:, 3:6], rbbox[:, 6], origin=origin, axis=z_axis)
    return points_in_convex_polygon_3d(points[:, :3], corner_to_surfaces_3d(corners))


def projection_matrix_to_CRT_kitti(proj):
    CR, CT = proj[0:3, 0:3], proj[0:3, 3]
    Rinv, Cinv = np.linalg.qr(np.linalg.inv(CR))
    return np.linalg.inv(Cinv), np.linalg.inv(Rinv), Cinv @ CT


def get_frustum(bbox_image, C, near_clip=0.001, far_clip=100):
    fku, fkv = C[0, 0], -C[1, 1]
    u0v0 = C[0:2, 2]
    z = np.array([near_clip] * 4 + [far_clip] * 4, dtype=C.dtype)[:, np.newaxis]
    b = bbox_image
    bc = np.array([[b[0], b[1]], [b[0], b[3]], [b[2], b[3]], [b[2], b[1]]], dtype=C.dtype)
    near = (bc - u0v0) / np.array([fku / near_clip, -fkv / near_clip], dtype=C.dtype)
    far = (bc - u0v0) / np.array([fku / far_clip, -fkv / far_clip], dtype=C.dtype)
    return np.concatenate([np.concatenate([near, far], axis=0), z], axis=1)


def camera_to_lidar(points, r_rect, velo2cam):
    shp = list(points.shape[0:-1])
    if points.shape[-1] == 3:
        points = np.concatenate([points, np.ones(shp + [1])], axis=-1)
    return (points @ np.linalg.inv((r_rect @ velo2cam).T))[..., :3]


def frustum_surfaces(rect, Trv2c, P2, bbox_shape):
    C, R, T = projection_matrix_to_CRT_kitti(P2)
    fr = get_frustum(np.asarray(bbox_shape).tolist(), C)
    fr -= T
    fr = np.linalg.inv(R) @ fr.T
    fr = camera_to_lidar(fr.T, rect, Trv2c)
    return corner_to_surfaces_3d(fr[np.newaxis, ...])


def points_in_frustum(points, rect, Trv2c, P2, bbox_shape):
    return points_in_convex_polygon_3d(points[:, :3], frustum_surfaces(rect, Trv2c, P2, bbox_shape))


def face_distances(points, boxes):
    """points [N, 3], boxes [M, 7] (gravity centre, dims, yaw) torch fp32 -> [N, M, 6]
    (dx_min, dx_max, dy_min, dy_max, dz_min, dz_max), fcaf3d_head.py:495-520 op for op."""
    n, m = points.shape[0], boxes.shape[0]
    P = points.unsqueeze(1).expand(n, m, 3)
    B = boxes.unsqueeze(0).expand(n, m, 7)
    shift = torch.stack((P[..., 0] - B[..., 0], P[..., 1] - B[..., 1], P[..., 2] - B[..., 2]), dim=-1).permute(1, 0, 2)
    ang = -B[0, :, 6]
    s, c = torch.sin(ang), torch.cos(ang)
    one, zero = torch.ones_like(c), torch.zeros_like(c)
    mt = torch.stack([torch.stack([c, s, zero]), torch.stack([-s, c, zero]), torch.stack([zero, zero, one])])
    shift = torch.einsum('aij,jka->aik', shift, mt).permute(1, 0, 2)
    cen = B[..., :3] + shift
    return torch.stack((cen[..., 0] - B[..., 0] + B[..., 3] / 2, B[..., 0] + B[..., 3] / 2 - cen[..., 0],
                        cen[..., 1] - B[..., 1] + B[..., 4] / 2, B[..., 1] + B[..., 4] / 2 - cen[..., 1],
                        cen[..., 2] - B[..., 2] + B[..., 5] / 2, B[..., 2] + B[..., 5] / 2 - cen[..., 2]), dim=-1)
