"""TEST INFRASTRUCTURE — CPU oracle for the membership contract (part 1).

Two independent restatements of the un-vendored ``mmcv.ops.points_in_boxes_cpu``
(arithmetic: SURVEY.md Appendix A.1; reference call sites
``mmdet3d/core/bbox/structures/base_box3d.py:534,566``, re-export
``mmdet3d/ops/__init__.py:12-13``):

* ``points_in_boxes_cpu`` / ``_all`` / ``_part`` — thin ctypes wrappers over
  ``oracle/pib_oracle.c`` (the literal box-major C loop, libm trig per pair);
* ``points_in_boxes_numpy`` — a vectorised numpy restatement of the same fp32/fp64
  arithmetic, used to cross-check the C file.

Pinned against ``/root/reference/tests/test_utils/test_box3d.py:1683-1797`` in
``tests/test_oracle_membership.py``.  Never imported by ``gga_b200/``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compiles oracle/_build/liboracle.so with the committed Makefile."""
    so = os.path.join(_HERE, '_build', 'liboracle.so')
    srcs = [os.path.join(_HERE, f) for f in ('pib_oracle.c', 'detmath_host.c', 'Makefile')]
    srcs.append(os.path.join(_HERE, '..', 'include', 'gga_detmath.h'))
    stale = force or not os.path.isfile(so) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.isfile(s))
    if stale:
        subprocess.run(['make', '-C', _HERE, '_build/liboracle.so'], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.gga_oracle_points_in_boxes_cpu.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int,
                                                     ctypes.c_int, ip, ctypes.c_int]
        L.gga_oracle_points_in_boxes_cpu.restype = None
        L.gga_oracle_points_in_boxes_part.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int,
                                                      ctypes.c_int, ip]
        L.gga_oracle_points_in_boxes_part.restype = None
        L.gga_oracle_box_sincos.argtypes = [fp, ctypes.c_int, fp, fp]
        L.gga_oracle_box_sincos.restype = None
        L.gga_oracle_det_sincos.argtypes = [fp, ctypes.c_long, fp, fp]
        L.gga_oracle_det_sincos.restype = None
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def points_in_boxes_all_np(points, boxes, nthreads=1):
    """points [M, >=3] (only xyz used), boxes [T, 7] -> int32 [M, T] (one frame)."""
    points, boxes = _f32(points), _f32(boxes)
    M, T = points.shape[0], boxes.shape[0]
    out = np.zeros((T, M), dtype=np.int32)
    if M and T:
        lib().gga_oracle_points_in_boxes_cpu(
            _fptr(boxes), _fptr(points), points.shape[1], T, M,
            out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), int(nthreads))
    return np.ascontiguousarray(out.T)  # the mmcv wrapper's .transpose(1, 2)


def points_in_boxes_part_np(points, boxes):
    points, boxes = _f32(points), _f32(boxes)
    M, T = points.shape[0], boxes.shape[0]
    out = np.full((M,), -1, dtype=np.int32)
    if M and T:
        lib().gga_oracle_points_in_boxes_part(
            _fptr(boxes), _fptr(points), points.shape[1], T, M,
            out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return out


def points_in_boxes_cpu(points, boxes, nthreads=1):
    """mmcv signature: points [B, M, 3], boxes [B, T, 7] -> int32 [B, M, T] (torch in/out)."""
    import torch
    assert boxes.shape[0] == points.shape[0] and boxes.shape[2] == 7 and points.shape[2] == 3
    p, b = points.detach().cpu().float().numpy(), boxes.detach().cpu().float().numpy()
    out = np.stack([points_in_boxes_all_np(p[i], b[i], nthreads) for i in range(p.shape[0])]) \
        if p.shape[0] else np.zeros((0, p.shape[1], b.shape[1]), np.int32)
    return torch.from_numpy(out)


def points_in_boxes_all(points, boxes):
    return points_in_boxes_cpu(points, boxes)


def points_in_boxes_part(points, boxes):
    import torch
    assert boxes.shape[0] == points.shape[0] and boxes.shape[2] == 7 and points.shape[2] == 3
    p, b = points.detach().cpu().float().numpy(), boxes.detach().cpu().float().numpy()
    out = np.stack([points_in_boxes_part_np(p[i], b[i]) for i in range(p.shape[0])]) \
        if p.shape[0] else np.zeros((0, p.shape[1]), np.int32)
    return torch.from_numpy(out)


def pack_bits(mask):
    """int/bool [M, T] -> uint32 [M, ceil(T/32)], bit (t & 31) of word (t >> 5)."""
    mask = np.asarray(mask) != 0
    M, T = mask.shape
    W = (T + 31) // 32
    pad = np.zeros((M, W * 32), dtype=bool)
    pad[:, :T] = mask
    w = pad.reshape(M, W, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)
    return w.sum(-1).astype(np.uint32)


def points_in_boxes_numpy(points, boxes):
    """Vectorised numpy restatement of Appendix A.1 (cross-check of pib_oracle.c)."""
    p, b = _f32(points), _f32(boxes)
    px, py, pz = p[:, None, 0], p[:, None, 1], p[:, None, 2]
    cx, cy, z, dx, dy, dz, rz = (b[None, :, i] for i in range(7))
    half_z = dz.astype(np.float64) / 2.0
    cz = (z.astype(np.float64) + half_z).astype(np.float32)
    with np.errstate(invalid='ignore', over='ignore'):
        zpass = ~(np.abs(pz - cz).astype(np.float64) > half_z)
        cosa = np.cos((-rz).astype(np.float64)).astype(np.float32)
        sina = np.sin((-rz).astype(np.float64)).astype(np.float32)
        sx, sy = px - cx, py - cy
        lx = (sx * cosa).astype(np.float32) + (sy * (-sina)).astype(np.float32)
        ly = (sx * sina).astype(np.float32) + (sy * cosa).astype(np.float32)
        hx, hy = dx.astype(np.float64) / 2.0, dy.astype(np.float64) / 2.0
        inside = zpass & (lx > -hx) & (lx < hx) & (ly > -hy) & (ly < hy)
    return inside.astype(np.int32)


def det_sincos(x):
    """Host build of include/gga_detmath.h: (sin, cos) as fp32 arrays."""
    x = _f32(x).ravel()
    s, c = np.empty_like(x), np.empty_like(x)
    lib().gga_oracle_det_sincos(_fptr(x), x.size, _fptr(s), _fptr(c))
    return s, c


def libm_box_sincos(rz):
    """(cosa, sina) of the contract evaluated with the host libm."""
    rz = _f32(rz).ravel()
    c, s = np.empty_like(rz), np.empty_like(rz)
    lib().gga_oracle_box_sincos(_fptr(rz), rz.size, _fptr(c), _fptr(s))
    return c, s
