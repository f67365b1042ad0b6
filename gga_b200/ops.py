"""Point -> 3D-box membership ops with the mmcv signatures.

Drop-in for ``from mmcv.ops import points_in_boxes_all, points_in_boxes_cpu,
points_in_boxes_part`` as re-exported by the reference at
``/root/reference/mmdet3d/ops/__init__.py:12-13`` and bound by its box classes at
``/root/reference/mmdet3d/core/bbox/structures/base_box3d.py:7`` (calls ``:534-536,566``).
Same names, argument order, shapes, dtypes, output layout and default values; shape
mismatches raise ``AssertionError`` like the mmcv Python wrappers, library failures raise
``RuntimeError`` like mmcv's ``TORCH_CHECK``.  The work runs on the current torch CUDA
stream of the points' device, asynchronously.

All three functions implement ONE contract, the CPU one, bit for
bit — unlike mmcv, whose CUDA and CPU kernels differ in the last ulp of the rotation.
There is no CPU fallback: ``points_in_boxes_cpu`` accepts and returns host tensors but
computes on the GPU through ``gga_points_in_boxes_all_host``.
"""
import torch

from . import _lib


def _check_shapes(points, boxes):
    assert points.dim() == 3 and boxes.dim() == 3, \
        f'points and boxes must be [B, M, 3] and [B, T, 7], got {tuple(points.shape)} and {tuple(boxes.shape)}'
    assert points.shape[0] == boxes.shape[0], \
        f'Points and boxes should have the same batch size, but got {points.shape[0]} and {boxes.shape[0]}'
    assert boxes.shape[2] == 7, f'boxes dimension should be 7, but got unexpected shape {boxes.shape[2]}'


def _as_f32_rows(t):
    """Returns (tensor, row_stride_in_floats) without copying when `t` is a last-dim slice
    of a contiguous fp32 tensor (the `points[..., :3]` view base_box3d.py:559 builds)."""
    if t.dtype != torch.float32:
        t = t.float()
    B, M, C = t.shape
    if M == 0 or B == 0:
        return t.contiguous(), max(C, 3)
    s0, s1, s2 = t.stride()
    if s2 == 1 and s1 >= C and s0 == s1 * M and (t.data_ptr() % 4 == 0):
        return t, s1
    if B == 1 and s2 == 1 and s1 >= C:
        return t, s1
    t = t.contiguous()
    return t, C


def _device_check(points, boxes):
    assert points.is_cuda and boxes.is_cuda, 'points and boxes must be CUDA tensors'
    assert points.device == boxes.device, 'Points and boxes should be put on the same device'


def points_in_boxes_bits(points, boxes):
    """Bit-packed membership (this library's native layout).

    Args:
        points (Tensor): [B, M, C>=3] fp32, xyz first (C = 4 for KITTI x,y,z,r).
        boxes (Tensor): [B, T, 7] (x, y, z_bottom, dx, dy, dz, rz), LiDAR/depth coordinates.
    Returns:
        Tensor: int32 [B, M, W]; box t of a point is bit (t & 31) of word (t >> 5);
        W = ``row_words(T)`` (1/2/4 for T <= 32/64/128, else 8*ceil(T/256)).
    """
    _check_shapes(points, boxes)
    assert points.shape[2] >= 3, f'points dimension should be >= 3, but got {points.shape[2]}'
    _device_check(points, boxes)
    L = _lib.load()
    B, M, _ = points.shape
    T = boxes.shape[1]
    W = L.gga_pib_row_words(T)
    pts, stride = _as_f32_rows(points)
    bx = boxes.float().contiguous()
    out = torch.empty((B, M, W), dtype=torch.int32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(L.gga_points_in_boxes_bits(_lib.ptr(pts), stride, _lib.ptr(bx), _lib.ptr(out), B, M, T,
                                              _lib.current_stream(points.device)), 'points_in_boxes_bits')
    return out


def hit_list(bits, num_boxes, capacity=None):
    """(row, box) pairs of the set bits of bit-packed rows.

    Args:
        bits (Tensor): int32 [..., W] as returned by ``points_in_boxes_bits`` (rows are numbered in
            memory order over the leading dimensions, e.g. ``frame * M + point``).
        num_boxes (int): T.
        capacity (int): pairs to make room for (default: 4 per row, at least 1024).
    Returns:
        Tensor: int32 [n_hits, 2] = (row, box); the order is unspecified.  Synchronises once (the
        count is read back); raises if the list did not fit ``capacity``.
    """
    assert bits.is_cuda and bits.dtype == torch.int32 and bits.is_contiguous()
    L = _lib.load()
    W = bits.shape[-1]
    assert W == L.gga_pib_row_words(int(num_boxes)), 'row width does not match num_boxes'
    rows = bits.numel() // max(W, 1)
    cap = int(capacity) if capacity is not None else max(1024, 4 * rows)
    pairs = torch.empty((cap, 2), dtype=torch.int32, device=bits.device)
    count = torch.empty((1,), dtype=torch.int32, device=bits.device)
    with torch.cuda.device(bits.device):
        _lib.check(L.gga_pib_hit_list(bits.data_ptr(), rows, int(num_boxes), 0, pairs.data_ptr(), cap, count.data_ptr(),
                                      1, _lib.current_stream(bits.device)), 'pib_hit_list')
    n = int(count.item())
    if n > cap:
        raise RuntimeError(f'hit list overflow: {n} pairs, capacity {cap}')
    return pairs[:n]


def row_words(num_boxes):
    return _lib.load().gga_pib_row_words(int(num_boxes))


def unpack_bits(bits, num_boxes):
    """int32 [..., W] -> int32 [..., T] 0/1 (test / debugging helper, plain torch)."""
    t = torch.arange(num_boxes, device=bits.device)
    words = bits[..., (t >> 5)]
    return ((words >> (t & 31)) & 1).to(torch.int32)


def points_in_boxes_all(points, boxes):
    """Find all boxes in which each point is (CUDA).  mmcv signature.

    Args:
        points (Tensor): [B, M, 3], [x, y, z] in LiDAR/DEPTH coordinate.
        boxes (Tensor): [B, T, 7], num_valid_boxes <= T,
            [x, y, z, x_size, y_size, z_size, rz], (x, y, z) is the bottom centre.
    Returns:
        Tensor: int32 [B, M, T], 1 where point m is in box t, else 0.
    """
    _check_shapes(points, boxes)
    assert points.shape[2] == 3, f'points dimension should be 3, but got unexpected shape {points.shape[2]}'
    _device_check(points, boxes)
    L = _lib.load()
    B, M, _ = points.shape
    T = boxes.shape[1]
    pts, stride = _as_f32_rows(points)
    bx = boxes.float().contiguous()
    out = torch.empty((B, M, T), dtype=torch.int32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(L.gga_points_in_boxes_all(_lib.ptr(pts), stride, _lib.ptr(bx), _lib.ptr(out), B, M, T,
                                             _lib.current_stream(points.device)), 'points_in_boxes_all')
    return out


def points_in_boxes_part(points, boxes):
    """Find the box in which each point is (CUDA).  mmcv signature.

    Returns:
        Tensor: int32 [B, M]; index of the first enclosing box, -1 if none.
    """
    _check_shapes(points, boxes)
    assert points.shape[2] == 3, f'points dimension should be 3, but got unexpected shape {points.shape[2]}'
    _device_check(points, boxes)
    L = _lib.load()
    B, M, _ = points.shape
    T = boxes.shape[1]
    pts, stride = _as_f32_rows(points)
    bx = boxes.float().contiguous()
    out = torch.empty((B, M), dtype=torch.int32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(L.gga_points_in_boxes_part(_lib.ptr(pts), stride, _lib.ptr(bx), _lib.ptr(out), B, M, T,
                                              _lib.current_stream(points.device)), 'points_in_boxes_part')
    return out


def points_in_boxes_cpu(points, boxes):
    """mmcv ``points_in_boxes_cpu`` signature: HOST tensors in, host int32 [B, M, T] out.

    The computation itself runs on the current CUDA device (H2D, kernel, D2H inside the
    C call); results are bit-identical to the CPU op's contract.
    """
    _check_shapes(points, boxes)
    assert points.shape[2] == 3, f'points dimension should be 3, but got unexpected shape {points.shape[2]}'
    assert not points.is_cuda and not boxes.is_cuda, 'points_in_boxes_cpu takes CPU tensors'
    L = _lib.load()
    B, M, _ = points.shape
    T = boxes.shape[1]
    pts = points.float().contiguous()
    bx = boxes.float().contiguous()
    out = torch.zeros((B, M, T), dtype=torch.int32)
    _lib.check(L.gga_points_in_boxes_all_host(_lib.ptr(pts), 3, _lib.ptr(bx), _lib.ptr(out), B, M, T),
               'points_in_boxes_cpu')
    return out
