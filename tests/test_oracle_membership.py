"""Pins the membership oracle (oracle/pib_oracle.c, oracle/membership.py) against the
reference's own golden masks: /root/reference/tests/test_utils/test_box3d.py:1683-1797."""
import numpy as np
import pytest
import torch

from oracle import membership as om

LIDAR_PTS = [[1.0, 4.3, 0.1], [1.0, 4.4, 0.1], [1.1, 4.3, 0.1], [0.9, 4.3, 0.1], [1.0, -0.3, 0.1],
             [1.0, -0.4, 0.1], [2.9, 0.1, 6.0], [-0.9, 3.9, 6.0]]
LIDAR_BOXES = [[1.0, 2.0, 0.0, 4.0, 4.0, 6.0, np.pi / 6], [1.0, 2.0, 0.0, 4.0, 4.0, 6.0, np.pi / 2],
               [1.0, 2.0, 0.0, 4.0, 4.0, 6.0, 7 * np.pi / 6], [1.0, 2.0, 0.0, 4.0, 4.0, 6.0, -np.pi / 6]]
LIDAR_ALL = [[1, 0, 1, 1], [0, 0, 0, 0], [1, 0, 1, 0], [0, 0, 0, 1], [1, 0, 1, 1], [0, 0, 0, 0],
             [0, 1, 0, 0], [0, 1, 0, 0]]                       # test_box3d.py:1699-1702
LIDAR_PART = [0, -1, 0, 3, 0, -1, 1, 1]                       # :1719
DEPTH_BOXES = [[1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 0.3], [-10.0, 23.0, 16.0, 10, 20, 20, 0.5]]
DEPTH_PTS = [[1, 2, 3.3], [1.2, 2.5, 3.0], [0.8, 2.1, 3.5], [1.6, 2.6, 3.6], [0.8, 1.2, 3.9],
             [-9.2, 21.0, 18.2], [3.8, 7.9, 6.3], [4.7, 3.5, -12.2], [3.8, 7.6, -2],
             [-10.6, -12.9, -20], [-16, -18, 9], [-21.3, -52, -5], [0, 0, 0], [6, 7, 8], [-2, -3, -4]]
DEPTH_ALL = [[1, 0]] * 5 + [[0, 1]] + [[0, 0]] * 9             # :1737-1739
DEPTH_PART = [0, 0, 0, 0, 0, 1] + [-1] * 9                      # :1745-1747
CAM_ALL = [[1, 0, 1, 1, 1, 1]] * 5 + [[0, 1, 0, 0, 0, 0]] + [[0] * 6] * 6 + [[0, 0, 0, 1, 0, 1]] + \
    [[0] * 6] * 2 + [[0, 0, 1, 1, 1, 1], [0, 0, 0, 1, 0, 0], [0, 0, 0, 1, 0, 1], [0, 0, 1, 1, 1, 0],
                     [0, 0, 1, 1, 1, 1], [0, 0, 0, 1, 0, 0], [1, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0]]
CAM_PART = [0, 0, 0, 0, 0, 1, -1, -1, -1, -1, -1, -1, 3, -1, -1, 2, 3, 3, 2, 2, 3, 0, 0]  # :1788-1791


def _f(a):
    return np.asarray(a, dtype=np.float32)


def test_lidar_golden_all_and_part():
    assert (om.points_in_boxes_all_np(_f(LIDAR_PTS), _f(LIDAR_BOXES)) == np.array(LIDAR_ALL)).all()
    assert (om.points_in_boxes_part_np(_f(LIDAR_PTS), _f(LIDAR_BOXES)) == np.array(LIDAR_PART)).all()
    assert (om.points_in_boxes_numpy(_f(LIDAR_PTS), _f(LIDAR_BOXES)) == np.array(LIDAR_ALL)).all()


def test_depth_golden_all_and_part():
    assert (om.points_in_boxes_all_np(_f(DEPTH_PTS), _f(DEPTH_BOXES)) == np.array(DEPTH_ALL)).all()
    assert (om.points_in_boxes_part_np(_f(DEPTH_PTS), _f(DEPTH_BOXES)) == np.array(DEPTH_PART)).all()


def test_torch_signature_batched():
    p = torch.tensor([LIDAR_PTS, LIDAR_PTS], dtype=torch.float32)
    b = torch.tensor([LIDAR_BOXES, LIDAR_BOXES[::-1]], dtype=torch.float32)
    out = om.points_in_boxes_cpu(p, b)
    assert out.dtype == torch.int32 and tuple(out.shape) == (2, 8, 4)
    assert (out[0].numpy() == np.array(LIDAR_ALL)).all()
    assert (out[1].numpy() == np.array(LIDAR_ALL)[:, ::-1]).all()


@pytest.mark.refonly
def test_reference_wrappers_with_oracle_injected():
    """The reference's OWN box classes (base_box3d.py:510-568, cam_box3d.py:303-354) reproduce
    every golden of test_box3d.py:1683-1797 when the oracle is injected as mmcv.ops."""
    from oracle import ref_loader
    ref = ref_loader.load_reference(om.points_in_boxes_all, om.points_in_boxes_part)
    lb = ref.LiDARInstance3DBoxes(torch.tensor(LIDAR_BOXES, dtype=torch.float32))
    pts = torch.tensor(LIDAR_PTS)
    assert (lb.points_in_boxes_all(pts).numpy() == np.array(LIDAR_ALL)).all()
    assert (lb.points_in_boxes_part(pts).numpy() == np.array(LIDAR_PART)).all()
    db = ref.DepthInstance3DBoxes(torch.tensor(DEPTH_BOXES, dtype=torch.float32))
    dpts = torch.tensor([DEPTH_PTS], dtype=torch.float32)
    assert (db.points_in_boxes_all(dpts).numpy() == np.array(DEPTH_ALL)).all()
    assert (db.points_in_boxes_part(dpts).numpy() == np.array(DEPTH_PART)).all()
    six = torch.tensor(DEPTH_BOXES + LIDAR_BOXES, dtype=torch.float32)
    cam_boxes = ref.DepthInstance3DBoxes(six).convert_to(ref.Box3DMode.CAM)
    cam_pts = ref.DepthPoints(torch.tensor(DEPTH_PTS + LIDAR_PTS, dtype=torch.float32)).convert_to(
        ref.Coord3DMode.CAM).tensor
    assert (cam_boxes.points_in_boxes_all(cam_pts).numpy() == np.array(CAM_ALL)).all()
    assert (cam_boxes.points_in_boxes_part(cam_pts).numpy() == np.array(CAM_PART)).all()


def test_c_oracle_equals_numpy_restatement_random_and_boundary():
    rng = np.random.default_rng(7)
    boxes = np.concatenate([rng.uniform(-20, 20, (40, 3)), rng.uniform(0.2, 6, (40, 3)),
                            rng.uniform(-7, 7, (40, 1))], 1).astype(np.float32)
    boxes[:8, 6] = [0, np.pi / 2, np.pi, -np.pi / 2, 0, 0, 1e-3, 2 * np.pi]
    boxes[0, :6] = [1.0, 2.0, 0.0, 4.0, 4.0, 6.0]   # exactly representable faces
    pts = rng.uniform(-25, 25, (5000, 3)).astype(np.float32)
    # points on faces / edges of the yaw-0 box 0 and on its closed z faces
    b = boxes[0]
    hx, hy = b[3] / 2, b[4] / 2
    edge = np.array([[b[0] + hx, b[1], b[2] + 0.1], [b[0] - hx, b[1], b[2] + 0.1],
                     [b[0], b[1] + hy, b[2] + 0.1], [b[0], b[1], b[2]], [b[0], b[1], b[2] + b[5]],
                     [np.nan, 0, 0], [0, np.nan, 0], [b[0], b[1], np.nan], [np.inf, 0, 0]], np.float32)
    pts = np.concatenate([pts, edge], 0)
    a = om.points_in_boxes_all_np(pts, boxes)
    assert (a == om.points_in_boxes_numpy(pts, boxes)).all()
    assert a.sum() > 0
    # closed z faces: bottom and top centre points are inside box 0
    assert a[5003, 0] == 1 and a[5004, 0] == 1
    # open x/y faces
    assert a[5000, 0] == 0 and a[5001, 0] == 0 and a[5002, 0] == 0
    # NaN x / y and inf are outside everything; a NaN *z* passes the closed-slab test
    # (`fabsf(z - cz) > dz/2` is false for NaN), so (cx, cy, NaN) is INSIDE — contract quirk.
    assert a[5005].sum() == 0 and a[5006].sum() == 0 and a[5008].sum() == 0
    assert a[5007, 0] == 1
    part = om.points_in_boxes_part_np(pts, boxes)
    first = np.where(a.any(1), a.argmax(1), -1)
    assert (part == first).all()
    assert (om.points_in_boxes_all_np(pts, boxes, nthreads=4) == a).all()


def test_detmath_equals_libm_on_samples():
    """Spot-check of the exhaustive run recorded in DESIGN.md (oracle/check_sincos.c: all 2^32
    fp32 inputs, 0 mismatches against glibc 2.39)."""
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2 ** 32, 400000, dtype=np.uint64).astype(np.uint32)
    x = bits.view(np.float32)
    x = np.concatenate([x, rng.uniform(-10, 10, 200000).astype(np.float32),
                        np.float32([0, -0.0, np.pi, np.pi / 2, np.pi / 4, 0.785398, 0.7853982,
                                    1e-30, 1e30, 3.4e38, np.inf, -np.inf, np.nan])])
    s, c = om.det_sincos(x)
    with np.errstate(invalid='ignore'):
        # the contract evaluates cos(-rz), sin(-rz)
        cl, sl = om.libm_box_sincos(-x)
    assert (s.view(np.uint32) == sl.view(np.uint32))[np.isfinite(x)].all()
    assert (c.view(np.uint32) == cl.view(np.uint32))[np.isfinite(x)].all()
    assert np.isnan(s[~np.isfinite(x)]).all() and np.isnan(c[~np.isfinite(x)]).all()


def test_pack_bits():
    m = np.zeros((3, 70), np.int32)
    m[0, 0] = m[1, 31] = m[1, 32] = m[2, 69] = 1
    w = om.pack_bits(m)
    assert w.shape == (3, 3) and w.dtype == np.uint32
    assert w[0, 0] == 1 and w[1, 0] == 1 << 31 and w[1, 1] == 1 and w[2, 2] == 1 << 5
