"""GPU parity tests of the numpy-side membership family (contract a6: points_in_rbbox /
points_in_convex_polygon_3d_jit / frustum membership) and of FCAF3D's face distances (a7)
against outputs of the reference itself (tests/golden/ref_convex.npz, ref_rbbox.npz) and the
oracle.  Masks bit-exact; distances within 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

import gga_b200 as G
from oracle import convex as oc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def C(golden_dir):
    return np.load(os.path.join(golden_dir, 'ref_convex.npz'))


def test_points_in_rbbox_vs_reference_all_dtypes(C, golden_dir):
    p, b = C['pts'], C['boxes']
    assert np.array_equal(G.points_in_rbbox(p, b), C['rbbox_f32'])
    assert np.array_equal(G.points_in_rbbox(p, b.astype(np.float64)), C['rbbox_f64boxes'])
    assert np.array_equal(G.points_in_rbbox(p.astype(np.float64), b.astype(np.float64)), C['rbbox_f64all'])
    assert np.array_equal(G.points_in_rbbox(p, C['boxes_cam'], z_axis=1, origin=(0.5, 1.0, 0.5)), C['rbbox_cam_axis1'])
    r = np.load(os.path.join(golden_dir, 'ref_rbbox.npz'))
    out = G.points_in_rbbox(r['pts'], r['boxes'])
    assert out.dtype == np.bool_ and np.array_equal(out, r['points_in_rbbox'])
    # CUDA tensor in -> CUDA bool tensor out
    t = G.points_in_rbbox(torch.from_numpy(p).cuda(), b)
    assert t.is_cuda and t.dtype == torch.bool and np.array_equal(t.cpu().numpy(), C['rbbox_f32'])


def test_contract_differs_from_mmcv_on_the_z_faces():
    """The reference's own vectors (tests/test_utils/test_box3d.py:1689-1690): points on the top
    face are inside for the mmcv op (closed z slab) and outside for points_in_rbbox (open)."""
    box = np.float32([[0.0, 0.0, 0.0, 2.0, 2.0, 2.0, 0.0]])
    pts = np.float32([[0.0, 0.0, 2.0, 0], [0.0, 0.0, 0.0, 0], [0.0, 0.0, 1.0, 0]])
    a1 = G.points_in_boxes_all(torch.from_numpy(pts[None, :, :3]).cuda(), torch.from_numpy(box[None]).cuda())[0, :, 0]
    a6 = G.points_in_rbbox(pts, box)[:, 0]
    assert a1.cpu().tolist() == [1, 1, 1] and a6.tolist() == [False, False, True]


def test_frustum_membership_vs_reference(C):
    for i, bb in enumerate(C['fr_bboxes']):
        got = G.points_in_frustm_indices(C['fr_pts'], C['fr_rect'], C['fr_Trv2c'], C['fr_P2'], bb)
        assert got.shape == (4000, 1) and np.array_equal(got[:, 0], C['fr_indices'][:, i])
    # given surfaces, num_surfaces quirk and many polygons / surfaces per polygon
    surf = np.concatenate([C['fr_surfaces']] * 30, 0)                       # 90 polygons: several smem tiles
    ns = np.full((90,), 9999999, np.int64); ns[1] = 2
    ref = oc.points_in_convex_polygon_3d(C['fr_pts'][:, :3], surf, ns)
    assert np.array_equal(G.points_in_convex_polygon_3d_jit(C['fr_pts'][:, :3], surf, ns), ref)


def test_convex_random_vs_oracle_and_nan_semantics():
    rng = np.random.default_rng(5)
    boxes = np.concatenate([rng.uniform(-20, 20, (300, 3)), rng.uniform(0.5, 6, (300, 3)), rng.uniform(-7, 7, (300, 1))], 1).astype(np.float32)
    pts = np.concatenate([boxes[rng.integers(0, 300, 20000), :3] + rng.normal(0, 1.5, (20000, 3)),
                          rng.uniform(0, 1, (20000, 1))], 1).astype(np.float32)
    pts[0, 0] = np.nan
    ref = oc.points_in_rbbox(pts, boxes)
    got = G.points_in_rbbox(pts, boxes)
    assert np.array_equal(got, ref) and ref.sum() > 5000
    assert got[0].all()            # NaN point: `sign >= 0` is false for every surface (reference semantics)
    assert G.points_in_rbbox(pts[:0], boxes).shape == (0, 300) and G.points_in_rbbox(pts, boxes[:0]).shape == (20000, 0)


def test_face_distances_vs_reference(C):
    p, b = torch.from_numpy(C['fd_pts']).cuda(), torch.from_numpy(C['fd_boxes']).cuda()
    fd, inside = G.face_distances(p, b, return_inside=True)
    ref = C['fd_dist']
    assert np.allclose(fd.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    clear = np.abs(ref.min(-1)) > 1e-5
    assert np.array_equal(inside.cpu().numpy()[clear], C['fd_inside'][clear])
    # SUN-RGBD-sized: 50k points x 64 boxes against the oracle
    rng = np.random.default_rng(1)
    bb = np.concatenate([rng.uniform(-4, 4, (64, 2)), rng.uniform(-1.5, 0.5, (64, 1)), rng.uniform(0.3, 2.5, (64, 3)),
                         rng.uniform(-np.pi, np.pi, (64, 1))], 1).astype(np.float32)
    pp = np.stack([rng.uniform(-5, 5, 50000), rng.uniform(-5, 5, 50000), rng.uniform(-2, 1, 50000)], 1).astype(np.float32)
    r = oc.face_distances(torch.from_numpy(pp), torch.from_numpy(bb)).numpy()
    g, gi = G.face_distances(torch.from_numpy(pp).cuda(), torch.from_numpy(bb).cuda(), return_inside=True)
    assert np.allclose(g.cpu().numpy(), r, rtol=1e-5, atol=1e-5)
    clear = np.abs(r.min(-1)) > 1e-5
    assert np.array_equal(gi.cpu().numpy()[clear], (r.min(-1) > 0)[clear])
