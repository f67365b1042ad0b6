"""Oracle pins for the convex-polygon membership family and FCAF3D face distances against
outputs of the reference itself (tests/golden/ref_convex.npz, ref_rbbox.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import convex as oc


@pytest.fixture(scope='module')
def C(golden_dir):
    return np.load(os.path.join(golden_dir, 'ref_convex.npz'))


def test_surfaces_and_planes_bit_exact(C):
    b = C['boxes']
    surf = oc.corner_to_surfaces_3d(oc.center_to_corner_box3d(b[:, :3], b[:, 3:6], b[:, 6], origin=(0.5, 0.5, 0), axis=2))
    assert np.array_equal(surf, C['surfaces_f32'])
    nv, d = oc.surface_equ_3d(surf[:, :, :3, :])
    assert np.array_equal(nv, C['normal_f32']) and np.array_equal(d, C['d_f32'])


def test_points_in_rbbox_all_dtype_combinations(C, golden_dir):
    p, b = C['pts'], C['boxes']
    assert np.array_equal(oc.points_in_rbbox(p, b), C['rbbox_f32'])
    assert np.array_equal(oc.points_in_rbbox(p, b.astype(np.float64)), C['rbbox_f64boxes'])
    assert np.array_equal(oc.points_in_rbbox(p.astype(np.float64), b.astype(np.float64)), C['rbbox_f64all'])
    assert np.array_equal(oc.points_in_rbbox(p, C['boxes_cam'], z_axis=1, origin=(0.5, 1.0, 0.5)), C['rbbox_cam_axis1'])
    r = np.load(os.path.join(golden_dir, 'ref_rbbox.npz'))
    assert np.array_equal(oc.points_in_rbbox(r['pts'], r['boxes']), r['points_in_rbbox'])
    # open faces: the points placed exactly on a face of the first three boxes are outside
    n_in = 1500
    assert not C['rbbox_f32'][n_in:n_in + 6].any()


def test_frustum_membership(C):
    for i, bb in enumerate(C['fr_bboxes']):
        surf = oc.frustum_surfaces(C['fr_rect'], C['fr_Trv2c'], C['fr_P2'], bb)
        assert np.allclose(surf[0], C['fr_surfaces'][i], rtol=1e-12, atol=1e-12)
        got = oc.points_in_convex_polygon_3d(C['fr_pts'][:, :3], C['fr_surfaces'][i][None])[:, 0]
        assert np.array_equal(got, C['fr_indices'][:, i])


def test_face_distances(C):
    fd = oc.face_distances(torch.from_numpy(C['fd_pts']), torch.from_numpy(C['fd_boxes']))
    assert np.allclose(fd.numpy(), C['fd_dist'], rtol=1e-6, atol=1e-6)
    assert np.array_equal((fd.min(dim=-1).values > 0).numpy(), C['fd_inside'])
