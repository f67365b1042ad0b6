"""Pins oracle/geometry.py and oracle/losses.py against (a) known-answer vectors copied from
the reference's own tests and (b) fixtures produced by running the reference itself
(oracle/gen_golden.py -> tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import geometry as og
from oracle import losses as ol


@pytest.fixture(scope='module')
def G(golden_dir):
    return np.load(os.path.join(golden_dir, 'ref_geometry.npz'))


@pytest.fixture(scope='module')
def I(golden_dir):
    return np.load(os.path.join(golden_dir, 'ref_iou.npz'))


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(a, b, rtol=1e-5, atol=1e-5):
    a, b = np.asarray(a), np.asarray(b)
    return np.allclose(a, b, rtol=rtol, atol=atol)


def test_lidar_corners_known_answer():
    # /root/reference/tests/test_utils/test_box3d.py:385-408 (boxes after translate), :482
    # (limit_yaw) -> :504-545 (corners, rtol 1e-4)
    boxes = torch.tensor([
        [1.1281544, -3.0507944, -1.9169292, 1.7597977, 3.4089797, 1.6592377, 1.9336663 - np.pi],
        [8.098079, -4.9332013, -1.8018866, 1.5486219, 4.0324507, 1.57879, 1.7936664 - np.pi],
        [27.64241, -7.2408795, -1.4676381, 1.4782301, 2.242485, 1.488286, 4.9836664 - np.pi],
        [20.018322, -28.477297, -1.9027928, 1.5687338, 3.4994833, 1.4078381, 5.1036663 - np.pi],
        [28.21472, -16.502048, -1.7878747, 1.7497417, 3.791107, 1.488286, 0.6236664 - np.pi]])
    boxes[:, 6] = og.limit_period(boxes[:, 6], 0.5, np.pi)
    c = og.corners_lidar(boxes)
    exp0 = torch.tensor([[-7.7767e-01, -2.8332e+00, -1.9169e+00], [-7.7767e-01, -2.8332e+00, -2.5769e-01],
                         [2.4093e+00, -1.6232e+00, -2.5769e-01], [2.4093e+00, -1.6232e+00, -1.9169e+00],
                         [-1.5301e-01, -4.4784e+00, -1.9169e+00], [-1.5301e-01, -4.4784e+00, -2.5769e-01],
                         [3.0340e+00, -3.2684e+00, -2.5769e-01], [3.0340e+00, -3.2684e+00, -1.9169e+00]])
    exp4 = torch.tensor([[2.8612e+01, -1.8552e+01, -1.7879e+00], [2.8612e+01, -1.8552e+01, -2.9959e-01],
                         [2.6398e+01, -1.5474e+01, -2.9959e-01], [2.6398e+01, -1.5474e+01, -1.7879e+00],
                         [3.0032e+01, -1.7530e+01, -1.7879e+00], [3.0032e+01, -1.7530e+01, -2.9959e-01],
                         [2.7818e+01, -1.4452e+01, -2.9959e-01], [2.7818e+01, -1.4452e+01, -1.7879e+00]])
    assert torch.allclose(c[0], exp0, rtol=1e-4, atol=1e-4)
    assert torch.allclose(c[4], exp4, rtol=1e-4, atol=1e-4)


def test_points_cam2img_known_answer():
    # /root/reference/tests/test_utils/test_box3d.py:1653-1661
    torch.manual_seed(0)
    points = torch.rand([5, 3])
    proj = torch.rand([4, 4])
    exp = torch.tensor([[0.5832, 0.6496], [0.6146, 0.7910], [0.6994, 0.7782], [0.5623, 0.6303],
                        [0.4359, 0.6532]])
    assert torch.allclose(og.points_cam2img(points, proj), exp, 1e-3)


def test_kitti_box_known_answer(G):
    # tests/test_data/test_datasets/test_kitti_dataset.py:378-379 -> :393
    box = torch.tensor([[8.7314, -1.8559, -1.5997, 1.2000, 0.4800, 1.8900, -1.5808]])
    b2d, valid, cam = og.project_kitti_cam(box, T(G['rect']), T(G['Trv2c']), T(G['P2']), (375, 1242),
                                           [0, -40, -3, 70.4, 40, 0.0])
    exp = np.array([[710.443, 144.00221, 820.29114, 307.58667]], np.float32)
    assert np.allclose(b2d.numpy(), exp, rtol=1e-5, atol=1e-3) and bool(valid[0])
    # :395 box3d_camera location
    assert np.allclose(cam[0, :3].numpy(), [1.8399826, 1.4700009, 8.410018], atol=1e-4)
    # variant A on the same box is a DIFFERENT function
    l2i = (G['P2'] @ G['rect'] @ G['Trv2c']).astype(np.float32)
    a = og.project_lidar_direct(box, T(l2i)[None])
    assert np.allclose(a.numpy(), [[708.5656, 143.4244, 820.2964, 308.1900]], atol=2e-2)


def test_corners_rotation_limit_period_vs_reference(G):
    b = T(G['boxes_lidar'])
    assert close(og.corners_lidar(b), G['corners_lidar'])
    assert close(og.corners_lidar(b), G['corners_depth'])
    assert close(og.corners_cam(b), G['corners_cam'])
    assert close(og.limit_period(T(G['lp_in'])), G['lp_pi'], atol=1e-6)
    assert close(og.limit_period(T(G['lp_in']), 0.5, np.pi * 2), G['lp_2pi'], atol=1e-6)
    p, a = T(G['rot_pts']), T(G['lp_in'])
    assert close(og.rotation_3d_in_axis(p, a, 2), G['rot_axis2'], atol=1e-6)
    assert close(og.rotation_3d_in_axis(p, a, 1), G['rot_axis1'], atol=1e-6)
    assert close(og.rotation_3d_in_axis(p, a, 2, clockwise=True), G['rot_axis2_cw'], atol=1e-6)


def test_convert_and_cam2img_vs_reference(G):
    rt = T(G['rect'] @ G['Trv2c'])
    assert close(og.lidar_to_cam_boxes(T(G['boxes_lidar']), rt), G['boxes_cam'])
    assert close(og.points_cam2img(T(G['c2i_pts']), T(G['P2'])), G['c2i_uv'], atol=1e-3)
    assert close(og.points_cam2img(T(G['c2i_pts']), T(G['P2'][:3])), G['c2i_uv_3x4'], atol=1e-3)


def test_variant_b_vs_reference(G):
    b2d, valid, _ = og.project_kitti_cam(T(G['boxes_lidar']), T(G['rect']), T(G['Trv2c']), T(G['P2']),
                                         G['img_hw'], G['pcd_range'])
    raw, _, _ = og.project_kitti_cam(T(G['boxes_lidar']), T(G['rect']), T(G['Trv2c']), T(G['P2']),
                                     G['img_hw'], G['pcd_range'], clamp=False)
    assert close(raw, G['varB_box2d_raw'], rtol=1e-5, atol=2e-3)
    assert close(b2d, G['varB_box2d_clamped'], rtol=1e-5, atol=2e-3)
    assert (valid.numpy() == G['varB_valid']).all()


def test_variant_a_c_forward_and_grad_vs_reference(G):
    b = T(G['boxes_lidar']).clone().requires_grad_(True)
    out = og.project_lidar_direct(b, T(G['varA_lidar2img']))
    assert close(out.detach(), G['varA_box2d'], rtol=1e-5, atol=2e-3)
    out.backward(T(G['varA_gout']))
    assert np.allclose(b.grad.numpy(), G['varA_grad_boxes'], rtol=1e-3, atol=1e-2)
    c = T(G['varC_boxes_cam_center']).clone().requires_grad_(True)
    outc = og.project_cam(c, T(G['P2']))
    assert close(outc.detach(), G['varC_box2d'], rtol=1e-5, atol=2e-3)
    outc.backward(T(G['varA_gout']))
    assert np.allclose(c.grad.numpy(), G['varC_grad_boxes'], rtol=1e-3, atol=1e-2)


def test_axis_aligned_iou_loss_known_answer():
    # /root/reference/tests/test_metrics/test_losses.py:178-189
    pred = torch.tensor([[0., 0., 0., 1., 1., 1.], [0., 0., 0., 1., 1., 1.]])
    target = torch.tensor([[0., 0., 0., 1., 1., 1.], [2., 2., 2., 3., 3., 3.]])
    # golden table of the reference's loss: identical -> 0, disjoint -> 1
    l = ol.axis_aligned_iou_loss(pred, target, reduction='none')
    assert torch.allclose(l, torch.tensor([0., 1.]))
    b1 = torch.tensor([[0., 0., 0., 2., 2., 2.]])
    b2 = torch.tensor([[1., 1., 1., 3., 3., 3.]])
    assert torch.allclose(ol.axis_aligned_iou_loss(b1, b2, reduction='none'), torch.tensor([14. / 15.]))


def test_iou_giou_vs_reference_formula(I):
    b1, b2 = T(I['b1'].astype(np.float32)), T(I['b2'].astype(np.float32))
    assert close(ol.bbox_overlaps_aligned(b1, b2, 'iou'), I['aa3d_iou'], atol=1e-6)
    assert close(ol.bbox_overlaps_aligned(b1, b2, 'giou'), I['aa3d_giou'], atol=1e-6)
    w = T(I['w'])
    p = b1.clone().requires_grad_(True)
    ol.giou_loss_module(p, b2, w, reduction='sum').backward()
    assert np.allclose(p.grad.numpy(), I['aa3d_giou_loss_grad_b1'], rtol=1e-4, atol=1e-6)
    q1, q2 = T(I['q1']), T(I['q2'])
    assert close(ol.axis_aligned_overlaps_3d_aligned(q1, q2, 'iou'), I['aa3d_iou_3d'], atol=1e-6)
    assert close(ol.axis_aligned_overlaps_3d_aligned(q1, q2, 'giou'), I['aa3d_giou_3d'], atol=1e-6)


def test_giou_cross_check_torchvision(I):
    tv = pytest.importorskip('torchvision.ops')
    b1, b2 = T(I['b1'].astype(np.float32))[24:], T(I['b2'].astype(np.float32))[24:]
    ours = 1 - ol.bbox_overlaps_aligned(b1, b2, 'giou')
    assert torch.allclose(ours, tv.generalized_box_iou_loss(b1, b2), atol=1e-5)


def test_image_box_overlap_vs_reference(I):
    assert np.array_equal(ol.image_box_overlap(I['b1'][:50], I['b2'][:30]), I['ibo_f64'])
    # zero unless strictly positive overlap
    assert I['ibo_f64'][20:24, 20:24].diagonal().max() == 0


def test_loss_module_conventions():
    pred = torch.tensor([[0., 0., 10., 10.], [0., 0., 10., 10.]], requires_grad=True)
    tgt = torch.tensor([[0., 0., 10., 10.], [5., 5., 15., 15.]])
    l = ol.giou_loss_module(pred, tgt, reduction='none')
    assert torch.allclose(l, torch.tensor([0., 1 - (25 / 175 - (225 - 175) / 225)]))
    w = torch.tensor([1.0, 0.5])
    assert torch.allclose(ol.giou_loss_module(pred, tgt, w, avg_factor=4.0, loss_weight=2.0),
                          2.0 * (l * w).sum() / 4.0)
    # weight [n,4] is averaged over the last dim; all-zero weight early-out keeps the graph
    z = ol.giou_loss_module(pred, tgt, torch.zeros(2, 4))
    assert float(z) == 0.0 and z.requires_grad
    l1 = ol.l1_loss_module(pred, tgt, torch.ones(2, 4), avg_factor=2.0, loss_weight=0.25)
    assert torch.allclose(l1, torch.tensor(0.25 * 20.0 / 2.0))
