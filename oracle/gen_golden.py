"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the REFERENCE ITSELF.

Run in the build container only (needs /root/reference):
    python oracle/gen_golden.py
It imports the reference's own geometry modules through ``oracle/ref_loader.py`` and
records their outputs (and torch-CPU autograd gradients through them) on seeded inputs.
The fixtures are committed; the tests that consume them run anywhere (the GPU box has no
/root/reference).  Every array is produced by reference code paths cited inline; where
the reference function cannot be imported (module header needs mmcv/mmdet registries)
the fixture is composed from the reference's own importable pieces exactly as that
function composes them, and says so.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

# KITTI 000000 calibration: rect / Trv2c from /root/reference/tests/test_utils/test_box_np_ops.py:8-15;
# P2 from the KITTI calib file of the same frame (P2 @ rect @ Trv2c reproduces
# expected_lidar2img of tests/test_data/test_datasets/test_kitti_dataset.py:226-230).
RECT = np.array([[0.9999128, 0.01009263, -0.00851193, 0.],
                 [-0.01012729, 0.9999406, -0.00403767, 0.],
                 [0.00847068, 0.00412352, 0.9999556, 0.],
                 [0., 0., 0., 1.]], dtype=np.float32)
TRV2C = np.array([[0.00692796, -0.9999722, -0.00275783, -0.02457729],
                  [-0.00116298, 0.00274984, -0.9999955, -0.06127237],
                  [0.9999753, 0.00693114, -0.0011439, -0.3321029],
                  [0., 0., 0., 1.]], dtype=np.float32)
P2 = np.array([[707.0493, 0., 604.0814, 45.75831],
               [0., 707.0493, 180.5066, -0.3454157],
               [0., 0., 1., 0.004981016],
               [0., 0., 0., 1.]], dtype=np.float32)
IMG_HW = (375, 1242)
PCD_RANGE = [0, -40, -3, 70.4, 40, 0.0]


def kitti_like_boxes(rng, n):
    cls = rng.integers(0, 3, n)
    mean = np.array([[3.9, 1.6, 1.56], [0.8, 0.6, 1.73], [1.76, 0.6, 1.73]], np.float32)[cls]
    dims = mean * rng.uniform(0.8, 1.2, (n, 3)).astype(np.float32)
    xyz = np.stack([rng.uniform(2, 68, n), rng.uniform(-30, 30, n), rng.uniform(-2, -1, n)], 1)
    yaw = rng.uniform(-2 * np.pi, 2 * np.pi, (n, 1))
    return np.concatenate([xyz, dims, yaw], 1).astype(np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load_reference()
    rng = np.random.default_rng(20261017)
    g = {}

    # --- corners / rotation / limit_period (lidar_box3d.py:49-89, cam_box3d.py:116-157,
    #     depth_box3d.py:51-91, utils.py:10-25, 28-117)
    boxes = kitti_like_boxes(rng, 64)
    tb = torch.from_numpy(boxes)
    g['boxes_lidar'] = boxes
    g['corners_lidar'] = ref.LiDARInstance3DBoxes(tb).corners.numpy()
    g['corners_depth'] = ref.DepthInstance3DBoxes(tb).corners.numpy()
    g['corners_cam'] = ref.CameraInstance3DBoxes(tb).corners.numpy()
    g['corners_cam_center_origin'] = ref.CameraInstance3DBoxes(
        tb, origin=(0.5, 0.5, 0.5)).corners.numpy()
    ang = torch.from_numpy(rng.uniform(-7, 7, 64).astype(np.float32))
    g['lp_in'] = ang.numpy()
    g['lp_pi'] = ref.limit_period(ang).numpy()
    g['lp_2pi'] = ref.limit_period(ang, 0.5, np.pi * 2).numpy()
    pts = torch.from_numpy(rng.normal(size=(64, 5, 3)).astype(np.float32))
    g['rot_pts'] = pts.numpy()
    g['rot_axis2'] = ref.rotation_3d_in_axis(pts, ang, axis=2).numpy()
    g['rot_axis1'] = ref.rotation_3d_in_axis(pts, ang, axis=1).numpy()
    g['rot_axis2_cw'] = ref.rotation_3d_in_axis(pts, ang, axis=2, clockwise=True).numpy()

    # --- LIDAR -> CAM conversion (box_3d_mode.py:117-123,162-173 via lidar_box3d.py:177-195)
    rt = torch.from_numpy(RECT @ TRV2C)
    cam = ref.LiDARInstance3DBoxes(tb).convert_to(ref.Box3DMode.CAM, rt)
    g['rect'], g['Trv2c'], g['P2'] = RECT, TRV2C, P2
    g['boxes_cam'] = cam.tensor.numpy()

    # --- points_cam2img (utils.py:175-214)
    p3 = torch.from_numpy((rng.normal(size=(40, 3)) * [5, 2, 10] + [0, 0, 20]).astype(np.float32))
    g['c2i_pts'] = p3.numpy()
    g['c2i_uv'] = ref.points_cam2img(p3, torch.from_numpy(P2)).numpy()
    g['c2i_uv_3x4'] = ref.points_cam2img(p3, torch.from_numpy(P2[:3])).numpy()

    # --- variant B: body of convert_valid_bboxes (kitti_dataset_GGA_match.py:713-748) and the
    #     clamp of bbox2result_kitti (:511-512), composed from the reference's own box classes
    #     (the dataset module itself needs mmdet registries and cannot be imported).
    lb = ref.LiDARInstance3DBoxes(tb.clone())
    lb.limit_yaw(offset=0.5, period=np.pi * 2)                                   # :713
    camb = lb.convert_to(ref.Box3DMode.CAM, rt)                                  # :730
    uv = ref.points_cam2img(camb.corners, torch.from_numpy(P2))                  # :732-733
    b2d = torch.cat([uv.min(dim=1)[0], uv.max(dim=1)[0]], dim=1)                 # :735-737
    ishape = torch.tensor(IMG_HW, dtype=torch.float32)
    vcam = ((b2d[:, 0] < ishape[1]) & (b2d[:, 1] < ishape[0]) & (b2d[:, 2] > 0) & (b2d[:, 3] > 0))
    lim = torch.tensor(PCD_RANGE)
    vp = ((lb.center > lim[:3]) & (lb.center < lim[3:])).all(-1)                 # :745-748
    bb = b2d.numpy().copy()
    bb[:, 2:] = np.minimum(bb[:, 2:], np.array(IMG_HW[::-1], np.float32))        # :511
    bb[:, :2] = np.maximum(bb[:, :2], [0, 0])                                    # :512
    g['varB_box2d_raw'] = b2d.numpy()
    g['varB_box2d_clamped'] = bb.astype(np.float32)
    g['varB_valid'] = (vcam & vp).numpy()
    g['varB_valid_cam'] = vcam.numpy()
    g['img_hw'] = np.array(IMG_HW)
    g['pcd_range'] = np.array(PCD_RANGE, np.float32)

    # the reference's own end-to-end golden for this path:
    # tests/test_data/test_datasets/test_kitti_dataset.py:378-379 (box) -> :393 (bbox)
    one = torch.tensor([[8.7314, -1.8559, -1.5997, 1.2000, 0.4800, 1.8900, -1.5808]])
    l1 = ref.LiDARInstance3DBoxes(one.clone())
    l1.limit_yaw(offset=0.5, period=np.pi * 2)
    c1 = l1.convert_to(ref.Box3DMode.CAM, rt)
    uv1 = ref.points_cam2img(c1.corners, torch.from_numpy(P2))
    g['kitti_one_box2d'] = torch.cat([uv1.min(1)[0], uv1.max(1)[0]], 1).numpy()

    # --- variant A: centerpoint_head_gga.py:252-275,317-338 composed from the reference's
    #     rotation_3d_in_axis exactly as the head composes it (head module not importable).
    l2i = (P2 @ RECT @ TRV2C).astype(np.float32)
    l2i_obj = np.repeat(l2i[None], 64, 0)
    l2i_obj[1::2, :3, 3] += rng.normal(size=(32, 3)).astype(np.float32) * 0.05  # per-object calib
    tb_req = tb.clone().requires_grad_(True)

    def variant_a(bx, l2):
        dims = bx[:, 3:6]
        cn = torch.from_numpy(np.stack(np.unravel_index(np.arange(8), [2] * 3), axis=1)).to(dims.dtype)
        cn = cn[[0, 1, 3, 2, 4, 5, 7, 6]] - dims.new_tensor([0.5, 0.5, 0])
        co = dims.view([-1, 1, 3]) * cn.reshape([1, 8, 3])
        co = ref.rotation_3d_in_axis(co, bx[:, 6], axis=2)
        co = co + bx[:, :3].view(-1, 1, 3)
        co = torch.cat((co, torch.ones(co.shape[0], co.shape[1], 1)), dim=-1)
        pim = torch.einsum('bij,bjk->bik', l2, co.permute(0, 2, 1))
        depth = torch.maximum(pim[:, 2, None, :], torch.tensor([0.1]))
        pix = (pim[:, :2, :] / depth).permute(0, 2, 1)
        return torch.cat((pix[..., 0].min(-1)[0][:, None], pix[..., 1].min(-1)[0][:, None],
                          pix[..., 0].max(-1)[0][:, None], pix[..., 1].max(-1)[0][:, None]), dim=-1)

    pa = variant_a(tb_req, torch.from_numpy(l2i_obj))
    g['varA_lidar2img'] = l2i_obj
    g['varA_box2d'] = pa.detach().numpy()
    gout = torch.from_numpy(rng.normal(size=(64, 4)).astype(np.float32))
    pa.backward(gout)
    g['varA_gout'] = gout.numpy()
    g['varA_grad_boxes'] = tb_req.grad.numpy()

    # --- variant C: pgd_head.py:413-427 (CAM boxes, origin (.5,.5,.5), cam2img 4x4)
    cam_c = cam.tensor.clone()
    cam_c[:, 1] -= cam_c[:, 4] * 0.5          # gravity-centre form, as PGD decodes it
    cam_c.requires_grad_(True)
    cc = ref.CameraInstance3DBoxes(cam_c, box_dim=7, origin=(0.5, 0.5, 0.5)).corners
    uvc = ref.points_cam2img(cc, torch.from_numpy(P2))
    pc = torch.cat([uvc.min(dim=1)[0], uvc.max(dim=1)[0]], dim=1)
    pc.backward(gout)
    g['varC_boxes_cam_center'] = cam_c.detach().numpy()
    g['varC_box2d'] = pc.detach().numpy()
    g['varC_grad_boxes'] = cam_c.grad.numpy()

    # --- variant B gradient (autograd through the reference classes; unclamped box)
    tb2 = tb.clone().requires_grad_(True)
    cb2 = ref.LiDARInstance3DBoxes(tb2).convert_to(ref.Box3DMode.CAM, rt)
    uv2 = ref.points_cam2img(cb2.corners, torch.from_numpy(P2))
    pb = torch.cat([uv2.min(dim=1)[0], uv2.max(dim=1)[0]], dim=1)
    pb.backward(gout)
    g['varB_noyawlimit_box2d'] = pb.detach().numpy()
    g['varB_grad_boxes'] = tb2.grad.numpy()
    np.savez_compressed(os.path.join(OUT, 'ref_geometry.npz'), **g)

    # --- IoU family
    h = {}
    a = rng.uniform(0, 100, (200, 2))
    b1 = np.concatenate([a, a + rng.uniform(1, 60, (200, 2))], 1)
    c = a + rng.normal(size=(200, 2)) * 15
    b2 = np.concatenate([c, c + rng.uniform(1, 60, (200, 2))], 1)
    b2[:10] = b1[:10]                      # identical pairs
    b2[10:20, :2] = b1[10:20, 2:] + 5      # disjoint pairs
    b2[10:20, 2:] = b2[10:20, :2] + 7
    b2[20:24] = b1[20:24]
    b2[20:24, 0] = b1[20:24, 2]            # touching edges (zero-width overlap)
    b2[20:24, 2] = b2[20:24, 0] + 3
    h['b1'], h['b2'] = b1, b2
    # image_box_overlap (kitti_utils/eval.py:85-114), numba, float64 and the mixed
    # float32-dt / float64-gt typing pseudo_label_matching_kitti feeds it
    # (tools/utils_pseudo_labels_gga.py:45 passes dt first; dt bbox is float32 from
    #  convert_valid_bboxes' .numpy(), gt bbox float64 from the info pkl).
    h['ibo_f64'] = ref.image_box_overlap(b1[:50], b2[:30])
    h['ibo_f32_f64'] = ref.image_box_overlap(b1[:50].astype(np.float32), b2[:30])
    h['ibo_f32_f32'] = ref.image_box_overlap(b1[:50].astype(np.float32), b2[:30].astype(np.float32))
    # axis_aligned_bbox_overlaps_3d (iou3d_calculator.py:210-329) on boxes with z in [0, 1]:
    # the z factors are exactly 1, so the values equal the 2D formula of mmdet bbox_overlaps.
    def to3(b):
        z0 = np.zeros((b.shape[0], 1))
        return torch.from_numpy(np.concatenate([b[:, :2], z0, b[:, 2:], z0 + 1], 1).astype(np.float32))
    t1, t2 = to3(b1).requires_grad_(True), to3(b2)
    iou = ref.axis_aligned_bbox_overlaps_3d(t1, t2, mode='iou', is_aligned=True)
    giou = ref.axis_aligned_bbox_overlaps_3d(t1, t2, mode='giou', is_aligned=True)
    h['aa3d_iou'], h['aa3d_giou'] = iou.detach().numpy(), giou.detach().numpy()
    w = torch.from_numpy(rng.uniform(0, 1, 200).astype(np.float32))
    ((1 - giou) * w).sum().backward()
    h['w'] = w.numpy()
    h['aa3d_giou_loss_grad_b1'] = t1.grad.numpy()[:, [0, 1, 3, 4]]
    t1b = to3(b1).requires_grad_(True)
    iou_b = ref.axis_aligned_bbox_overlaps_3d(t1b, t2, mode='iou', is_aligned=True)
    ((1 - iou_b) * w).sum().backward()
    h['aa3d_iou_loss_grad_b1'] = t1b.grad.numpy()[:, [0, 1, 3, 4]]
    # true 3D boxes for the AxisAlignedIoULoss twin
    a3 = rng.uniform(0, 10, (64, 3))
    q1 = np.concatenate([a3, a3 + rng.uniform(0.5, 5, (64, 3))], 1).astype(np.float32)
    c3 = a3 + rng.normal(size=(64, 3))
    q2 = np.concatenate([c3, c3 + rng.uniform(0.5, 5, (64, 3))], 1).astype(np.float32)
    h['q1'], h['q2'] = q1, q2
    h['aa3d_iou_3d'] = ref.axis_aligned_bbox_overlaps_3d(
        torch.from_numpy(q1), torch.from_numpy(q2), mode='iou', is_aligned=True).numpy()
    h['aa3d_giou_3d'] = ref.axis_aligned_bbox_overlaps_3d(
        torch.from_numpy(q1), torch.from_numpy(q2), mode='giou', is_aligned=True).numpy()
    np.savez_compressed(os.path.join(OUT, 'ref_iou.npz'), **h)

    # --- numpy/numba membership twin (box_np_ops.py:353-376) — the a6 contract ("next" row)
    m = {}
    mb = kitti_like_boxes(rng, 16)
    n_in = 600
    sel = rng.integers(0, 16, n_in)
    loc = (rng.uniform(-0.6, 0.6, (n_in, 3)) * mb[sel, 3:6]).astype(np.float32)
    cs, sn = np.cos(mb[sel, 6]), np.sin(mb[sel, 6])
    pin = np.stack([mb[sel, 0] + loc[:, 0] * cs - loc[:, 1] * sn,
                    mb[sel, 1] + loc[:, 0] * sn + loc[:, 1] * cs,
                    mb[sel, 2] + mb[sel, 5] * 0.5 + loc[:, 2]], 1).astype(np.float32)
    pout = np.stack([rng.uniform(0, 70, 400), rng.uniform(-40, 40, 400), rng.uniform(-3, 1, 400)], 1)
    mp = np.concatenate([pin, pout.astype(np.float32)], 0)
    m['pts'], m['boxes'] = mp, mb
    m['points_in_rbbox'] = ref.box_np_ops.points_in_rbbox(mp, mb)
    np.savez_compressed(os.path.join(OUT, 'ref_rbbox.npz'), **m)
    # --- GGA head functions, executed from the reference's own source text
    #     (centerpoint_head_gga.py:167-182 GGA_calculate_rotation, :184-248 get_distance_single/_bev
    #      = Point-to-Box Alignment distances, :250-341 get_prediction_single), with torch-CPU autograd
    #     through them; train_cfg of configs/gga/gga_kitti_config.py
    cfg = dict(grid_size=[1408, 1600, 40], out_size_factor=8, voxel_size=[0.05, 0.05, 0.1],
               point_cloud_range=[0, -40, -3, 70.4, 40, 1])
    H = ref_loader.load_head_functions(cfg, norm_bbox=True)
    hd = {}
    Bq, Kq = 2, 24
    fmx = 1408 // 8
    pred = np.zeros((Bq, Kq, 8), np.float32)
    pred[..., 0:2] = rng.uniform(0, 1, (Bq, Kq, 2))
    pred[..., 2] = rng.uniform(-1.5, -0.3, (Bq, Kq))
    pred[..., 3:6] = np.log(kitti_like_boxes(rng, Bq * Kq)[:, 3:6]).reshape(Bq, Kq, 3)
    ang = rng.uniform(-np.pi, np.pi, (Bq, Kq))
    pred[..., 6], pred[..., 7] = np.sin(ang), np.cos(ang)
    ind = (rng.integers(20, 180, (Bq, Kq)) * fmx + rng.integers(5, 170, (Bq, Kq))).astype(np.int64)
    l2i = (P2 @ RECT @ TRV2C).astype(np.float32)
    l2i_all = np.broadcast_to(l2i, (Bq, Kq, 4, 4)).copy()
    l2i_all[1, ::3, :3, 3] += rng.normal(0, 0.05, (Kq // 3, 3)).astype(np.float32)  # copy-pasted objects keep their own calib
    tp = torch.from_numpy(pred).clone().requires_grad_(True)
    rot, rmat = H.GGA_calculate_rotation(tp[..., 6:])
    ratio, piou, pbev = H.get_prediction_single(tp, torch.from_numpy(ind), torch.from_numpy(l2i_all), rot)
    gi = torch.from_numpy(rng.normal(size=tuple(piou.shape)).astype(np.float32))
    gb = torch.from_numpy(rng.normal(size=tuple(pbev.shape)).astype(np.float32))
    gr = torch.from_numpy(rng.normal(size=tuple(ratio.shape)).astype(np.float32))
    ((piou * gi).sum() + (pbev * gb).sum() + (ratio * gr).sum()).backward()
    hd['gps_pred'], hd['gps_ind'], hd['gps_lidar2img'] = pred, ind, l2i_all
    hd['gps_rot'], hd['gps_ratio'], hd['gps_iou'], hd['gps_bev'] = (rot.detach().numpy(), ratio.detach().numpy(),
                                                                     piou.detach().numpy(), pbev.detach().numpy())
    hd['gps_gi'], hd['gps_gb'], hd['gps_gr'], hd['gps_grad_pred'] = gi.numpy(), gb.numpy(), gr.numpy(), tp.grad.numpy()
    # the same reference source text evaluated in float64 (same inputs, .double()): the yardstick the
    # 1e-5 parity tests hold the CUDA kernels to (tests/parity.py); consumes no random numbers
    try:
        tp64 = torch.from_numpy(pred).double().requires_grad_(True)
        rot64, _ = H.GGA_calculate_rotation(tp64[..., 6:])
        ratio64, piou64, pbev64 = H.get_prediction_single(tp64, torch.from_numpy(ind), torch.from_numpy(l2i_all).double(), rot64)
        assert piou64.dtype == torch.float64
        ((piou64 * gi.double()).sum() + (pbev64 * gb.double()).sum() + (ratio64 * gr.double()).sum()).backward()
        hd['gps_iou_f64'], hd['gps_bev_f64'], hd['gps_ratio_f64'] = (piou64.detach().numpy(), pbev64.detach().numpy(),
                                                                     ratio64.detach().numpy())
        hd['gps_grad_pred_f64'] = tp64.grad.numpy()
    except Exception as e:  # noqa: BLE001
        print('get_prediction_single does not run in float64:', repr(e))
    # Point-to-Box Alignment: ragged in-box point lists (float64 [n_i, 4] = x, y, z, 1 like
    # kitti_converter_gga.py:245-247), some empty, some far outside the predicted box
    bev = pbev.detach().clone().requires_grad_(True)
    lists, counts = [], []
    for b_ in range(Bq):
        row = []
        for k_ in range(Kq):
            n_i = int(rng.choice([0, 1, 3, 17, 64, 300, 1500]))
            cx, cy, w_, h_, r_ = bev[b_, k_].detach().numpy()
            loc = rng.uniform(-0.9, 0.9, (n_i, 2)) * np.array([w_, h_]) * rng.choice([0.5, 1.0, 3.0])
            c_, s_ = np.cos(r_), np.sin(r_)
            xy = np.stack([cx + loc[:, 0] * c_ - loc[:, 1] * s_, cy + loc[:, 0] * s_ + loc[:, 1] * c_], 1)
            pts4 = np.concatenate([xy, rng.uniform(-2, 0, (n_i, 1)), np.ones((n_i, 1))], 1).astype(np.float64)
            row.append(torch.from_numpy(pts4))
            counts.append(n_i)
        lists.append(row)
    dmin, dxs, dys = H.get_distance_bev(lists, bev)
    cm, cx_, cy_ = (torch.from_numpy(rng.uniform(0.5, 1.5, tuple(dmin.shape)).astype(np.float32)) for _ in range(3))
    ((dmin * cm).sum() + (dxs * cx_).sum() + (dys * cy_).sum()).backward()
    hd['pal_bev'] = bev.detach().numpy()
    hd['pal_counts'] = np.asarray(counts, np.int32)
    hd['pal_points_xy'] = np.concatenate([t[:, :2].numpy() for row in lists for t in row], 0)   # float64, object-major
    hd['pal_min'], hd['pal_x'], hd['pal_y'] = dmin.detach().numpy(), dxs.detach().numpy(), dys.detach().numpy()
    hd['pal_cm'], hd['pal_cx'], hd['pal_cy'] = cm.numpy(), cx_.numpy(), cy_.numpy()
    hd['pal_grad_bev'] = bev.grad.numpy()
    np.savez_compressed(os.path.join(OUT, 'ref_head.npz'), **hd)

    # --- convex-polygon membership family (contract a6 of SURVEY.md §8a and §8f rank 2):
    #     box_np_ops.points_in_rbbox (:353-376) with float32 and float64 boxes, the frustum
    #     membership of tools/data_converter/utils_gga.py:88-101, and FCAF3D's face distances
    #     (fcaf3d_head.py:495-520, extracted from its source text like the head functions above)
    cp = {}
    nb = ref.box_np_ops
    cb = kitti_like_boxes(rng, 24)
    cb[:3, 6] = 0.0
    cb[:3, :3] = np.round(cb[:3, :3] * 4) / 4
    cb[:3, 3:6] = [4.0, 2.0, 1.5]
    n_in = 1500
    sel = rng.integers(0, 24, n_in)
    loc = (rng.uniform(-0.6, 0.6, (n_in, 3)) * cb[sel, 3:6]).astype(np.float32)
    cs, sn = np.cos(cb[sel, 6]), np.sin(cb[sel, 6])
    pin = np.stack([cb[sel, 0] + loc[:, 0] * cs - loc[:, 1] * sn, cb[sel, 1] + loc[:, 0] * sn + loc[:, 1] * cs,
                    cb[sel, 2] + cb[sel, 5] * 0.5 + loc[:, 2]], 1)
    face = np.stack([cb[:3, 0] + 2.0, cb[:3, 1], cb[:3, 2] + 0.75], 1)            # exactly on a face: outside (open)
    face2 = np.stack([cb[:3, 0], cb[:3, 1], cb[:3, 2]], 1)                         # on the bottom face: outside too
    pout = np.stack([rng.uniform(0, 70, 600), rng.uniform(-40, 40, 600), rng.uniform(-3, 1, 600)], 1)
    cpts = np.concatenate([pin, face, face2, pout], 0).astype(np.float32)
    cpts4 = np.concatenate([cpts, rng.uniform(0, 1, (len(cpts), 1)).astype(np.float32)], 1)
    cp['pts'], cp['boxes'] = cpts4, cb
    cp['rbbox_f32'] = nb.points_in_rbbox(cpts4, cb)
    cp['rbbox_f64boxes'] = nb.points_in_rbbox(cpts4, cb.astype(np.float64))
    cp['rbbox_f64all'] = nb.points_in_rbbox(cpts4.astype(np.float64), cb.astype(np.float64))
    camb = kitti_like_boxes(rng, 8)
    cp['boxes_cam'] = camb
    cp['rbbox_cam_axis1'] = nb.points_in_rbbox(cpts4, camb, z_axis=1, origin=(0.5, 1.0, 0.5))
    surf = nb.corner_to_surfaces_3d(nb.center_to_corner_box3d(cb[:, :3], cb[:, 3:6], cb[:, 6], origin=(0.5, 0.5, 0), axis=2))
    nv, dd = nb.surface_equ_3d(surf[:, :, :3, :])
    cp['surfaces_f32'], cp['normal_f32'], cp['d_f32'] = surf, nv, dd
    # frustum membership (utils_gga.points_in_frustm_indices)
    ug = ref_loader._load('gga_tools_utils_gga', 'tools/data_converter/utils_gga.py')
    fpts = np.stack([rng.uniform(0, 70, 4000), rng.uniform(-40, 40, 4000), rng.uniform(-3, 1, 4000),
                     rng.uniform(0, 1, 4000)], 1).astype(np.float32)
    rect64, trv64, p264 = RECT.astype(np.float64), TRV2C.astype(np.float64), P2.astype(np.float64)
    bbs = np.array([[100.0, 120.0, 400.0, 300.0], [600.0, 150.0, 700.0, 250.0], [0.0, 0.0, 1242.0, 375.0]])
    cp['fr_pts'], cp['fr_bboxes'] = fpts, bbs
    cp['fr_rect'], cp['fr_Trv2c'], cp['fr_P2'] = rect64, trv64, p264
    cp['fr_indices'] = np.stack([ug.points_in_frustm_indices(fpts, rect64, trv64, p264, bb)[:, 0] for bb in bbs], 1)
    C_, R_, T_ = nb.projection_matrix_to_CRT_kitti(p264)
    fr_surf = []
    for bb in bbs:
        fr = nb.get_frustum(bb.tolist(), C_)
        fr -= T_
        fr = np.linalg.inv(R_) @ fr.T
        fr = nb.camera_to_lidar(fr.T, rect64, trv64)
        fr_surf.append(nb.corner_to_surfaces_3d_jit(fr[np.newaxis, ...])[0])
    cp['fr_surfaces'] = np.stack(fr_surf)
    # FCAF3D face distances
    import ast
    import types as _types
    fpath = os.path.join(ref_loader.REF_ROOT, 'mmdet3d/models/dense_heads/fcaf3d_head.py')
    tree = ast.parse(open(fpath).read())
    fn = [n for c in tree.body if isinstance(c, ast.ClassDef) and c.name == 'FCAF3DHead'
          for n in c.body if isinstance(n, ast.FunctionDef) and n.name == '_get_face_distances'][0]
    fn.decorator_list = []
    glb = {'torch': torch, 'rotation_3d_in_axis': ref.rotation_3d_in_axis}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), fpath, 'exec'), glb)
    fb = np.concatenate([rng.uniform(-4, 4, (12, 2)), rng.uniform(-1.5, 0.5, (12, 1)), rng.uniform(0.3, 2.5, (12, 3)),
                         rng.uniform(-np.pi, np.pi, (12, 1))], 1).astype(np.float32)      # gravity centre boxes
    fp = np.stack([rng.uniform(-5, 5, 900), rng.uniform(-5, 5, 900), rng.uniform(-2, 1, 900)], 1).astype(np.float32)
    tpts = torch.from_numpy(fp).unsqueeze(1).expand(900, 12, 3)
    tbox = torch.from_numpy(fb).expand(900, 12, 7)
    fd = glb['_get_face_distances'](tpts, tbox)
    cp['fd_pts'], cp['fd_boxes'], cp['fd_dist'] = fp, fb, fd.numpy()
    cp['fd_inside'] = (fd.min(dim=-1).values > 0).numpy()
    np.savez_compressed(os.path.join(OUT, 'ref_convex.npz'), **cp)
    gen_targets()
    gen_format()
    gen_npops()
    print('wrote', sorted(os.listdir(OUT)))


TARGET_CASES = (
    # name, class_names, train_cfg overrides, objects per frame, pseudo dtype
    ('kitti64', [['Pedestrian'], ['Cyclist'], ['Car']], {}, (40, 0, 15, 3), np.float64),
    ('kitti32', [['Pedestrian'], ['Cyclist'], ['Car']], {}, (33, 12), np.float32),
    ('multi64', [['a', 'b'], ['c'], ['d', 'e', 'f']], dict(max_objs=6, min_radius=1, gaussian_overlap=0.3), (50, 21),
     np.float64),
    ('multi32', [['a', 'b', 'c'], ['d']], dict(max_objs=9, dense_reg=2, out_size_factor=4, gaussian_overlap=0.5),
     (60, 14, 0), np.float32),
)


def gen_targets():
    """tests/golden/ref_targets.npz: the reference's own get_targets_single (source text executed
    by ref_loader.load_target_functions) on synthetic GGA annotations — KITTI tasks and multi-class
    tasks, fp64 and fp32 pseudo labels, slots beyond max_objs, degenerate and border objects."""
    import types
    from gga_b200 import synth
    from gga_b200.targets import semantic_ratio_samples
    g = {}
    for name, class_names, over, counts, dt in TARGET_CASES:
        cfg = dict(synth.KITTI_TRAIN_CFG, **over)
        fn = ref_loader.load_target_functions(cfg, class_names)
        rng = np.random.default_rng(sum(map(ord, name)))
        n_classes = sum(len(c) for c in class_names)
        for f, n in enumerate(counts):
            fr = synth.make_target_frame(rng, n, n_classes, dt)
            for k, v in fr.items():
                g[f'{name}_f{f}_{k}'] = v
            gt = types.SimpleNamespace(gravity_center=torch.zeros((n, 3)), tensor=torch.zeros((n, 7)))
            ibp = [torch.full((1 + i % 3, 4), float(i), dtype=torch.float64) for i in range(n)]
            torch.manual_seed(1000 + f)
            g[f'{name}_f{f}_srl'] = semantic_ratio_samples(1, len(class_names))[0].numpy()
            torch.manual_seed(1000 + f)
            out = fn(gt, torch.from_numpy(fr['labels']), torch.from_numpy(fr['boxes_img']),
                     torch.from_numpy(fr['lidar2img']), torch.from_numpy(fr['pseudo']), torch.from_numpy(fr['bdry']),
                     ibp, dict(lidar2img=fr['base_lidar2img']))
            heat, anno, ind, mask, l2i, tibp, bm = out
            for t in range(len(class_names)):
                g[f'{name}_f{f}_t{t}_heatmap'] = heat[t].numpy()
                g[f'{name}_f{f}_t{t}_anno_box'] = anno[t].numpy()
                g[f'{name}_f{f}_t{t}_ind'] = ind[t].numpy()
                g[f'{name}_f{f}_t{t}_mask'] = mask[t].numpy()
                g[f'{name}_f{f}_t{t}_lidar2img'] = l2i[t].numpy()
                g[f'{name}_f{f}_t{t}_bmask'] = bm[t].numpy()
                g[f'{name}_f{f}_t{t}_ibp_order'] = np.asarray([int(p[0, 0]) for p in tibp[t]], np.int32)
    np.savez_compressed(os.path.join(OUT, 'ref_targets.npz'), **g)


FORMAT_COUNTS = (24, 0, 7, 2, 40)
FORMAT_CLASSES = ['Pedestrian', 'Cyclist', 'Car']
ANNO_KEYS = ('name', 'truncated', 'occluded', 'alpha', 'bbox', 'dimensions', 'location', 'rotation_y', 'score',
             'sample_idx')


def gen_format():
    """tests/golden/ref_format.npz: the reference's own bbox2result_kitti / convert_valid_bboxes
    (kitti_dataset_GGA_match.py:458-571, 685-765) and pseudo_label_matching_kitti
    (tools/utils_pseudo_labels_gga.py:17-88), source text executed through
    ref_loader.load_format_functions on synthetic detections."""
    import copy
    import tempfile
    from gga_b200 import synth
    infos, dets = synth.make_detection_frames(2024, FORMAT_COUNTS)
    R = ref_loader.load_format_functions(infos, list(synth.KITTI_MATCH_RANGE))
    net = [dict(boxes_3d=R.LiDARInstance3DBoxes(torch.from_numpy(d['boxes_3d']).clone()),
                scores_3d=torch.from_numpy(d['scores_3d']), labels_3d=torch.from_numpy(d['labels_3d'])) for d in dets]
    g = {}
    with tempfile.TemporaryDirectory() as tmp:
        annos = R.bbox2result_kitti(net, FORMAT_CLASSES, submission_prefix=tmp)
        for f, a in enumerate(annos):
            for k in ANNO_KEYS:
                g[f'f{f}_{k}'] = np.asarray(a[k])
            g[f'f{f}_txt'] = np.array(open(os.path.join(tmp, f"{infos[f]['image']['image_idx']:06d}.txt")).read())
    gt_infos = copy.deepcopy(infos)
    # frames without detections keep their annotations' keys but no rows (:52-57); frame 3 may lose
    # every detection to the validity test
    cleaned = R.pseudo_label_matching_kitti(gt_infos, copy.deepcopy(annos))
    (path, new_infos), = R.dumped
    for f in range(len(infos)):
        for k, v in new_infos[f]['annos'].items():
            g[f'f{f}_new_{k}'] = np.asarray(v)
        for k, v in cleaned[f].items():
            g[f'f{f}_clean_{k}'] = np.asarray(v)
    g['dump_path'] = np.array(path)
    np.savez_compressed(os.path.join(OUT, 'ref_format.npz'), **g)


def gen_npops():
    """tests/golden/ref_npops.npz: the reference's own box_np_ops.box3d_to_bbox (:311-328) and
    iou_jit (:482-523) on camera boxes / 2D boxes (rows a16 and a25 of SURVEY.md §8a)."""
    ref = ref_loader.load_reference()
    rng = np.random.default_rng(1625)
    n = 96
    cam = np.concatenate([rng.uniform(-15, 15, (n, 1)), rng.uniform(0.5, 2.5, (n, 1)), rng.uniform(4, 60, (n, 1)),
                          rng.uniform(0.5, 4.5, (n, 3)), rng.uniform(-np.pi, np.pi, (n, 1))], 1)
    g = {}
    for name, dt in (('f32', np.float32), ('f64', np.float64)):
        b = cam.astype(dt)
        g[f'cam_{name}'] = b
        g[f'bbox_{name}'] = ref.box_np_ops.box3d_to_bbox(b, P2.astype(dt))
    g['P2'] = P2
    x1y1 = np.stack([rng.uniform(0, 1100, 70), rng.uniform(0, 300, 70)], 1)
    bx = np.concatenate([x1y1, x1y1 + rng.uniform(1, 250, (70, 2))], 1)
    bx[5] = bx[6]                       # identical boxes
    bx[7, 2:] = bx[7, :2]               # zero-area box
    q = bx[::2] + rng.normal(0, 8, (35, 4))
    for name, dt in (('f32', np.float32), ('f64', np.float64)):
        g[f'iou_boxes_{name}'], g[f'iou_query_{name}'] = bx.astype(dt), q.astype(dt)
        for mode in ('iou', 'iof'):
            for eps in (0.0, 1.0):
                g[f'iou_{name}_{mode}_{int(eps)}'] = ref.box_np_ops.iou_jit(bx.astype(dt), q.astype(dt), mode, eps)
    np.savez_compressed(os.path.join(OUT, 'ref_npops.npz'), **g)


if __name__ == '__main__':
    if sys.argv[1:] == ['npops']:
        gen_npops()
    elif sys.argv[1:] == ['targets']:
        gen_targets()
    elif sys.argv[1:] == ['format']:
        gen_format()
    else:
        main()
