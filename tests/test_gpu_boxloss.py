"""GPU parity tests, parts 2 + 3: projection variants, 2D losses, fused projection+loss with
backward, head-level entry points and matching — CUDA (through the C ABI) against the CPU
oracle and against the fixtures generated from the reference itself.  Tolerance: 1e-5
relative (BASELINE.json north_star), with an absolute floor tied to the magnitude of the
quantity (pixels ~1e3 -> 2e-3 px).  Gradients (and the loss values they belong to) are held to
1e-5 of their largest entry against a FLOAT64 evaluation of the same formulation (tests/parity.py)."""
import os

import numpy as np
import pytest
import torch

import gga_b200 as G
from gga_b200 import synth
from oracle import geometry as og
from oracle import losses as ol
from parity import close64, d64

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def cu(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype).cuda()


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(got, ref, rtol=RTOL, atol_scale=1e-5):
    got = got.detach().cpu().double().numpy() if torch.is_tensor(got) else np.asarray(got, np.float64)
    ref = ref.detach().cpu().double().numpy() if torch.is_tensor(ref) else np.asarray(ref, np.float64)
    scale = np.abs(ref).max() if ref.size else 1.0
    ok = np.allclose(got, ref, rtol=rtol, atol=atol_scale * max(scale, 1e-30))
    if not ok:
        err = np.abs(got - ref)
        i = np.unravel_index(err.argmax(), err.shape)
        print('max abs err', err.max(), 'at', i, 'got', got[i], 'ref', ref[i], 'scale', scale)
    return ok


@pytest.fixture(scope='module')
def GEO(golden_dir):
    return np.load(os.path.join(golden_dir, 'ref_geometry.npz'))


@pytest.fixture(scope='module')
def IOU(golden_dir):
    return np.load(os.path.join(golden_dir, 'ref_iou.npz'))


def test_variant_a_forward_backward_vs_reference_fixture(GEO):
    b = cu(GEO['boxes_lidar']).requires_grad_(True)
    out, valid = G.box3d_project(b, cu(GEO['varA_lidar2img']), mode='lidar_direct', depth_clamp=0.1)
    assert close(out, GEO['varA_box2d'])
    out.backward(cu(GEO['varA_gout']))
    b64, l64, g64 = d64(T(GEO['boxes_lidar']).requires_grad_(True), GEO['varA_lidar2img'], GEO['varA_gout'])
    o64 = og.project_lidar_direct(b64, l64)
    o64.backward(g64)
    assert close64(out, o64, GEO['varA_box2d'], what='variant A box2d')
    assert close64(b.grad, b64.grad, GEO['varA_grad_boxes'], what='variant A grad')
    assert valid.all()


def test_variant_b_vs_reference_fixture_and_known_answer(GEO):
    b = cu(GEO['boxes_lidar'])
    rt = cu(GEO['rect'] @ GEO['Trv2c'])
    raw, valid = G.box3d_project(b, cu(GEO['P2']), mode='kitti_cam', rt=rt, img_hw=GEO['img_hw'],
                                 pcd_range=GEO['pcd_range'], clamp=False)
    assert close(raw, GEO['varB_box2d_raw'])
    assert np.array_equal(valid.cpu().numpy(), GEO['varB_valid'])
    cl, _ = G.box3d_project(b, cu(GEO['P2']), mode='kitti_cam', rt=rt, img_hw=GEO['img_hw'],
                            pcd_range=GEO['pcd_range'], clamp=True)
    assert close(cl, GEO['varB_box2d_clamped'])
    # the reference's own end-to-end golden: test_kitti_dataset.py:378-379 -> :393
    one = cu([[8.7314, -1.8559, -1.5997, 1.2000, 0.4800, 1.8900, -1.5808]])
    o, v = G.box3d_project(one, cu(GEO['P2'][:3]), mode='kitti_cam', rt=rt, img_hw=(375, 1242), clamp=True)
    assert np.allclose(o.cpu().numpy(), [[710.443, 144.00221, 820.29114, 307.58667]], rtol=1e-5, atol=1e-3)
    # gradient of the unclamped box w.r.t. the LiDAR box (yaw already inside (-pi, pi])
    bb = cu(GEO['boxes_lidar']).requires_grad_(True)
    r2, _ = G.box3d_project(bb, cu(GEO['P2']), mode='kitti_cam', rt=rt)
    ref_b = T(GEO['boxes_lidar']).clone().requires_grad_(True)
    rr, _, _ = og.project_kitti_cam(ref_b, T(GEO['rect']), T(GEO['Trv2c']), T(GEO['P2']), GEO['img_hw'],
                                    clamp=False)
    r2.backward(cu(GEO['varA_gout']))
    rr.backward(T(GEO['varA_gout']))
    b64 = d64(ref_b)
    r64, _, _ = og.project_kitti_cam(b64, *d64(GEO['rect'], GEO['Trv2c'], GEO['P2']), GEO['img_hw'], clamp=False)
    r64.backward(d64(GEO['varA_gout']))
    assert close64(bb.grad, b64.grad, ref_b.grad, what='variant B grad')


def test_variant_c_and_cam_bottom_vs_reference_fixture(GEO):
    c = cu(GEO['varC_boxes_cam_center']).requires_grad_(True)
    out, _ = G.box3d_project(c, cu(GEO['P2']), mode='cam_center')
    assert close(out, GEO['varC_box2d'])
    out.backward(cu(GEO['varA_gout']))
    c64 = d64(T(GEO['varC_boxes_cam_center']).requires_grad_(True))
    o64 = og.project_cam(c64, d64(GEO['P2']))
    o64.backward(d64(GEO['varA_gout']))
    assert close64(c.grad, c64.grad, GEO['varC_grad_boxes'], what='variant C grad')
    cb = cu(GEO['boxes_cam'])
    ref = og.minmax_box(og.points_cam2img(og.corners_cam(T(GEO['boxes_cam'])), T(GEO['P2'])))
    assert close(G.box3d_project(cb, cu(GEO['P2']), mode='cam_bottom')[0], ref)


def test_depth_clamp_and_behind_camera_boxes():
    rng = np.random.default_rng(4)
    boxes = synth.make_boxes(rng, 300)
    boxes[:100, 0] = rng.uniform(-3, 3, 100)          # straddling / behind the image plane
    l2i = np.repeat(synth.kitti_lidar2img()[None], 300, 0)
    b = cu(boxes).requires_grad_(True)
    out, _ = G.box3d_project(b, cu(l2i), mode='lidar_direct', depth_clamp=0.1)
    rb = T(boxes).clone().requires_grad_(True)
    ref = og.project_lidar_direct(rb, T(l2i))
    assert close(out, ref)
    g = rng.normal(size=(300, 4)).astype(np.float32)
    out.backward(cu(g))
    ref.backward(T(g))
    b64 = d64(rb)
    og.project_lidar_direct(b64, d64(l2i)).backward(d64(g))
    assert close64(b.grad, b64.grad, rb.grad, what='depth clamp grad')


@pytest.mark.parametrize('kind,mod', [('giou', 'giou'), ('iou_linear', 'linear'), ('iou_square', 'square'),
                                      ('iou_log', 'log'), ('l1', 'l1')])
@pytest.mark.parametrize('wmode', ['none', 'vec', 'mat'])
def test_box2d_losses_vs_oracle(IOU, kind, mod, wmode):
    b1 = IOU['b1'].astype(np.float32)
    b2 = IOU['b2'].astype(np.float32)
    n = b1.shape[0]
    rng = np.random.default_rng(8)
    w = None if wmode == 'none' else (rng.uniform(0, 1, n) if wmode == 'vec' else rng.uniform(0, 1, (n, 4)))
    w = None if w is None else w.astype(np.float32)
    if kind == 'l1' and wmode == 'vec':
        w = np.repeat(w[:, None], 4, 1)
    for avg, red in [(None, 'mean'), (37.5, 'mean'), (None, 'sum')]:
        p = cu(b1).requires_grad_(True)
        t = cu(b2).requires_grad_(True)
        rp = T(b1).clone().requires_grad_(True)
        rt = T(b2).clone().requires_grad_(True)
        wt = None if w is None else T(w)
        if kind == 'giou':
            ref = ol.giou_loss_module(rp, rt, wt, avg, red, 2.0)
        elif kind == 'l1':
            ref = ol.l1_loss_module(rp, rt, wt, avg, red, 2.0)
        else:
            ref = ol.iou_loss_module(rp, rt, wt, avg, red, 2.0, mode=mod)
        got = G.box2d_loss(p, t, None if w is None else cu(w), avg, kind, red, 2.0)
        assert close(got, ref)
        got.backward()
        ref.backward()
        p64, t64, w64 = d64(rp, rt, wt)
        if kind == 'giou':
            r64 = ol.giou_loss_module(p64, t64, w64, avg, red, 2.0)
        elif kind == 'l1':
            r64 = ol.l1_loss_module(p64, t64, w64, avg, red, 2.0)
        else:
            r64 = ol.iou_loss_module(p64, t64, w64, avg, red, 2.0, mode=mod)
        r64.backward()
        assert close64(got, r64, ref, what=f'{kind} loss')
        assert close64(p.grad, p64.grad, rp.grad, what=f'{kind} grad pred')
        assert close64(t.grad, t64.grad, rt.grad, what=f'{kind} grad target')


def test_giou_matches_reference_axis_aligned_formula(IOU):
    b1, b2 = cu(IOU['b1'].astype(np.float32)), cu(IOU['b2'].astype(np.float32))
    l = G.box2d_loss(b1, b2, kind='giou', reduction='none')
    assert close(1 - l, IOU['aa3d_giou'], atol_scale=2e-6)
    li = G.box2d_loss(b1, b2, kind='iou_linear', reduction='none')
    assert close(1 - li, np.maximum(IOU['aa3d_iou'], 1e-6), atol_scale=2e-6)
    p = b1.clone().requires_grad_(True)
    G.box2d_loss(p, b2, cu(IOU['w']), kind='giou', reduction='sum').backward()
    p64 = d64(T(IOU['b1'].astype(np.float32)).requires_grad_(True))
    ol.giou_loss_module(p64, d64(IOU['b2'].astype(np.float32)), d64(IOU['w']), None, 'sum').backward()
    assert close64(p.grad, p64.grad, IOU['aa3d_giou_loss_grad_b1'], what='giou grad vs reference twin')


def test_loss_modules_signature_and_early_out():
    pred = cu([[0., 0., 10., 10.], [0., 0., 10., 10.]]).requires_grad_(True)
    tgt = cu([[0., 0., 10., 10.], [5., 5., 15., 15.]])
    m = G.ProjectedGIoULoss(loss_weight=2.0)
    l = m(pred, tgt, reduction_override='none')
    assert torch.allclose(l.cpu(), 2.0 * torch.tensor([0., 1 - (25 / 175 - (225 - 175) / 225)]))
    w = cu([1.0, 0.5])
    assert torch.allclose(m(pred, tgt, w, avg_factor=4.0).cpu(), (l.cpu() * w.cpu()).sum() / 4.0)
    z = m(pred, tgt, torch.zeros(2, 4).cuda())
    assert float(z.detach()) == 0.0 and z.requires_grad
    l1 = G.ProjectedL1Loss(loss_weight=0.25)(pred, tgt, torch.ones(2, 4).cuda(), avg_factor=2.0)
    assert torch.allclose(l1.cpu(), torch.tensor(0.25 * 20.0 / 2.0))
    il = G.ProjectedIoULoss(mode='log')(pred, tgt)
    assert torch.allclose(il.cpu(), ol.iou_loss_module(pred.detach().cpu(), tgt.cpu()))


@pytest.mark.parametrize('kind', ['giou', 'l1', 'iou_log'])
def test_fused_projection_loss_equals_oracle_chain(kind):
    bt = synth.make_batch(2, 100, 2, N=0)
    boxes = bt['boxes'].reshape(-1, 7)
    l2i = bt['lidar2img'].reshape(-1, 4, 4)
    tgt = bt['target'].reshape(-1, 4)
    n = boxes.shape[0]
    rng = np.random.default_rng(2)
    w = rng.uniform(0, 1, (n, 4)).astype(np.float32) if kind == 'l1' else rng.uniform(0, 1, n).astype(np.float32)
    b = cu(boxes).requires_grad_(True)
    t = cu(tgt).requires_grad_(True)
    got, box2d, valid = G.projected_box_loss(b, cu(l2i), t, cu(w), avg_factor=float(n), kind=kind,
                                             loss_weight=1.5, return_box2d=True)
    rb = T(boxes).clone().requires_grad_(True)
    rtg = T(tgt).clone().requires_grad_(True)
    proj = og.project_lidar_direct(rb, T(l2i))
    fn = {'giou': ol.giou_loss_module, 'l1': ol.l1_loss_module,
          'iou_log': lambda *a, **k: ol.iou_loss_module(*a, mode='log', **k)}[kind]
    ref = fn(proj, rtg, T(w), avg_factor=float(n), loss_weight=1.5)
    assert close(box2d, proj)
    assert close(got, ref)
    got.backward()
    ref.backward()
    b64, t64 = d64(rb, rtg)
    r64 = fn(og.project_lidar_direct(b64, d64(l2i)), t64, d64(w), avg_factor=float(n), loss_weight=1.5)
    r64.backward()
    assert close64(got, r64, ref, what=f'fused {kind} loss')
    assert close64(b.grad, b64.grad, rb.grad, what=f'fused {kind} grad boxes')
    assert close64(t.grad, t64.grad, rtg.grad, what=f'fused {kind} grad target')
    # the fused launch equals the two-step public API
    b2 = cu(boxes).requires_grad_(True)
    two = G.box2d_loss(G.box3d_project(b2, cu(l2i))[0], cu(tgt), cu(w), float(n), kind, 'mean', 1.5)
    two.backward()
    assert close(two, got) and close(b2.grad, b.grad, rtol=1e-5, atol_scale=1e-6)


def test_get_prediction_single_and_bpl_vs_oracle():
    rng = np.random.default_rng(12)
    cfg = dict(grid_size=[1408, 1600, 40], out_size_factor=8, voxel_size=[0.05, 0.05, 0.1],
               point_cloud_range=[0, -40, -3, 70.4, 40, 1])
    B, K = 2, 500
    ind = rng.integers(0, 176 * 200, (B, K))
    pred = np.concatenate([rng.uniform(0, 1, (B, K, 2)), rng.uniform(-1.5, -0.3, (B, K, 1)),
                           np.log(rng.uniform(0.5, 4.5, (B, K, 3))), rng.normal(size=(B, K, 2))], -1).astype(np.float32)
    l2i = np.broadcast_to(synth.kitti_lidar2img(), (B, K, 4, 4)).copy()
    p = cu(pred).requires_grad_(True)
    rot, _ = G.gga_calculate_rotation(p[..., 6:])
    ratio, piou, bev = G.get_prediction_single(p, cu(ind, torch.int64), cu(l2i), rot, cfg)
    rp = T(pred).clone().requires_grad_(True)
    rrot = torch.atan2(rp[..., 6], rp[..., 7])
    r_ratio, r_iou, r_bev = og.get_prediction_single(rp, T(ind), T(l2i), rrot, cfg)
    assert close(piou, r_iou, atol_scale=2e-6) and close(ratio, r_ratio) and close(bev, r_bev)
    tb = np.concatenate([rng.uniform(0, 1242, (B, K, 2)), rng.uniform(0, 375, (B, K, 2)),
                         rng.uniform(1, 3, (B, K, 1))], -1).astype(np.float32)
    tb = tb[..., [0, 2, 1, 3, 4]]
    mask = (rng.uniform(size=(B, K)) < 0.1)
    bmask = (rng.uniform(size=(B, K, 4)) < 0.7)
    got = G.boundary_projection_loss(piou, cu(tb), cu(mask, torch.uint8), cu(bmask, torch.uint8))
    ref = ol.boundary_projection_loss(r_iou, T(tb), T(mask.astype(np.uint8)), T(bmask.astype(np.uint8)))
    assert close(got, ref)
    got.backward()
    ref.backward()
    p64 = d64(rp)
    _, i64, _ = og.get_prediction_single(p64, T(ind), d64(l2i), torch.atan2(p64[..., 6], p64[..., 7]), cfg)
    r64 = ol.boundary_projection_loss(i64, d64(tb), T(mask.astype(np.uint8)), T(bmask.astype(np.uint8)))
    r64.backward()
    assert close64(got, r64, ref, what='BPL loss')
    assert close64(p.grad, p64.grad, rp.grad, what='BPL grad')


def test_matching_vs_oracle_and_reference_fixture(IOU):
    # float32 detections vs float64 annotations: the typing pseudo_label_matching_kitti produces
    dt = IOU['b1'][:50].astype(np.float32)
    gt = IOU['b2'][:30]
    do, go = torch.tensor([0, 50], dtype=torch.int32), torch.tensor([0, 30], dtype=torch.int32)
    m, best, ov, oo = G.match_dt_to_gt(cu(dt), do, cu(gt, torch.float64), go, return_overlaps=True)
    assert np.array_equal(ov.cpu().numpy().reshape(50, 30), IOU['ibo_f32_f64'])       # bit-exact
    assert np.array_equal(m.cpu().numpy(), IOU['ibo_f32_f64'].argmax(-1))
    # dense float64 twin
    d = G.image_box_overlap(cu(IOU['b1'][:50], torch.float64), cu(IOU['b2'][:30], torch.float64))
    assert np.array_equal(d.cpu().numpy(), IOU['ibo_f64'])
    # ragged frames incl. empty ones
    rng = np.random.default_rng(6)
    nd = [0, 5, 512, 3, 0, 77]
    ng = [3, 0, 8, 1, 0, 12]
    a = rng.uniform(0, 1000, (sum(nd), 2))
    dts = np.concatenate([a, a + rng.uniform(1, 200, a.shape)], 1).astype(np.float32)
    c = rng.uniform(0, 1000, (sum(ng), 2))
    gts = np.concatenate([c, c + rng.uniform(1, 200, c.shape)], 1)
    do = np.concatenate([[0], np.cumsum(nd)])
    go = np.concatenate([[0], np.cumsum(ng)])
    m, best = G.match_dt_to_gt(cu(dts), torch.as_tensor(do), cu(gts, torch.float64), torch.as_tensor(go))
    m, best = m.cpu().numpy(), best.cpu().numpy()
    for f in range(len(nd)):
        sl = slice(do[f], do[f + 1])
        if ng[f] == 0:
            assert (m[sl] == -1).all()
            continue
        o = ol.image_box_overlap(dts[sl], gts[go[f]:go[f + 1]])
        assert o.dtype == np.float32
        # float32 dt area like numba: recompute with the mixed typing via the oracle in float64
        ref_m = np.argmax(o, -1) if nd[f] else np.zeros((0,), np.int64)
        if nd[f] == 0:
            continue
        agree = (m[sl] == ref_m)
        # the numpy oracle computes the dt area in float64; ties within 1 ulp may flip — compare IoU values
        assert agree.mean() > 0.99
        assert np.allclose(best[sl], o.max(-1), rtol=1e-6, atol=1e-7)


def test_convert_valid_bboxes_batch_matches_oracle():
    rng = np.random.default_rng(3)
    F = 6
    counts = [0, 40, 512, 1, 300, 64]
    boxes = np.concatenate([synth.make_boxes(rng, c) for c in counts if c]).astype(np.float32)
    boxes[:, 6] = rng.uniform(-7, 7, len(boxes))
    fob = np.concatenate([np.full(c, i) for i, c in enumerate(counts)]).astype(np.int32)
    rect = np.repeat(synth.KITTI_RECT[None], F, 0)
    trv = np.repeat(synth.KITTI_TRV2C[None], F, 0)
    trv[1:, :3, 3] += rng.normal(size=(F - 1, 3)).astype(np.float32) * 0.05
    p2 = np.repeat(synth.KITTI_P2[None], F, 0)
    hw = np.float32([[375, 1242]] * F)
    hw[3] = [370, 1224]
    out = G.convert_valid_bboxes_batch(cu(boxes), cu(fob, torch.int32), cu(rect), cu(trv), cu(p2), cu(hw),
                                       synth.KITTI_MATCH_RANGE)
    for f in range(F):
        sel = fob == f
        if not sel.any():
            continue
        b2d, valid, _ = og.project_kitti_cam(T(boxes[sel]), T(rect[f]), T(trv[f]), T(p2[f]), hw[f],
                                             synth.KITTI_MATCH_RANGE)
        assert close(out['bbox'][torch.as_tensor(sel).cuda()], b2d)
        assert np.array_equal(out['valid'].cpu().numpy()[sel], valid.numpy())


def test_axis_aligned_iou_loss_3d_vs_reference_golden_and_oracle(golden_dir):
    """AxisAlignedIoULoss (FCAF3D): known answer of the reference's test_losses.py:178-189,
    golden IoU / GIoU of axis_aligned_bbox_overlaps_3d, and oracle gradients."""
    import os
    d = np.load(os.path.join(golden_dir, 'ref_iou.npz'))
    q1, q2 = torch.from_numpy(d['q1']).cuda(), torch.from_numpy(d['q2']).cuda()
    for mode, key in (('iou', 'aa3d_iou_3d'), ('giou', 'aa3d_giou_3d')):
        l = G.axis_aligned_iou_loss(q1, q2, reduction='none', mode=mode)
        assert np.allclose(1.0 - l.cpu().numpy(), d[key], rtol=1e-5, atol=1e-6)
    # the reference's own unit test vector
    loss = G.AxisAlignedIoULoss(reduction='mean')
    pred = torch.tensor([[0., 0, 0, 1, 1, 1], [0, 0, 0, 1, 1, 1], [0, 0, 0, 1, 1, 1]]).cuda()
    tgt = torch.tensor([[0., 0, 0, 1, 1, 1], [-4, -4, -4, -1, -1, -1], [0, 0, 0, 0.5, 0.5, 0.5]]).cuda()   # disjoint / contained
    out = loss(pred, tgt, reduction_override='none')
    assert np.allclose(out.cpu().numpy(), [0.0, 1.0, 1.0 - 0.125], atol=1e-6)
    # gradients and weighted reduction against the oracle (torch autograd on the restated formula)
    rng = np.random.default_rng(9)
    lo = rng.uniform(-2, 2, (500, 3)); a = np.concatenate([lo, lo + rng.uniform(0.1, 3, (500, 3))], 1).astype(np.float32)
    lo2 = lo + rng.normal(0, 0.7, (500, 3)); b = np.concatenate([lo2, lo2 + rng.uniform(0.1, 3, (500, 3))], 1).astype(np.float32)
    a[:5] = b[:5]                                     # exact ties
    w = rng.uniform(0, 2, 500).astype(np.float32)
    pa = torch.from_numpy(a).requires_grad_(True); pb = torch.from_numpy(b).requires_grad_(True)
    ref = ol.axis_aligned_iou_loss(pa, pb, torch.from_numpy(w), avg_factor=321.0, loss_weight=0.7)
    ref.backward()
    ga = torch.from_numpy(a).cuda().requires_grad_(True); gb = torch.from_numpy(b).cuda().requires_grad_(True)
    got = G.AxisAlignedIoULoss(loss_weight=0.7)(ga, gb, torch.from_numpy(w).cuda(), avg_factor=321.0)
    got.backward()
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref)) + 1e-7
    a64, b64 = d64(pa, pb)
    ol.axis_aligned_iou_loss(a64, b64, d64(w), avg_factor=321.0, loss_weight=0.7).backward()
    for g, r64, r in ((ga.grad, a64.grad, pa.grad), (gb.grad, b64.grad, pb.grad)):
        assert close64(g, r64, r, what='AxisAlignedIoULoss grad')
    # early-out: no positive weight
    z = G.AxisAlignedIoULoss()(ga, gb, torch.zeros(500).cuda())
    assert float(z) == 0.0 and z.requires_grad
