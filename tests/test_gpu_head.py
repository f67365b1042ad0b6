"""GPU parity tests of the GGA head functions against golden vectors produced by the
reference's own source text (tests/golden/ref_head.npz) and against the CPU oracle:
get_prediction_single (projection, variant A) and the Point-to-Box Alignment distances with
their gradients.  Tolerance: 1e-5 relative (north_star), absolute floor scaled to the data."""
import os

import numpy as np
import pytest
import torch

import gga_b200 as G
from oracle import geometry as og
from oracle import losses as ol
from parity import close64, d64

pytestmark = pytest.mark.gpu
CFG = dict(grid_size=[1408, 1600, 40], out_size_factor=8, voxel_size=[0.05, 0.05, 0.1],
           point_cloud_range=[0, -40, -3, 70.4, 40, 1])


@pytest.fixture(scope='module')
def H(golden_dir):
    return np.load(os.path.join(golden_dir, 'ref_head.npz'))


def test_get_prediction_single_vs_reference_golden(H):
    pred = torch.from_numpy(H['gps_pred']).cuda().requires_grad_(True)
    rot, _ = G.gga_calculate_rotation(pred[..., 6:])
    assert np.allclose(rot.detach().cpu().numpy(), H['gps_rot'], rtol=1e-6, atol=1e-6)
    ratio, iou, bev = G.get_prediction_single(pred, torch.from_numpy(H['gps_ind']).cuda(),
                                              torch.from_numpy(H['gps_lidar2img']).cuda(), rot, CFG)
    assert np.allclose(ratio.detach().cpu().numpy(), H['gps_ratio'], rtol=1e-6)
    assert np.allclose(bev.detach().cpu().numpy(), H['gps_bev'], rtol=1e-6, atol=1e-6)
    assert np.allclose(iou.detach().cpu().numpy(), H['gps_iou'], rtol=1e-5, atol=2e-3)
    ((iou * torch.from_numpy(H['gps_gi']).cuda()).sum() + (bev * torch.from_numpy(H['gps_gb']).cuda()).sum() +
     (ratio * torch.from_numpy(H['gps_gr']).cuda()).sum()).backward()
    # float64 yardstick: the reference's OWN source text (centerpoint_head_gga.py:250-341) evaluated on the
    # same inputs in double (oracle/gen_golden.py -> gps_*_f64)
    assert close64(iou, H['gps_iou_f64'], H['gps_iou'], what='get_prediction_single box2d')
    assert close64(bev, H['gps_bev_f64'], H['gps_bev'], what='get_prediction_single bev')
    assert close64(pred.grad, H['gps_grad_pred_f64'], H['gps_grad_pred'], what='get_prediction_single grad')


def _lists(H, device):
    counts = H['pal_counts']
    off = np.concatenate([[0], np.cumsum(counts)])
    xy = H['pal_points_xy']
    B, K = H['pal_bev'].shape[:2]
    flat = [torch.from_numpy(np.concatenate([xy[off[i]:off[i + 1]], np.zeros((counts[i], 2))], 1)) for i in range(B * K)]
    return [flat[b * K:(b + 1) * K] for b in range(B)]


def test_point_box_alignment_vs_reference_golden(H):
    lists = _lists(H, 'cuda')
    bev = torch.from_numpy(H['pal_bev']).cuda().requires_grad_(True)
    dmin, dx, dy = G.get_distance_bev(lists, bev)                       # the reference's signature
    for got, key in ((dmin, 'pal_min'), (dx, 'pal_x'), (dy, 'pal_y')):
        ref = H[key]
        assert got.shape == ref.shape
        assert np.allclose(got.detach().cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(ref).max())), key
    ((dmin * torch.from_numpy(H['pal_cm']).cuda()).sum() + (dx * torch.from_numpy(H['pal_cx']).cuda()).sum() +
     (dy * torch.from_numpy(H['pal_cy']).cuda()).sum()).backward()
    counts = H['pal_counts']
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    b64 = d64(torch.from_numpy(H['pal_bev']).reshape(-1, 5).requires_grad_(True))
    m64, x64, y64 = ol.point_box_distances(d64(H['pal_points_xy']), off, b64)
    shp = H['pal_min'].shape
    ((m64.reshape(shp) * d64(H['pal_cm'])).sum() + (x64.reshape(shp) * d64(H['pal_cx'])).sum() +
     (y64.reshape(shp) * d64(H['pal_cy'])).sum()).backward()
    assert close64(dmin, m64.reshape(shp), H['pal_min'], what='PAL min distance')
    assert close64(bev.grad, b64.grad.reshape(H['pal_grad_bev'].shape), H['pal_grad_bev'], what='PAL grad')


def test_get_distance_bev_accepts_ragged_lists_like_the_reference(H):
    """The real loss() call: a frame lists n_obj <= K clusters (centerpoint_head_gga.py:463-479,693);
    the rows of the missing objects stay zero (:190-199) and receive no gradient."""
    lists = _lists(H, 'cuda')
    B, K = H['pal_bev'].shape[:2]
    keep = [K - 5, K // 2]
    ragged = [lists[b][:keep[b % 2]] for b in range(B)]
    bev = torch.from_numpy(H['pal_bev']).cuda().requires_grad_(True)
    dmin, dx, dy = G.get_distance_bev(ragged, bev)
    assert dmin.shape == (B, K, 1)
    for b in range(B):
        n = keep[b % 2]
        for got, key in ((dmin, 'pal_min'), (dx, 'pal_x'), (dy, 'pal_y')):
            ref = H[key][b, :n]
            assert np.allclose(got[b, :n].detach().cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(ref).max()))
            assert (got[b, n:] == 0).all()
    (dmin.sum() + dx.sum() + dy.sum()).backward()
    for b in range(B):
        assert (bev.grad[b, keep[b % 2]:] == 0).all() and bev.grad[b, :keep[b % 2]].abs().sum() > 0
    with pytest.raises(AssertionError):
        G.get_distance_bev([lists[0] + lists[0][:1]] + lists[1:], bev)      # more clusters than objects


def test_point_box_alignment_vs_oracle_large_and_edge_cases():
    rng = np.random.default_rng(3)
    n_obj = 700                                        # > 500 objects, ragged, up to 6000 points
    counts = rng.choice([0, 0, 1, 2, 31, 32, 33, 500, 6000], n_obj)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    bev = np.stack([rng.uniform(0, 70, n_obj), rng.uniform(-40, 40, n_obj), rng.uniform(0.4, 5, n_obj),
                    rng.uniform(0.4, 2.5, n_obj), rng.uniform(-7, 7, n_obj)], 1).astype(np.float32)
    xy = np.concatenate([bev[i, :2] + rng.normal(0, 2.0, (counts[i], 2)) for i in range(n_obj)], 0).astype(np.float32)
    # points exactly on the centre / on a face of an axis-aligned box: abs'(0) = 0, relu'(0) = 0, first-min ties
    bev[0] = [10, 5, 4, 2, 0]
    xy[off[0]:off[1]] = 0
    k0 = 2 if counts[0] >= 2 else 0
    tb = torch.from_numpy(bev).clone().requires_grad_(True)
    rmin, rx, ry = ol.point_box_distances(torch.from_numpy(xy), off, tb)
    coef = torch.from_numpy(rng.uniform(0.5, 1.5, (n_obj, 3)).astype(np.float32))
    (torch.stack([rmin, rx, ry], 1) * coef).sum().backward()
    gb = torch.from_numpy(bev).cuda().requires_grad_(True)
    d = G.point_box_distances(torch.from_numpy(xy).cuda(), torch.from_numpy(off.astype(np.int32)).cuda(), gb)
    (d * coef.cuda()).sum().backward()
    ref = torch.stack([rmin, rx, ry], 1).detach().numpy()
    got = d.detach().cpu().numpy()
    b64 = d64(tb)
    m64, x64, y64 = ol.point_box_distances(d64(xy), off, b64)
    (torch.stack([m64, x64, y64], 1) * d64(coef)).sum().backward()
    assert close64(got, torch.stack([m64, x64, y64], 1), ref, what='PAL distances (large)')
    assert close64(gb.grad, b64.grad, tb.grad, what='PAL grad (large)')
    assert (got[counts == 0] == 0).all()
    # losses (weighted L1 against zero) agree with the oracle's mmdet restatement
    mask = torch.from_numpy((rng.uniform(size=(1, n_obj)) < 0.7).astype(np.float32))
    lo = ol.point_alignment_losses(rmin.detach()[None], rx.detach()[None], ry.detach()[None], mask)
    lg = G.point_alignment_losses(d[None, :, 0:1], d[None, :, 1:2], d[None, :, 2:3], mask.cuda())
    for a, b in zip(lo, lg):
        assert abs(float(a) - float(b)) <= 2e-5 * abs(float(a)) + 1e-7
