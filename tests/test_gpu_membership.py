"""GPU parity tests, part 1: the CUDA membership kernels (through the C ABI) against the CPU
oracle — bit-exact — on the reference's golden vectors, adversarial boundary inputs, synthetic
KITTI / SUN-RGBD frames, and full-size frames through size-independent properties."""
import numpy as np
import pytest
import torch

import gga_b200 as G
from gga_b200 import synth
from oracle import membership as om
from test_oracle_membership import (CAM_ALL, CAM_PART, DEPTH_ALL, DEPTH_BOXES, DEPTH_PART, DEPTH_PTS,
                                    LIDAR_ALL, LIDAR_BOXES, LIDAR_PART, LIDAR_PTS)

pytestmark = pytest.mark.gpu


def cu(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype).cuda()


def run_all(points, boxes):
    """points [M, C], boxes [T, 7] numpy -> (all int32 [M,T], part [M], bits-unpacked [M,T])"""
    p3 = cu(points[:, :3])[None]
    b = cu(boxes)[None]
    a = G.points_in_boxes_all(p3, b)[0].cpu().numpy()
    part = G.points_in_boxes_part(p3, b)[0].cpu().numpy()
    bits = G.points_in_boxes_bits(cu(points)[None], b)
    ub = G.unpack_bits(bits, boxes.shape[0])[0].cpu().numpy()
    return a, part, ub, bits[0].cpu().numpy()


def check_against_oracle(points, boxes):
    a, part, ub, bits = run_all(points, boxes)
    ref = om.points_in_boxes_all_np(points[:, :3], boxes, nthreads=8)
    assert np.array_equal(a, ref)
    assert np.array_equal(ub, ref)
    assert np.array_equal(part, np.where(ref.any(1), ref.argmax(1), -1))
    # padding bits are zero
    W = bits.shape[1]
    T = boxes.shape[0]
    if W * 32 > T:
        full = np.unpackbits(bits.view(np.uint8), axis=1, bitorder='little')
        assert full[:, T:].sum() == 0
    return ref


def test_device_detmath_equals_host_build():
    rng = np.random.default_rng(11)
    bits = np.concatenate([np.arange(0, 2 ** 32, 4099, dtype=np.uint64),
                           rng.integers(0, 2 ** 32, 1 << 20, dtype=np.uint64)]).astype(np.uint32)
    x = np.concatenate([bits.view(np.float32), rng.uniform(-8, 8, 1 << 20).astype(np.float32)])
    xs = cu(x)
    s = torch.empty_like(xs)
    c = torch.empty_like(xs)
    L = G._lib.load()
    G._lib.check(L.gga_test_sincos(xs.data_ptr(), xs.numel(), s.data_ptr(), c.data_ptr(), None))
    torch.cuda.synchronize()
    hs, hc = om.det_sincos(x)
    ds, dc = s.cpu().numpy(), c.cpu().numpy()
    fin = np.isfinite(x)
    assert np.array_equal(ds.view(np.uint32)[fin], hs.view(np.uint32)[fin])
    assert np.array_equal(dc.view(np.uint32)[fin], hc.view(np.uint32)[fin])
    assert np.isnan(ds[~fin]).all() and np.isnan(dc[~fin]).all()


def test_box_prep_matches_contract_terms():
    rng = np.random.default_rng(5)
    boxes = synth.make_boxes(rng, 512)
    boxes[:4, 3:6] = [[1e-45, 3e-45, 5e-45], [0, 1, 1], [-1, 2, 2], [np.inf, 1, 1]]
    b = cu(boxes)
    prep = torch.empty((512, 8), device='cuda')
    G._lib.check(G._lib.load().gga_test_box_prep(b.data_ptr(), 512, prep.data_ptr(), None))
    p = prep.cpu().numpy()
    cosa, sina = om.libm_box_sincos(boxes[:, 6])
    assert np.array_equal(p[:, 4].view(np.uint32), cosa.view(np.uint32))
    assert np.array_equal(p[:, 5].view(np.uint32), sina.view(np.uint32))
    cz = (boxes[:, 2].astype(np.float64) + boxes[:, 5].astype(np.float64) / 2.0).astype(np.float32)
    assert np.array_equal(p[:, 2].view(np.uint32), cz.view(np.uint32))
    # subnormal extents: the fp32 thresholds reproduce the double compares
    assert p[0, 6] >= np.float64(boxes[0, 3]) / 2 and p[0, 3] <= np.float64(boxes[0, 5]) / 2


def test_reference_golden_lidar_depth():
    a, part, ub, _ = run_all(np.float32(LIDAR_PTS), np.float32(LIDAR_BOXES))
    assert np.array_equal(a, np.array(LIDAR_ALL)) and np.array_equal(ub, np.array(LIDAR_ALL))
    assert np.array_equal(part, np.array(LIDAR_PART))
    a, part, ub, _ = run_all(np.float32(DEPTH_PTS), np.float32(DEPTH_BOXES))
    assert np.array_equal(a, np.array(DEPTH_ALL)) and np.array_equal(part, np.array(DEPTH_PART))


@pytest.mark.refonly
def test_reference_own_box_classes_with_cuda_op_injected():
    """Drop-in acceptance: the reference's box classes + our op reproduce test_box3d.py:1683-1797."""
    from oracle import ref_loader
    ref = ref_loader.load_reference(G.points_in_boxes_all, G.points_in_boxes_part)
    six = torch.tensor(DEPTH_BOXES + LIDAR_BOXES, dtype=torch.float32).cuda()
    cam_boxes = ref.DepthInstance3DBoxes(six).convert_to(ref.Box3DMode.CAM)
    cam_pts = ref.DepthPoints(torch.tensor(DEPTH_PTS + LIDAR_PTS, dtype=torch.float32).cuda()).convert_to(
        ref.Coord3DMode.CAM).tensor
    assert np.array_equal(cam_boxes.points_in_boxes_all(cam_pts).cpu().numpy(), np.array(CAM_ALL))
    assert np.array_equal(cam_boxes.points_in_boxes_part(cam_pts).cpu().numpy(), np.array(CAM_PART))
    lb = ref.LiDARInstance3DBoxes(torch.tensor(LIDAR_BOXES, dtype=torch.float32).cuda())
    assert np.array_equal(lb.points_in_boxes_all(torch.tensor(LIDAR_PTS).cuda()).cpu().numpy(),
                          np.array(LIDAR_ALL))


def test_camera_golden_through_coordinate_conversion():
    """Same golden without the reference tree: CAM->LiDAR conversion of cam_box3d.py:330-354
    restated inline (points and box xyz rotated by [[0,0,1],[-1,0,0],[0,-1,0]], dims/yaw kept)."""
    six = np.float32(DEPTH_BOXES + LIDAR_BOXES)
    pts = np.float32(DEPTH_PTS + LIDAR_PTS)
    d2c = np.float32([[1, 0, 0], [0, 0, -1], [0, 1, 0]])       # DEPTH -> CAM (box_3d_mode.py:131)
    c2l = np.float32([[0, 0, 1], [-1, 0, 0], [0, -1, 0]])      # CAM -> LIDAR (coord_3d_mode.py)
    cam_xyz = six[:, :3] @ d2c.T
    cam = np.concatenate([cam_xyz, six[:, [3, 5, 4]], -six[:, 6:7]], 1)
    cam_pts = pts @ d2c.T
    boxes_l = np.concatenate([cam[:, :3] @ c2l.T, cam[:, 3:]], 1).astype(np.float32)
    pts_l = (cam_pts @ c2l.T).astype(np.float32)
    a, part, _, _ = run_all(pts_l, boxes_l)
    assert np.array_equal(a, np.array(CAM_ALL)) and np.array_equal(part, np.array(CAM_PART))


@pytest.mark.parametrize('M,T', [(1, 1), (31, 7), (1000, 32), (1025, 33), (4097, 64), (3000, 65), (2048, 128),
                                 (5000, 129), (7777, 256), (3001, 257), (2000, 700), (1500, 1024), (600, 1500),
                                 (4096, 300), (8192, 512), (2048, 513), (2016, 768), (6400, 1024)])
def test_random_shapes_bit_exact(M, T):
    rng = np.random.default_rng(M * 31 + T)
    boxes = synth.make_boxes(rng, T)
    boxes[:, 6] = rng.uniform(-10, 10, T)
    pts = synth.make_points(rng, M, boxes, sort_azimuth=bool(T % 2))
    ref = check_against_oracle(pts, boxes)
    if M >= 1000:
        assert ref.sum() > 0


@pytest.mark.parametrize('T', [40, 300, 700, 1100])
def test_adversarial_boxes_and_points(T):
    """T = 40: cell-table path; 300 / 700: axis masks (x, y and z) with 16- / 24-word rows; 1100: two sweeps."""
    rng = np.random.default_rng(99)
    boxes = synth.make_boxes(rng, T)
    boxes[4] = [10, 0, -1, 0, 2, 2, 0.3]            # zero width: contains nothing
    boxes[5] = [10, 0, -1, -2, 2, 2, 0.3]           # negative size
    boxes[6] = [np.nan, 0, -1, 2, 2, 2, 0.3]        # NaN centre
    boxes[7] = [10, 0, -1, 2, 2, 2, np.nan]         # NaN yaw
    boxes[8] = [10, 0, -1, 2, 2, 2, np.inf]         # inf yaw
    boxes[9] = [10, 0, -1, np.inf, 2, 2, 0.0]       # infinite extent: a slab across the scene
    boxes[10] = [35, 0, -1, 500, 500, 50, 0.7]      # covers everything
    boxes[11] = [10, 0, -1, 2, 2, np.inf, 0.1]      # infinite height
    boxes[12] = [10, 0, -1, 2, 2, np.nan, 0.1]      # NaN height: z test passes for every point
    boxes[13] = [1e30, 0, -1, 1e30, 2, 2, 0.0]      # overflowing rectangle
    boxes[14] = [10, 0, -1, 2, 2, -1, 0.1]          # negative height
    boxes[15] = [3e38, 3e38, 0, 3e38, 3e38, 1, 0.5]
    boxes[16] = [20, 5, -1.5, 3, 1.5, 1.5, 1e4]     # huge yaw (Payne-Hanek path)
    boxes[17] = [20, 5, -1.5, 3, 1.5, 1.5, -3e38]
    # z slabs: far above / below the others, stacked at one place, a subnormal height, non-finite z
    boxes[18] = [30, 10, 40, 4, 4, 2, 0.2]
    boxes[19] = [30, 10, -60, 4, 4, 2, 0.2]
    boxes[20] = [30, 10, 0, 4, 4, 1, 0.2]
    boxes[21] = [30, 10, 1, 4, 4, 1, 0.2]           # its bottom face is box 20's top face (both closed)
    boxes[22] = [30, 10, 2, 4, 4, 1e-42, 0.2]       # subnormal height: only z == 2 exactly
    boxes[23] = [30, 10, np.inf, 4, 4, 1, 0.2]
    boxes[24] = [30, 10, -np.inf, 4, 4, np.inf, 0.2]
    boxes[25] = [30, 10, 1e30, 4, 4, 1e30, 0.2]
    boxes[26] = [30, 10, np.nan, 4, 4, 1, 0.2]      # NaN z: the z test passes for every point
    pts = synth.make_points(rng, 20000, boxes)
    pts[20:34, :3] = [[30, 10, 40], [30, 10, 42], [30, 10, 42.00001], [30, 10, -60], [30, 10, -58], [30, 10, 1],
                      [30, 10, 0], [30, 10, 2], [30, 10, np.nan], [30, 10, 1e30], [30, 10, 2e30], [30, 10, 3e38],
                      [30, 10, -3e38], [30, 10, np.float32(2) + np.float32(2.4e-7)]]
    pts[:6, :3] = [[np.nan, 0, 0], [0, np.nan, 0], [10, 0, np.nan], [np.inf, 0, 0], [0, -np.inf, 0],
                   [10, 0, np.inf]]
    pts[6:12, :3] = [[10, 0, -1], [10, 0, 1], [11, 0, 0], [9, 0, 0], [1e30, 0, -0.5], [20, 5, -1]]
    ref = check_against_oracle(pts, boxes)
    assert ref[:, 10].sum() > 10000 and ref[:, 4].sum() == 0 and ref[:, 6].sum() == 0
    assert ref[2, 12] == 1  # NaN z inside the NaN-height box at its centre (contract quirk)
    assert ref[20, 18] and ref[21, 18] and not ref[22, 18] and ref[23, 19] and ref[24, 19]   # closed z faces
    assert ref[25, 20] and ref[25, 21] and ref[27, 21] and ref[27, 22] and ref[28, 20]        # shared face, NaN z
    assert ref[20:34, 26].all() and ref[30, 25]


def test_only_degenerate_boxes_and_empty_inputs():
    boxes = np.float32([[0, 0, 0, 0, 0, 0, 0], [1, 1, 1, -1, -1, -1, 1]])
    pts = np.float32(np.random.default_rng(0).normal(size=(500, 4)))
    ref = check_against_oracle(pts, boxes)
    assert ref.sum() == 0
    # empty points / empty boxes
    e = G.points_in_boxes_all(torch.zeros((1, 0, 3)).cuda(), cu(boxes)[None])
    assert tuple(e.shape) == (1, 0, 2)
    e = G.points_in_boxes_all(cu(pts[:, :3])[None], torch.zeros((1, 0, 7)).cuda())
    assert tuple(e.shape) == (1, 500, 0)
    e = G.points_in_boxes_part(cu(pts[:, :3])[None], torch.zeros((1, 0, 7)).cuda())
    assert (e == -1).all() and tuple(e.shape) == (1, 500)


def test_batched_frames_and_strided_views():
    fr = [synth.make_frame(2, i, N=5000, M=200) for i in range(5)]
    P = cu(np.stack([f['points'] for f in fr]))          # [5, 5000, 4]
    B = cu(np.stack([f['boxes'] for f in fr]))
    ref = np.stack([om.points_in_boxes_all_np(f['points'], f['boxes'], 8) for f in fr])
    assert np.array_equal(G.points_in_boxes_all(P[..., :3], B).cpu().numpy(), ref)          # strided view
    assert np.array_equal(G.points_in_boxes_all(P[..., :3].contiguous(), B).cpu().numpy(), ref)
    assert np.array_equal(G.points_in_boxes_all(P[..., 1:4].contiguous(), B).cpu().numpy(),
                          np.stack([om.points_in_boxes_all_np(f['points'][:, 1:4], f['boxes'], 8) for f in fr]))
    bits = G.points_in_boxes_bits(P, B)
    assert np.array_equal(G.unpack_bits(bits, 200).cpu().numpy(), ref)
    part = G.points_in_boxes_part(P[..., :3], B).cpu().numpy()
    assert np.array_equal(part, np.where(ref.any(2), ref.argmax(2), -1))
    # the mmcv shape asserts
    with pytest.raises(AssertionError):
        G.points_in_boxes_all(P, B)                       # last dim 4
    with pytest.raises(AssertionError):
        G.points_in_boxes_all(P[:2, :, :3], B)


def test_points_in_boxes_cpu_signature_host_tensors():
    f = synth.make_frame(1, 3, N=20000)
    p = torch.from_numpy(f['points'][None, :, :3].copy())
    b = torch.from_numpy(f['boxes'][None])
    out = G.points_in_boxes_cpu(p, b)
    assert not out.is_cuda and out.dtype == torch.int32 and tuple(out.shape) == (1, 20000, 64)
    assert np.array_equal(out[0].numpy(), om.points_in_boxes_all_np(f['points'], f['boxes'], 8))


@pytest.mark.parametrize('T,frames,M', [(200, 3, 4000), (256, 5, 4096), (400, 3, 2048), (512, 2, 4128),
                                        (700, 3, 2080), (1024, 2, 3200), (1024, 3, 999), (256, 3, 4001), (400, 2, 1000),
                                        (130, 4, 33), (512, 9, 31)])
def test_multi_frame_bits_lean_widths(T, frames, M):
    """[F, N, 4] points, every row width of the lean stream variants (8/16/24/32 words), frames
    with different boxes in one call."""
    rng = np.random.default_rng(T * 7 + frames)
    bx, px = [], []
    for f in range(frames):
        b = synth.make_boxes(rng, T)
        b[:, 6] = rng.uniform(-7, 7, T)
        bx.append(b)
        px.append(synth.make_points(rng, M, b, sort_azimuth=bool(f % 2)))
    bits = G.points_in_boxes_bits(cu(np.stack(px)), cu(np.stack(bx)))
    ub = G.unpack_bits(bits, T).cpu().numpy()
    for f in range(frames):
        ref = om.points_in_boxes_all_np(px[f][:, :3], bx[f], nthreads=8)
        assert np.array_equal(ub[f], ref), f
        assert ref.any()


@pytest.mark.parametrize('cfg,frames', [(1, 1), (2, 2), (3, 1)])
def test_config_shapes_bit_exact(cfg, frames):
    for i in range(frames):
        f = synth.make_frame(cfg, i)
        ref = check_against_oracle(f['points'], f['boxes'])
        assert ref.any(1).mean() > 0.05


@pytest.mark.parametrize('T,N', [(8000, 700), (32768 + 5, 300)])
def test_many_boxes_are_served_in_chunks(T, N):
    """More boxes than one sweep indexes (1024): the mmcv op accepts any T, so does this one."""
    rng = np.random.default_rng(T)
    boxes = synth.make_boxes(rng, T)
    pts = synth.make_points(rng, N, boxes[:2000])
    ref = om.points_in_boxes_all_np(pts[:, :3], boxes, nthreads=8)
    P, B = cu(pts)[None], cu(boxes)[None]
    assert np.array_equal(G.unpack_bits(G.points_in_boxes_bits(P, B), T)[0].cpu().numpy(), ref)
    part = G.points_in_boxes_part(P[..., :3], B)[0].cpu().numpy()
    assert np.array_equal(part, np.where(ref.any(1), ref.argmax(1), -1))
    if T <= 8000:
        assert np.array_equal(G.points_in_boxes_all(P[..., :3], B)[0].cpu().numpy(), ref)


def test_extreme_extents_keep_the_index_conservative():
    """Box extents spanning many orders of magnitude (one far-away box stretches the bin grid so
    that every other box shares a bin) and points outside every bin."""
    rng = np.random.default_rng(123)
    boxes = synth.make_boxes(rng, 200)
    boxes[0, :2] = [5e6, -7e6]
    boxes[1, :2] = [-3e7, 2e7]
    boxes[2, 3:5] = [1e5, 3e4]
    pts = synth.make_points(rng, 6000, boxes)
    pts[:50, :2] = rng.uniform(-1e7, 1e7, (50, 2))
    pts[50:60, :3] = boxes[:10, :3] + [0, 0, 0.5]
    ref = check_against_oracle(pts, boxes)
    assert ref[50:60].any()


def test_full_size_stress_properties():
    """Config 5 (2M points x 1024 boxes): size-independent checks — sampled rows against the
    oracle, permutation invariance over points, and popcount = sum over boxes."""
    f = synth.make_frame(5, 0)
    P, B = cu(f['points'])[None], cu(f['boxes'])[None]
    bits = G.points_in_boxes_bits(P, B)
    rng = np.random.default_rng(1)
    idx = np.sort(rng.choice(f['points'].shape[0], 30000, replace=False))
    ref = om.points_in_boxes_all_np(f['points'][idx], f['boxes'], 8)
    got = G.unpack_bits(bits[:, torch.as_tensor(idx).cuda()], 1024)[0].cpu().numpy()
    assert np.array_equal(got, ref)
    perm = torch.randperm(P.shape[1], device='cuda')
    bits_p = G.points_in_boxes_bits(P[:, perm].contiguous(), B)
    assert torch.equal(bits_p, bits[:, perm])
    part = G.points_in_boxes_part(P[..., :3], B)
    has = (bits != 0).any(-1)
    assert torch.equal(part >= 0, has)
