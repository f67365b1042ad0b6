"""Synthetic KITTI / SUN-RGBD shaped frames for tests and bench.

fp32, ``numpy.random.default_rng(1000 * cfg + frame)``.  Pure data generation (numpy on the
host); nothing here is on the measured path.
"""
import numpy as np

# KITTI 000000 calibration (rect / Trv2c: /root/reference/tests/test_utils/test_box_np_ops.py:8-15;
# P2 of the same frame; P2 @ rect @ Trv2c reproduces expected_lidar2img of
# /root/reference/tests/test_data/test_datasets/test_kitti_dataset.py:226-230).
KITTI_RECT = np.array([[0.9999128, 0.01009263, -0.00851193, 0.], [-0.01012729, 0.9999406, -0.00403767, 0.],
                       [0.00847068, 0.00412352, 0.9999556, 0.], [0., 0., 0., 1.]], dtype=np.float32)
KITTI_TRV2C = np.array([[0.00692796, -0.9999722, -0.00275783, -0.02457729],
                        [-0.00116298, 0.00274984, -0.9999955, -0.06127237],
                        [0.9999753, 0.00693114, -0.0011439, -0.3321029], [0., 0., 0., 1.]], dtype=np.float32)
KITTI_P2 = np.array([[707.0493, 0., 604.0814, 45.75831], [0., 707.0493, 180.5066, -0.3454157],
                     [0., 0., 1., 0.004981016], [0., 0., 0., 1.]], dtype=np.float32)
KITTI_IMG_HW = (375, 1242)
KITTI_PCD_RANGE = (0.0, -40.0, -3.0, 70.4, 40.0, 1.0)   # gga_kitti_config.py:2
KITTI_MATCH_RANGE = (0.0, -40.0, -3.0, 70.4, 40.0, 0.0)  # pcd_limit_range of the KITTI datasets

CLASS_DIMS = np.array([[3.9, 1.6, 1.56], [0.8, 0.6, 1.73], [1.76, 0.6, 1.73]], dtype=np.float32)

# the named configurations of BASELINE.json (cfg id -> shape)
CONFIGS = {
    1: dict(name='kitti_single_frame_cpu', N=120000, M=64, G=8, frames_per_gpu=1, kind='kitti'),
    2: dict(name='gga_kitti_train', N=120000, M=256, G=8, frames_per_gpu=8, kind='kitti'),
    3: dict(name='fcaf3d_sunrgbd', N=50000, M=512, G=8, frames_per_gpu=8, kind='sunrgbd'),
    4: dict(name='pseudo_label_matching', N=0, M=512, G=8, frames_per_gpu=464, kind='kitti'),
    5: dict(name='roofline_stress', N=2000000, M=1024, G=8, frames_per_gpu=1, kind='kitti'),
}


def kitti_lidar2img():
    return (KITTI_P2 @ KITTI_RECT @ KITTI_TRV2C).astype(np.float32)


def sunrgbd_depth2img():
    K = np.array([[529.5, 0, 365.0], [0, 529.5, 265.0], [0, 0, 1]], dtype=np.float32)
    t = np.deg2rad(6.0)
    # depth (x right, y forward, z up) -> camera (x right, y down, z forward), tilted about x
    Rt = np.array([[1, 0, 0], [0, -np.sin(t), -np.cos(t)], [0, np.cos(t), -np.sin(t)]], dtype=np.float32)
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = K @ Rt
    return m


def make_boxes(rng, M, kind='kitti'):
    if kind == 'sunrgbd':
        xyz = np.stack([rng.uniform(-4, 4, M), rng.uniform(1, 7, M), rng.uniform(-1.5, 0.5, M)], 1)
        dims = rng.uniform(0.3, 2.5, (M, 3))
    else:
        xyz = np.stack([rng.uniform(1, 69.4, M), rng.uniform(-39, 39, M), rng.uniform(-2, -1, M)], 1)
        dims = CLASS_DIMS[rng.integers(0, 3, M)] * rng.uniform(0.8, 1.2, (M, 3))
    yaw = rng.uniform(-np.pi, np.pi, (M, 1))
    boxes = np.concatenate([xyz, dims, yaw], 1).astype(np.float32)
    # a few axis-aligned boxes with exactly representable faces carry the adversarial points
    k = min(4, M)
    boxes[:k, 6] = 0.0
    boxes[:k, 3:6] = np.float32([4.0, 2.0, 1.5])
    boxes[:k, :3] = np.round(boxes[:k, :3] * 4) / 4
    return boxes


def make_points(rng, N, boxes, kind='kitti', sort_azimuth=True):
    """[N, 4] (x, y, z, r): 80 % uniform in range, 20 % inside 1.2x enlarged boxes, plus up to
    64 adversarial points exactly on faces / edges of the yaw-0 boxes."""
    M = boxes.shape[0]
    n_adv = min(64, N // 8) if M >= 1 else 0
    n_in = int(0.2 * N) if M >= 1 else 0
    n_bg = N - n_in - n_adv
    if kind == 'sunrgbd':
        bg = np.stack([rng.uniform(-5, 5, n_bg), rng.uniform(0, 8, n_bg), rng.uniform(-2, 1, n_bg)], 1)
    else:
        bg = np.stack([rng.uniform(0, 70.4, n_bg), rng.uniform(-40, 40, n_bg), rng.uniform(-3, 1, n_bg)], 1)
    parts = [bg]
    if n_in:
        sel = rng.integers(0, M, n_in)
        loc = rng.uniform(-0.6, 0.6, (n_in, 3)) * boxes[sel, 3:6]
        c, s = np.cos(boxes[sel, 6]), np.sin(boxes[sel, 6])
        parts.append(np.stack([boxes[sel, 0] + loc[:, 0] * c - loc[:, 1] * s,
                               boxes[sel, 1] + loc[:, 0] * s + loc[:, 1] * c,
                               boxes[sel, 2] + boxes[sel, 5] * 0.5 + loc[:, 2]], 1))
    if n_adv:
        k = min(4, M)
        adv = []
        for j in range(n_adv):
            b = boxes[j % k]
            hx, hy, dz = b[3] / 2, b[4] / 2, b[5]
            case = (j // k) % 8
            off = [(hx, 0, dz / 2), (-hx, 0, dz / 2), (0, hy, dz / 2), (0, -hy, dz / 2),   # open x/y faces
                   (0, 0, 0), (0, 0, dz),                                                  # closed z faces
                   (hx, hy, dz), (hx * 0.5, hy * 0.5, dz * 0.5)][case]
            adv.append([b[0] + off[0], b[1] + off[1], b[2] + off[2]])
        parts.append(np.asarray(adv))
    xyz = np.concatenate(parts, 0).astype(np.float32)
    pts = np.concatenate([xyz, rng.uniform(0, 1, (N, 1)).astype(np.float32)], 1)
    if sort_azimuth:
        pts = pts[np.argsort(np.arctan2(pts[:, 1], pts[:, 0]), kind='stable')]
    else:
        pts = pts[rng.permutation(N)]
    return np.ascontiguousarray(pts, dtype=np.float32)


def _project_kitti_cam_np(boxes):
    """numpy restatement of variant B (target generation only)."""
    b = boxes.astype(np.float32).copy()
    two_pi = np.float32(2 * np.pi)
    b[:, 6] = b[:, 6] - np.floor(b[:, 6] / two_pi + np.float32(0.5)) * two_pi
    rt = KITTI_RECT @ KITTI_TRV2C
    xyz = np.concatenate([b[:, :3], np.ones((len(b), 1), np.float32)], 1) @ rt.T
    dims = b[:, [3, 5, 4]]
    yaw = -b[:, 6] - np.float32(np.pi / 2)
    yaw = yaw - np.floor(yaw / two_pi + np.float32(0.5)) * two_pi
    bits = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 1], [0, 1, 0], [1, 0, 0], [1, 0, 1], [1, 1, 1], [1, 1, 0]],
                    np.float32) - np.float32([0.5, 1.0, 0.5])
    loc = dims[:, None, :] * bits[None]
    c, s = np.cos(yaw)[:, None], np.sin(yaw)[:, None]
    cor = np.stack([loc[..., 0] * c + loc[..., 2] * s, loc[..., 1], -loc[..., 0] * s + loc[..., 2] * c], -1)
    cor = cor + xyz[:, None, :3]
    q = np.concatenate([cor, np.ones(cor.shape[:2] + (1,), np.float32)], -1) @ KITTI_P2.T
    uv = q[..., :2] / q[..., 2:3]
    return np.concatenate([uv.min(1), uv.max(1)], 1).astype(np.float32)


def make_targets(rng, boxes, G, img_hw=KITTI_IMG_HW):
    """[G, 4] 2D boxes: variant-B projection of a jittered copy of G boxes, clamped."""
    G = min(G, boxes.shape[0])
    jb = boxes[:G].copy()
    jb[:, :3] += rng.normal(0, 0.2, (G, 3)).astype(np.float32)
    jb[:, 6] += rng.normal(0, 0.1, G).astype(np.float32)
    t = _project_kitti_cam_np(jb)
    H, W = img_hw
    t[:, 0] = np.clip(t[:, 0], 0, W - 2); t[:, 1] = np.clip(t[:, 1], 0, H - 2)
    t[:, 2] = np.clip(t[:, 2], t[:, 0] + 1, W); t[:, 3] = np.clip(t[:, 3], t[:, 1] + 1, H)
    return t.astype(np.float32)


def make_frame(cfg, frame, N=None, M=None, G=None, sort_azimuth=True):
    """One synthetic frame of configuration `cfg` (1..5)."""
    c = CONFIGS[cfg]
    N = c['N'] if N is None else N
    M = c['M'] if M is None else M
    G = c['G'] if G is None else G
    kind = c['kind']
    rng = np.random.default_rng(1000 * cfg + frame)
    boxes = make_boxes(rng, M, kind)
    pts = make_points(rng, N, boxes, kind, sort_azimuth) if N else np.zeros((0, 4), np.float32)
    if kind == 'sunrgbd':
        l2i = sunrgbd_depth2img()
        hw = (530, 730)
        tg = np.stack([rng.uniform(0, 600, G), rng.uniform(0, 400, G)], 1)
        tg = np.concatenate([tg, tg + rng.uniform(20, 200, (G, 2))], 1).astype(np.float32)
    else:
        l2i = kitti_lidar2img()
        hw = KITTI_IMG_HW
        tg = make_targets(rng, boxes, G)
    target = tg[np.arange(M) % len(tg)] if len(tg) else np.zeros((M, 4), np.float32)
    return dict(points=pts, boxes=boxes, lidar2img=l2i, img_hw=hw, gt2d=tg,
                target=np.ascontiguousarray(target, dtype=np.float32),
                weight=np.ones((M,), np.float32))


def make_batch(cfg, frame0, n_frames, **kw):
    """Stacks `n_frames` frames: points [F, N, 4], boxes [F, M, 7], lidar2img [F, M, 4, 4]
    (one calib per object, the GGA_lidar2img layout), target [F, M, 4], weight [F, M]."""
    fr = [make_frame(cfg, frame0 + i, **kw) for i in range(n_frames)]
    M = fr[0]['boxes'].shape[0]
    return dict(
        points=np.stack([f['points'] for f in fr]),
        boxes=np.stack([f['boxes'] for f in fr]),
        lidar2img=np.stack([np.repeat(f['lidar2img'][None], M, 0) for f in fr]),
        target=np.stack([f['target'] for f in fr]),
        weight=np.stack([f['weight'] for f in fr]),
        gt2d=[f['gt2d'] for f in fr],
        img_hw=fr[0]['img_hw'])


# ---------------------------------------------------------------------------------------------
# target packing (CenterHead_GGA.get_targets_single inputs, centerpoint_head_gga.py:401-413)
# ---------------------------------------------------------------------------------------------
KITTI_TRAIN_CFG = dict(grid_size=[1408, 1600, 40], out_size_factor=8, voxel_size=[0.05, 0.05, 0.1],
                       point_cloud_range=[0, -40, -3, 70.4, 40, 1], dense_reg=1, gaussian_overlap=0.1,
                       max_objs=500, min_radius=2)      # configs/gga/gga_kitti_config.py:63-75
KITTI_TASKS = [['Pedestrian'], ['Cyclist'], ['Car']]     # configs/gga/gga_kitti_config.py:39-43


def make_target_frame(rng, n, n_classes=3, dtype=np.float64, adversarial=True):
    """One frame of GGA annotations: labels int64 [n] (a few -1 = ignored), boxes_img [n,4],
    lidar2img float32 [n,4,4] (per-object calib), pseudo `dtype` [n,7], bdry bool [n,4],
    base lidar2img [4,4].  `adversarial` adds objects with degenerate extents and centres on /
    outside the map borders (cell coordinate in (-1, 0), last cell, beyond)."""
    boxes = make_boxes(rng, n).astype(np.float64) if n else np.zeros((0, 7))
    boxes[:, 0:2] += rng.normal(0, 1e-3, (n, 2))
    labels = rng.integers(0, n_classes, n).astype(np.int64)
    if adversarial and n >= 12:
        labels[rng.integers(0, n, 2)] = -1
        boxes[0, 3] = 0.0                     # zero width: slot stays empty
        boxes[1, 4] = -1.0                    # negative length
        boxes[2, 0] = -0.2                    # cell coordinate in (-1, 0): truncates to cell 0
        boxes[3, 0] = 70.39                   # last column
        boxes[4, 1] = 39.99                   # last row
        boxes[5, 0] = 70.41                   # one cell beyond
        boxes[6, 1] = -40.5                   # below the map
        boxes[7, 0:2] = (0.1, -39.9)          # corner: window clipped on two sides
        boxes[8, 3:5] = (12.0, 30.0)          # very large object: radius >> min_radius
        boxes[9, 3:5] = (0.05, 0.05)          # tiny: min_radius applies
        boxes[10, 0:2] = (35.2, 0.0)          # exactly on a cell edge
    l2i = kitti_lidar2img()
    lidar2img = np.broadcast_to(l2i, (n, 4, 4)).copy()
    lidar2img[::3, :3, 3] += rng.normal(0, 0.05, (len(range(0, n, 3)), 3)).astype(np.float32)
    x1y1 = np.stack([rng.uniform(0, 1100, n), rng.uniform(0, 300, n)], 1)
    boxes_img = np.concatenate([x1y1, x1y1 + rng.uniform(8, 200, (n, 2))], 1).astype(np.float32)
    bdry = rng.random((n, 4)) < 0.15
    return dict(labels=labels, boxes_img=boxes_img, lidar2img=lidar2img.astype(np.float32), pseudo=boxes.astype(dtype),
                bdry=bdry, base_lidar2img=l2i)


def make_detection_frames(seed, counts, gts=6):
    """Synthetic inputs of the result-formatting / pseudo-label rewrite step: per frame a KITTI
    info dict (calib float64 4x4 like the info pkls, image idx / shape, annos with GGA fields and
    trailing DontCare objects) and a detection dict (boxes_3d [n, 7] LiDAR bottom-centre boxes, some
    outside the image / range, scores_3d, labels_3d)."""
    rng = np.random.default_rng(seed)
    infos, dets = [], []
    names = np.array(['Pedestrian', 'Cyclist', 'Car'])
    for f, n in enumerate(counts):
        P2 = KITTI_P2.astype(np.float64).copy()
        P2[0, 3] += rng.normal(0, 1.0)
        calib = dict(R0_rect=KITTI_RECT.astype(np.float64), Tr_velo_to_cam=KITTI_TRV2C.astype(np.float64), P2=P2)
        boxes = make_boxes(rng, n).astype(np.float32) if n else np.zeros((0, 7), np.float32)
        boxes[:, 6] = rng.uniform(-7, 7, n)
        if n >= 6:
            boxes[0, 0] = -5.0          # behind the sensor: outside the range and the image
            boxes[1, 1] = 45.0          # outside the point-cloud range
            boxes[2, 0:2] = (3.0, 9.0)  # in range, projects left of the image
            boxes[3, 0:2] = (6.0, -2.0) # close: box clipped by the image borders
        g = min(gts, max(n, 1)) + 2
        gb = (boxes[rng.integers(0, max(n, 1), g)] if n else make_boxes(rng, g)).copy()
        gb[:, :3] += rng.normal(0, 0.3, (g, 3)).astype(np.float32)
        bbox = _project_kitti_cam_np(gb).astype(np.float64)
        bbox[:, [0, 2]] = np.clip(bbox[:, [0, 2]], 0, 1242)
        bbox[:, [1, 3]] = np.clip(bbox[:, [1, 3]], 0, 375)
        gname = names[rng.integers(0, 3, g)].astype('<U10')
        gname[-2:] = 'DontCare'
        if g > 4:
            gname[1] = 'Van'            # a class the matching drops
        annos = dict(name=gname, bbox=bbox, truncated=rng.uniform(0, 1, g), occluded=rng.integers(0, 3, g),
                     alpha=rng.uniform(-3, 3, g), dimensions=rng.uniform(0.5, 4, (g, 3)),
                     location=rng.uniform(-10, 40, (g, 3)), rotation_y=rng.uniform(-3, 3, g),
                     index=np.arange(g, dtype=np.int32), group_ids=np.arange(g, dtype=np.int32),
                     difficulty=rng.integers(-1, 3, g).astype(np.int32),
                     num_points_in_gt=rng.integers(0, 400, g).astype(np.int32),
                     GGA_init_pseudo_label=rng.uniform(-1, 1, (g, 7)), GGA_boxes_img=bbox.copy(),
                     GGA_bdry_masks=rng.random((g, 4)) < 0.2,
                     GGA_in_box_points=[np.ones((1 + i, 4)) for i in range(g)])
        infos.append(dict(image=dict(image_idx=100 + 7 * f, image_shape=np.array([375, 1242], np.int32)), calib=calib,
                          annos=annos))
        dets.append(dict(boxes_3d=boxes, scores_3d=rng.uniform(0.05, 1, n).astype(np.float32),
                         labels_3d=rng.integers(0, 3, n).astype(np.int64)))
    return infos, dets
