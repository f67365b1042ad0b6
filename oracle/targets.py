"""TEST INFRASTRUCTURE — CPU restatement of GGA's training-target packing.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU baseline may import this; the product never does.

Follows ``CenterHead_GGA.get_targets_single``
(/root/reference/mmdet3d/models/dense_heads/centerpoint_head_gga.py:401-627) and
``gaussian_2d / draw_heatmap_gaussian / gaussian_radius``
(/root/reference/mmdet3d/core/utils/gaussian.py:6-86) in numpy scalars of the dtype torch's
type promotion gives the reference's 0-dim operands (fp32 pseudo labels -> fp32, fp64 -> fp64).

Parity: pinned.  tests/golden/ref_targets.npz holds outputs of the reference's own
``get_targets_single`` source text executed by oracle/gen_golden.py (fp32 and fp64 pseudo
labels, multi-class tasks, out-of-range / degenerate / overflowing objects);
tests/test_oracle_targets.py checks this restatement against them bit for bit.
"""
import numpy as np


def gaussian_radius(height, width, min_overlap):
    """gaussian.py:58-86; `height`, `width` numpy scalars (np.float32 / np.float64); the Python
    scalars are evaluated in double first and cast to the tensor dtype, like torch does."""
    T = type(height)
    b1 = height + width
    c1 = width * height * T(1 - min_overlap) / T(1 + min_overlap)
    sq1 = np.sqrt(b1 * b1 - T(4) * c1)
    r1 = (b1 + sq1) / T(2)
    b2 = T(2) * (height + width)
    c2 = T(1 - min_overlap) * width * height
    sq2 = np.sqrt(b2 * b2 - T(16) * c2)
    r2 = (b2 + sq2) / T(2)
    a3 = 4 * min_overlap
    b3 = T(-2 * min_overlap) * (height + width)
    c3 = T(min_overlap - 1) * width * height
    sq3 = np.sqrt(b3 * b3 - T(4 * a3) * c3)
    r3 = (b3 + sq3) / T(2)
    r = r1
    if r2 < r:
        r = r2
    if r3 < r:
        r = r3
    return r


def gaussian_2d(radius):
    """gaussian.py:6-22 for shape (2r+1, 2r+1), sigma = (2r+1)/6; float64."""
    d = 2 * radius + 1
    sigma = d / 6
    y, x = np.ogrid[-radius:radius + 1, -radius:radius + 1]
    h = np.exp(-(x * x + y * y).astype(np.float64) / (2 * sigma * sigma))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    return h


def draw_heatmap_gaussian(heatmap, cx, cy, radius):
    """gaussian.py:25-55 (k = 1) on a float32 [H, W] numpy map, in place."""
    g = gaussian_2d(radius)
    H, W = heatmap.shape
    left, right = min(cx, radius), min(W - cx, radius + 1)
    top, bottom = min(cy, radius), min(H - cy, radius + 1)
    mh = heatmap[cy - top:cy + bottom, cx - left:cx + right]
    mg = g[radius - top:radius + bottom, radius - left:radius + right].astype(np.float32)
    if min(mg.shape) > 0 and min(mh.shape) > 0:
        np.maximum(mh, mg, out=mh)


def get_targets_single(labels, boxes_img, lidar2img, pseudo, bdry, base_lidar2img, srl, class_names, train_cfg):
    """One frame.  labels int [n]; boxes_img [n,4]; lidar2img [n,4,4]; pseudo [n,7] float32 or
    float64; bdry bool [n,4]; base_lidar2img [4,4]; srl [n_tasks].  Returns per-task lists
    (heatmaps, anno_boxes, inds, masks, anno_lidar2imgs, src_index, boundary_masks)."""
    T = np.float64 if pseudo.dtype == np.float64 else np.float32
    pseudo = pseudo.astype(T)
    K = int(train_cfg['max_objs']) * int(train_cfg['dense_reg'])
    osf = int(train_cfg['out_size_factor'])
    fm_w, fm_h = int(train_cfg['grid_size'][0]) // osf, int(train_cfg['grid_size'][1]) // osf
    vx, vy = (T(np.float32(v)) for v in train_cfg['voxel_size'][:2])
    x0, y0 = (T(np.float32(v)) for v in train_cfg['point_cloud_range'][:2])
    ov = float(train_cfg['gaussian_overlap'])
    outs = ([], [], [], [], [], [], [])
    flag = 0
    for t, names in enumerate(class_names):
        order = [i for c in range(len(names)) for i in np.nonzero(labels == c + flag)[0]]
        cls = [c for c in range(len(names)) for _ in np.nonzero(labels == c + flag)[0]]
        flag += len(names)
        heat = np.zeros((len(names), fm_h, fm_w), np.float32)
        anno = np.zeros((K, 5), np.float32)
        ind = np.zeros((K,), np.int64)
        mask = np.zeros((K,), np.uint8)
        l2i = np.repeat(base_lidar2img.astype(np.float32)[None], K, 0)
        bm = np.zeros((K, 4), np.uint8)
        src = np.full((K,), -1, np.int32)
        for k in range(min(len(order), K)):
            i = order[k]
            src[k] = i
            w = pseudo[i, 3] / vx / T(osf)
            ln = pseudo[i, 4] / vy / T(osf)
            if not (w > 0 and ln > 0):
                continue
            radius = max(int(train_cfg['min_radius']), int(gaussian_radius(ln, w, ov)))
            fx = np.float32((pseudo[i, 0] - x0) / vx / T(osf))
            fy = np.float32((pseudo[i, 1] - y0) / vy / T(osf))
            cx, cy = int(np.trunc(fx)), int(np.trunc(fy))
            if not (0 <= cx < fm_w and 0 <= cy < fm_h):
                continue
            draw_heatmap_gaussian(heat[cls[k]], cx, cy, radius)
            ind[k] = cy * fm_w + cx
            mask[k] = 1
            l2i[k] = lidar2img[i]
            bm[k] = ~bdry[i].astype(bool)
            anno[k, :4] = boxes_img[i]
            anno[k, 4] = srl[t]
        for o, v in zip(outs, (heat, anno, ind, mask, l2i, src, bm)):
            o.append(v)
    return outs
