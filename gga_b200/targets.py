"""GGA training-target packing, same call shape as ``CenterHead_GGA.get_targets``
(``/root/reference/mmdet3d/models/dense_heads/centerpoint_head_gga.py:343-399``; per frame
``get_targets_single`` ``:401-627``).

The reference loops in Python over frames x tasks x objects, a few dozen 0-dim tensor ops per
object (``gaussian_radius`` / ``draw_heatmap_gaussian`` of ``core/utils/gaussian.py:6-86``).
Here the per-frame lists are concatenated (CSR by frame), copied to the device once, and ONE
kernel launch (``gga_pack_targets``) fills every output of every task.

The Semantic-Ratio sample ``srl`` of each (frame, task) is drawn by the reference with
``torch.normal`` inside the loop (``:515-527``); :func:`semantic_ratio_samples` draws the same
numbers in the same order from torch's global generator, so a seeded run reproduces the
reference's ``anno_box[..., 4]`` bit for bit.
"""
import torch

from . import _lib

# (mean, std) of the Semantic Ratio prior by task index: 0 Pedestrian, 1 Cyclist, others Car
# (centerpoint_head_gga.py:516-527)
SRL_PRIORS = ((1.35, 0.48), (3.60, 0.68), (2.40, 0.28))


def semantic_ratio_samples(num_frames, n_tasks):
    """[F, n_tasks] float32 on the CPU, drawn exactly like the reference's loop (frame major,
    task minor; ``torch.clamp(torch.normal(mean, std), min=1e-3)``)."""
    out = torch.empty((num_frames, n_tasks), dtype=torch.float32)
    for f in range(num_frames):
        for t in range(n_tasks):
            m, s = SRL_PRIORS[min(t, 2)]
            out[f, t] = torch.clamp(torch.normal(torch.tensor(m), torch.tensor(s)), min=1e-3)
    return out


class PackedTargets:
    """Raw task-major outputs of one ``gga_pack_targets`` launch (see include/gga_b200.h)."""

    def __init__(self, heatmap, anno_box, ind, mask, anno_lidar2img, boundary_mask, src_index, task_channels,
                 frame_offsets):
        self.heatmap, self.anno_box, self.ind, self.mask = heatmap, anno_box, ind, mask
        self.anno_lidar2img, self.boundary_mask, self.src_index = anno_lidar2img, boundary_mask, src_index
        self.task_channels, self.frame_offsets = task_channels, frame_offsets


def pack_targets(labels, frame_offsets, boxes_img, lidar2img, pseudo, bdry, base_lidar2img, srl, class_names,
                 train_cfg, device=None):
    """Concatenated inputs (see ``gga_target_args``) -> :class:`PackedTargets` on `device`.

    labels int [n]; frame_offsets int [F+1]; boxes_img [n,4]; lidar2img [n,4,4]; pseudo [n,7]
    fp32 or fp64 (the dtype the radius / centre arithmetic runs in); bdry bool [n,4];
    base_lidar2img [F,4,4]; srl [F,n_tasks]; class_names: list of per-task name lists
    (``self.class_names``); train_cfg: the head's ``train_cfg`` dict."""
    L = _lib.load()
    dev = torch.device(device if device is not None else 'cuda')
    assert dev.type == 'cuda', 'pack_targets runs on a CUDA device (there is no CPU path)'
    n_tasks = len(class_names)
    counts = [len(c) for c in class_names]
    n_classes = sum(counts)
    class_task = [t for t, c in enumerate(counts) for _ in range(c)]
    class_cls = [i for c in counts for i in range(c)]
    chan0 = [sum(counts[:t]) for t in range(n_tasks)]
    fo = torch.as_tensor(frame_offsets, dtype=torch.int32).cpu()
    F = fo.numel() - 1
    n = int(fo[-1]) if F >= 0 and fo.numel() else 0
    assert labels.shape[0] == n and boxes_img.shape[0] == n and pseudo.shape[0] == n
    max_frame = int((fo[1:] - fo[:-1]).max()) if F > 0 else 0
    osf = int(train_cfg['out_size_factor'])
    fm_w = int(train_cfg['grid_size'][0]) // osf
    fm_h = int(train_cfg['grid_size'][1]) // osf
    K = int(train_cfg['max_objs']) * int(train_cfg['dense_reg'])
    pdt = pseudo.dtype if pseudo.dtype in (torch.float32, torch.float64) else torch.float32

    def dv(x, dtype):
        x = torch.as_tensor(x).to(device=dev, dtype=dtype).contiguous()
        return x if x.numel() else torch.zeros((16,), dtype=dtype, device=dev)   # the C ABI takes no null pointers

    t_lab = dv(labels, torch.int32)
    t_fo = fo.to(dev)
    t_img, t_l2i, t_ps = dv(boxes_img, torch.float32), dv(lidar2img, torch.float32), dv(pseudo, pdt)
    t_bd, t_base, t_srl = dv(bdry, torch.uint8), dv(base_lidar2img, torch.float32), dv(srl, torch.float32)
    assert n == 0 or (t_img.numel() == 4 * n and t_l2i.numel() == 16 * n and t_ps.numel() == 7 * n and t_bd.numel() == 4 * n)
    assert F == 0 or (t_base.numel() == 16 * F and t_srl.numel() == F * n_tasks)
    t_ct, t_cc, t_c0 = dv(class_task, torch.int32), dv(class_cls, torch.int32), dv(chan0, torch.int32)

    heatmap = torch.empty((F, n_classes, fm_h, fm_w), dtype=torch.float32, device=dev)
    anno_box = torch.empty((n_tasks, F, K, 5), dtype=torch.float32, device=dev)
    ind = torch.empty((n_tasks, F, K), dtype=torch.int64, device=dev)
    mask = torch.empty((n_tasks, F, K), dtype=torch.uint8, device=dev)
    anno_l2i = torch.empty((n_tasks, F, K, 4, 4), dtype=torch.float32, device=dev)
    bmask = torch.empty((n_tasks, F, K, 4), dtype=torch.uint8, device=dev)
    src = torch.empty((n_tasks, F, K), dtype=torch.int32, device=dev)

    a = _lib.TargetArgs()
    a.labels, a.frame_offsets, a.boxes_img, a.lidar2img = t_lab.data_ptr(), t_fo.data_ptr(), t_img.data_ptr(), t_l2i.data_ptr()
    a.pseudo, a.bdry, a.base_lidar2img, a.srl = t_ps.data_ptr(), t_bd.data_ptr(), t_base.data_ptr(), t_srl.data_ptr()
    a.class_task, a.class_cls, a.task_channel0 = t_ct.data_ptr(), t_cc.data_ptr(), t_c0.data_ptr()
    a.pseudo_dtype = _lib.F64 if pdt == torch.float64 else _lib.F32
    a.num_frames, a.n_tasks, a.n_classes, a.n_channels = F, n_tasks, n_classes, n_classes
    a.max_frame_objs, a.max_objs, a.fm_w, a.fm_h = max_frame, K, fm_w, fm_h
    a.out_size_factor, a.min_radius = osf, int(train_cfg['min_radius'])
    # the reference holds these config lists as fp32 torch tensors (:421-422)
    a.pc_x0, a.pc_y0 = float(train_cfg['point_cloud_range'][0]), float(train_cfg['point_cloud_range'][1])
    a.voxel_x, a.voxel_y = float(train_cfg['voxel_size'][0]), float(train_cfg['voxel_size'][1])
    a.gaussian_overlap = float(train_cfg['gaussian_overlap'])
    a.heatmap, a.anno_box, a.ind, a.mask = heatmap.data_ptr(), anno_box.data_ptr(), ind.data_ptr(), mask.data_ptr()
    a.anno_lidar2img, a.boundary_mask, a.src_index = anno_l2i.data_ptr(), bmask.data_ptr(), src.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(L.gga_pack_targets(a, torch.cuda.current_stream(dev).cuda_stream), 'pack_targets')
    return PackedTargets(heatmap, anno_box, ind, mask, anno_l2i, bmask, src, list(zip(chan0, counts)), fo)


def get_targets(gt_labels_3d, GGA_boxes_img, GGA_lidar2img, GGA_init_pseudo_labels, GGA_bdry_masks,
                GGA_in_box_points, img_metas, class_names, train_cfg, srl=None, device=None):
    """Drop-in for ``CenterHead_GGA.get_targets`` (``self.class_names`` / ``self.train_cfg`` passed
    explicitly; ``gt_bboxes_3d`` is debug-only in the reference, ``:404``, and is not an input).

    Per-frame lists in, per-task lists out: ``(heatmaps, anno_boxes, inds, masks, anno_lidar2imgs,
    ibp_points, anno_bound_masks)`` with ``heatmaps[t] [B, C_t, H, W]``, ``anno_boxes[t] [B, K, 5]``,
    ``inds[t] [B, K]`` int64, ``masks[t] [B, K]`` uint8, ``anno_lidar2imgs[t] [B, K, 4, 4]``,
    ``ibp_points[t][b]`` = the frame's in-box point clusters in task order (``:463-479``) and
    ``anno_bound_masks[t] [B, K, 4]`` uint8."""
    B = len(gt_labels_3d)
    dev = torch.device(device if device is not None else gt_labels_3d[0].device)
    n_tasks = len(class_names)
    counts = [int(x.shape[0]) for x in gt_labels_3d]
    fo = [0]
    for c in counts:
        fo.append(fo[-1] + c)

    def cat(xs, shape_tail, dtype=None):
        xs = [torch.as_tensor(x).reshape((-1,) + shape_tail) for x in xs]
        if dtype is not None:
            xs = [x.to(dtype) for x in xs]
        return torch.cat(xs, 0) if xs else torch.zeros((0,) + shape_tail)

    pseudo = cat(GGA_init_pseudo_labels, (7,))
    base = torch.stack([torch.as_tensor(m['lidar2img']).to(torch.float32) for m in img_metas]) if B else \
        torch.zeros((0, 4, 4))
    if srl is None:
        srl = semantic_ratio_samples(B, n_tasks)
    p = pack_targets(cat(gt_labels_3d, ()), fo, cat(GGA_boxes_img, (4,), torch.float32),
                     cat(GGA_lidar2img, (4, 4), torch.float32), pseudo, cat(GGA_bdry_masks, (4,), torch.uint8),
                     base, srl, class_names, train_cfg, device=dev)
    heatmaps = [p.heatmap[:, c0:c0 + c] for c0, c in p.task_channels]
    src = p.src_index.cpu()
    ibp = []
    for t in range(n_tasks):
        per_frame = []
        for b in range(B):
            row = src[t, b]
            per_frame.append([GGA_in_box_points[b][int(i) - fo[b]].to(dev) for i in row[row >= 0]])
        ibp.append(per_frame)
    return (heatmaps, list(p.anno_box), list(p.ind), list(p.mask), list(p.anno_lidar2img), ibp,
            list(p.boundary_mask))
