"""KITTI result formatting and the pseudo-label annos rewrite — the step
after the matching path.

* :func:`bbox2result_kitti` — ``KittiDataset_GGA_match.bbox2result_kitti``
  (``/root/reference/mmdet3d/datasets/kitti_dataset_GGA_match.py:458-571``) over
  ``convert_valid_bboxes`` (``:685-765``).  The reference converts, projects and filters one frame
  at a time on CPU tensors and then appends one detection at a time to Python lists; here all
  frames go through ONE projection launch (``box3d_project(mode='kitti_cam')``: corners, P2
  projection, min/max, image / range validity, clamp) plus a handful of batched torch ops for the
  LiDAR -> camera box conversion (``box_3d_mode.py:117-123,162-173``), one D2H copy, and the
  per-frame dicts are cut out of the batch arrays with numpy slicing.  Same keys, dtypes and KITTI
  text lines as the reference.
* :func:`pseudo_label_matching_kitti` — ``tools/utils_pseudo_labels_gga.py:17-88``: DontCare
  removal, 2D IoU matching of the projected detections against the 2D annotations
  (``gga_match_dt_gt``: block-diagonal IoU + argmax on the GPU instead of
  ``calculate_iou_partly``), the annos rewrite and the l/w swap (``:59-78``).
"""
import copy
import os
import pickle

import numpy as np
import torch

from .matching import fix_matched_dims, match_dt_to_gt
from .project import box3d_project

USED_CLASSES = ('Pedestrian', 'Car', 'Cyclist')   # drop_arrays_by_name default, utils_pseudo_labels_gga.py:11


def _limit_period(val, offset, period):
    return val - torch.floor(val / period + offset) * period


def _box_tensor(b):
    return b.tensor if hasattr(b, 'tensor') else torch.as_tensor(b)


def _empty_anno():
    # kitti_dataset_GGA_match.py:527-538
    return {'name': np.array([]), 'truncated': np.array([]), 'occluded': np.array([]), 'alpha': np.array([]),
            'bbox': np.zeros([0, 4]), 'dimensions': np.zeros([0, 3]), 'location': np.zeros([0, 3]),
            'rotation_y': np.array([]), 'score': np.array([])}


REFERENCE_DUMP_PATH = './data/kitti_pesudo/kitti_infos_trainval_GGA_pseudo.pkl'   # utils_pseudo_labels_gga.py:69


class KittiResultFormatter:
    """Carries what the reference method reads from ``self`` (``self.data_infos``,
    ``self.pcd_limit_range``) so that :meth:`bbox2result_kitti` has the reference's exact signature
    (``kitti_dataset_GGA_match.py:458-462``); a dataset object can equally be passed to
    :func:`bbox2result_kitti` through ``dataset=``."""

    def __init__(self, data_infos, pcd_limit_range, device=None):
        self.data_infos, self.pcd_limit_range, self.device = data_infos, pcd_limit_range, device

    def bbox2result_kitti(self, net_outputs, class_names, pklfile_prefix=None, submission_prefix=None):
        return bbox2result_kitti(net_outputs, class_names, pklfile_prefix, submission_prefix, dataset=self,
                                 device=self.device)


def bbox2result_kitti(net_outputs, class_names, pklfile_prefix=None, submission_prefix=None, *, dataset=None,
                      data_infos=None, pcd_limit_range=None, device=None):
    """The reference method's arguments in its order (``net_outputs, class_names, pklfile_prefix,
    submission_prefix``); what it reads from ``self`` comes keyword-only: ``dataset`` (any object with
    ``data_infos`` and ``pcd_limit_range``, e.g. the reference's dataset) or the two values themselves.
    ``net_outputs[i]`` = dict(boxes_3d = LiDAR boxes (an object with ``.tensor`` or a [n, 7] tensor,
    bottom centre), scores_3d [n], labels_3d [n]).  Returns ``list[dict]`` in KITTI format; writes
    ``{submission_prefix}/{idx:06d}.txt`` and ``{pklfile_prefix}.pkl`` when asked."""
    if dataset is not None:
        data_infos, pcd_limit_range = dataset.data_infos, dataset.pcd_limit_range
    assert data_infos is not None and pcd_limit_range is not None, 'pass dataset= or data_infos= and pcd_limit_range='

    assert len(net_outputs) == len(data_infos), 'invalid list length of network outputs'
    dev = torch.device(device if device is not None else 'cuda')
    assert dev.type == 'cuda', 'bbox2result_kitti runs the projection on a CUDA device (there is no CPU path)'
    if submission_prefix is not None:
        os.makedirs(submission_prefix, exist_ok=True)
    F = len(net_outputs)
    boxes = [_box_tensor(o['boxes_3d']).detach().reshape(-1, 7).float().cpu() for o in net_outputs]
    counts = [int(b.shape[0]) for b in boxes]
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    n = int(off[-1])
    det_annos = []
    if n:
        lidar = torch.cat(boxes, 0).to(dev)
        lidar[:, 6] = _limit_period(lidar[:, 6], 0.5, np.pi * 2)      # limit_yaw, :713
        fob = torch.from_numpy(np.repeat(np.arange(F, dtype=np.int32), counts)).to(dev)
        # rect @ Trv2c in float32 numpy, the reference's own expression (:724-730)
        rt = torch.from_numpy(np.stack([i['calib']['R0_rect'].astype(np.float32) @
                                        i['calib']['Tr_velo_to_cam'].astype(np.float32) for i in data_infos])).to(dev)
        P2 = torch.from_numpy(np.stack([i['calib']['P2'].astype(np.float32) for i in data_infos])).to(dev)
        hw = torch.tensor([list(i['image']['image_shape'][:2]) for i in data_infos], dtype=torch.float32, device=dev)
        bbox, valid = box3d_project(lidar, P2, mode='kitti_cam', rt=rt, img_hw=hw, pcd_range=pcd_limit_range,
                                    clamp=True, frame_of_box=fob)
        # Box3DMode.convert(LIDAR -> CAM, rt): xyz through rt, (x, y, z) sizes -> (x, z, y), yaw -> -yaw - pi/2
        xyz1 = torch.cat([lidar[:, :3], lidar.new_ones((n, 1))], 1)
        xyz = torch.bmm(xyz1[:, None, :], rt[fob.long()].transpose(1, 2))[:, 0, :3]
        yaw = _limit_period(-lidar[:, 6:7] - np.pi / 2, 0.5, np.pi * 2)
        cam = torch.cat([xyz, lidar[:, 3:4], lidar[:, 5:6], lidar[:, 4:5], yaw], 1)
        packed = torch.cat([bbox, cam, lidar[:, :2], valid.float()[:, None]], 1).cpu().numpy()   # one D2H
        bbox_h, cam_h, lidar_xy, valid_h = packed[:, :4], packed[:, 4:11], packed[:, 11:13], packed[:, 13] > 0
    for idx in range(F):
        info = data_infos[idx]
        sample_idx = info['image']['image_idx']
        anno = _empty_anno()
        if counts[idx]:
            sl = slice(off[idx], off[idx + 1])
            keep = valid_h[sl]
            if keep.any():
                cam_f, xy = cam_h[sl][keep], lidar_xy[sl][keep]
                scores = torch.as_tensor(net_outputs[idx]['scores_3d']).detach().cpu().numpy()[keep]
                labels = torch.as_tensor(net_outputs[idx]['labels_3d']).detach().cpu().numpy()[keep]
                k = int(keep.sum())
                anno = {
                    'name': np.array([class_names[int(lb)] for lb in labels]),
                    'truncated': np.zeros(k, np.float64),
                    'occluded': np.zeros(k, np.int64),
                    'alpha': -np.arctan2(-xy[:, 1], xy[:, 0]) + cam_f[:, 6],     # :516-517
                    'bbox': bbox_h[sl][keep].copy(),
                    'dimensions': cam_f[:, 3:6].copy(),
                    'location': cam_f[:, :3].copy(),
                    'rotation_y': cam_f[:, 6].copy(),
                    'score': scores,
                }
        if submission_prefix is not None:
            with open(f'{submission_prefix}/{sample_idx:06d}.txt', 'w') as f:
                f.write(kitti_lines(anno))
        anno['sample_idx'] = np.array([sample_idx] * len(anno['score']), dtype=np.int64)
        det_annos.append(anno)
    if pklfile_prefix is not None:
        out = pklfile_prefix if pklfile_prefix.endswith(('.pkl', '.pickle')) else f'{pklfile_prefix}.pkl'
        with open(out, 'wb') as f:
            pickle.dump(det_annos, f)
    return det_annos


def kitti_lines(anno):
    """The KITTI submission text of one frame (``:541-558``; dims printed as h w l)."""
    bbox, loc, dims = anno['bbox'], anno['location'], anno['dimensions']
    lines = []
    for i in range(len(bbox)):
        lines.append('{} -1 -1 {:.4f} {:.4f} {:.4f} {:.4f} {:.4f} {:.4f} {:.4f} {:.4f} {:.4f} {:.4f} {:.4f} {:.4f} {:.4f}\n'.format(
            anno['name'][i], anno['alpha'][i], bbox[i][0], bbox[i][1], bbox[i][2], bbox[i][3], dims[i][1], dims[i][2],
            dims[i][0], loc[i][0], loc[i][1], loc[i][2], anno['rotation_y'][i], anno['score'][i]))
    return ''.join(lines)


def _strip_dontcare(anno, keys=None):
    """``utils_pseudo_labels_gga.py:28-37``: keep the leading non-DontCare objects, then only the
    used classes."""
    num_obj = len([n for n in anno['name'] if n != 'DontCare'])
    for key in list(anno.keys()):
        if keys is None or key in keys:
            anno[key] = anno[key][:num_obj]
    select = np.array([i for i, x in enumerate(anno['name']) if x in USED_CLASSES], dtype=np.int64)
    for key in list(anno.keys()):
        if keys is None or key in keys:
            anno[key] = anno[key][select]


def pseudo_label_matching_kitti(gt_infos, dt_annos, metric=0, num_parts=200, *, out_path=REFERENCE_DUMP_PATH,
                                device=None, return_infos=False):
    """``pseudo_label_matching_kitti(gt_infos, dt_annos, metric=0, num_parts=200) -> gt_annos``
    (``tools/utils_pseudo_labels_gga.py:17-84``), same positional signature and return value.

    Modifies ``gt_infos[i]['annos']`` in place exactly like the reference (pops
    ``GGA_in_box_points``, removes DontCare / unused classes), rewrites the annotations of a deep
    copy of the infos from the matched detections and dumps that copy to ``out_path`` (default: the
    reference's fixed file, its directory created if needed; ``None`` = no file).  Returns the
    cleaned ``gt_annos`` like the reference; ``return_infos=True`` returns ``(gt_annos, new_infos)``.
    ``metric`` must be 0 (2D image boxes, the only one the GGA tool uses, ``:45``); ``num_parts`` only
    chunks the reference's IoU computation and has no effect on the result."""
    if metric != 0:
        raise NotImplementedError('pseudo-label matching uses the 2D image-box IoU (metric=0)')
    gt_annos = [info['annos'] for info in gt_infos]
    assert len(gt_annos) == len(dt_annos)
    new_infos = copy.deepcopy(gt_infos)
    for a in gt_annos:
        a.pop('GGA_in_box_points')
        _strip_dontcare(a)
    dev = torch.device(device if device is not None else 'cuda')
    F = len(gt_annos)
    dt_counts = [len(a['name']) for a in dt_annos]
    gt_counts = [len(a['name']) for a in gt_annos]
    for f in range(F):
        assert dt_counts[f] == 0 or gt_counts[f] > 0, f'frame {f}: detections but no annotation to match'
    do = torch.from_numpy(np.concatenate([[0], np.cumsum(dt_counts)]).astype(np.int32))
    go = torch.from_numpy(np.concatenate([[0], np.cumsum(gt_counts)]).astype(np.int32))
    match_h = np.zeros((0,), np.int64)
    if int(do[-1]):
        dt_boxes = torch.from_numpy(np.concatenate([np.asarray(a['bbox'], np.float32).reshape(-1, 4) for a in dt_annos], 0))
        gt_boxes = torch.from_numpy(np.concatenate([np.asarray(a['bbox'], np.float64).reshape(-1, 4) for a in gt_annos], 0))
        match, _ = match_dt_to_gt(dt_boxes.to(dev), do, gt_boxes, go)
        match_h = match.cpu().numpy().astype(np.int64)
    new_annos = []
    for f in range(F):
        gt, dt = gt_annos[f], dt_annos[f]
        if dt_counts[f] == 0:
            new_annos.append({k: v[:0] for k, v in gt.items()})
            continue
        m = match_h[int(do[f]):int(do[f + 1])]
        new = {k: (dt[k] if k in dt else v[m]) for k, v in gt.items()}
        new['dimensions'], new['rotation_y'] = fix_matched_dims(new['dimensions'], new['rotation_y'])   # :70-78
        new_annos.append(new)
    for f, sample in enumerate(new_infos):
        sample.pop('annos')
        sample['annos'] = new_annos[f]
    if out_path is not None:
        d = os.path.dirname(out_path)
        if d:
            os.makedirs(d, exist_ok=True)
        with open(out_path, 'wb') as fh:
            pickle.dump(new_infos, fh)
    return (gt_annos, new_infos) if return_infos else gt_annos
