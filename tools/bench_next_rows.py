#!/usr/bin/env python
"""Timings of the 'next' rows 3 and 4 of SURVEY.md §8f on the GPU box, beside their CPU
restatements (oracle = the reference's per-object Python loops restated in numpy):
target packing (get_targets) and KITTI result formatting + pseudo-label rewrite.
One JSON line per row."""
import copy
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gga_b200 import kitti_format as KF  # noqa: E402
from gga_b200 import synth  # noqa: E402
from gga_b200 import targets as T  # noqa: E402
from oracle import targets as ot  # noqa: E402


def wall(fn, reps):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def targets_row():
    rng = np.random.default_rng(5)
    counts = [120] * 8                       # 8 frames x 120 objects (copy-paste augmented KITTI batch)
    frames = [synth.make_target_frame(rng, n, 3, np.float64) for n in counts]
    srl = rng.uniform(0.5, 4, (8, 3)).astype(np.float32)
    fo = np.concatenate([[0], np.cumsum(counts)])
    cat = {k: torch.from_numpy(np.concatenate([fr[k] for fr in frames], 0)) for k in ('labels', 'boxes_img', 'lidar2img', 'pseudo', 'bdry')}
    base = np.stack([fr['base_lidar2img'] for fr in frames])

    def ours():
        return T.pack_targets(cat['labels'], fo, cat['boxes_img'], cat['lidar2img'], cat['pseudo'], cat['bdry'], base, srl,
                              synth.KITTI_TASKS, synth.KITTI_TRAIN_CFG, device='cuda')
    ms = wall(ours, 50)
    dev = {k: v.cuda() for k, v in cat.items()}
    bd, sd, fod = torch.from_numpy(base).cuda(), torch.from_numpy(srl).cuda(), torch.from_numpy(fo).int()

    def ours_dev():
        return T.pack_targets(dev['labels'], fod, dev['boxes_img'], dev['lidar2img'], dev['pseudo'], dev['bdry'], bd, sd,
                              synth.KITTI_TASKS, synth.KITTI_TRAIN_CFG, device='cuda')
    ms_dev = wall(ours_dev, 50)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ours_dev()
    e1.record()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f, fr in enumerate(frames):
        ot.get_targets_single(fr['labels'], fr['boxes_img'], fr['lidar2img'], fr['pseudo'], fr['bdry'], fr['base_lidar2img'],
                              srl[f], synth.KITTI_TASKS, synth.KITTI_TRAIN_CFG)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    return dict(row='get_targets (8 frames x 120 objects, 3 tasks, 176x200 heatmaps)', ms_host_inputs=round(ms, 3),
                ms_device_inputs=round(ms_dev, 3), ms_device_events=round(e0.elapsed_time(e1) / 20, 3),
                cpu_numpy_restatement_ms=round(cpu_ms, 1), speedup=round(cpu_ms / ms, 1))


def format_row():
    counts = [60] * 200                      # 200 frames x 60 detections
    infos, dets = synth.make_detection_frames(7, counts)

    def ours():
        return KF.bbox2result_kitti(dets, infos, ['Pedestrian', 'Cyclist', 'Car'], list(synth.KITTI_MATCH_RANGE))
    ms = wall(ours, 5)
    annos = ours()

    def rewrite():
        return KF.pseudo_label_matching_kitti(copy.deepcopy(infos), annos)
    t0 = time.perf_counter()
    rewrite()
    ms2 = (time.perf_counter() - t0) * 1e3
    return dict(row='bbox2result_kitti + pseudo_label_matching_kitti (200 frames x 60 detections)',
                format_ms=round(ms, 2), rewrite_ms_incl_deepcopy=round(ms2, 2), detections_kept=int(sum(len(a['name']) for a in annos)))


if __name__ == '__main__':
    print(json.dumps(targets_row()), flush=True)
    print(json.dumps(format_row()), flush=True)
