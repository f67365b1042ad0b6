#!/usr/bin/env python
"""Secondary measurements on the other BASELINE.json configs (the judged line is bench.py's c2):
c1 single KITTI frame, c3 SUN-RGBD shape, c5 roofline stress (membership + box loss), and c4 the
pseudo-label matching pass (variant-B projection + block-diagonal IoU + argmax).  One JSON line
per config: CUDA-event time of a CUDA-graph replay, algorithmic bytes, fraction of the measured
HBM peak."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gga_b200 as G  # noqa: E402
from gga_b200 import synth  # noqa: E402
from gga_b200.step import GeometryStep  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        return 6650.0


def time_graph(fn, reps=20, iters=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * iters)


def step_config(cfg, frames):
    c = synth.CONFIGS[cfg]
    N, M = c['N'], c['M']
    W = G.row_words(M)
    n_sets = max(2, int(300e6 // (frames * (16 * N + 4 * N * W))) + 1)
    n_sets = min(n_sets, 8)
    sets, steps = [], []
    for k in range(n_sets):
        bt = synth.make_batch(cfg, 50 * k, frames)
        t = {n: torch.from_numpy(np.ascontiguousarray(bt[n])).cuda() for n in ('points', 'boxes', 'lidar2img', 'target', 'weight')}
        sets.append(t)
        steps.append(GeometryStep(frames, N, M, 'cuda', kind='giou', mode='lidar_direct'))
    it = [0]

    def fn():
        k = it[0] % n_sets
        it[0] += 1
        t = sets[k]
        steps[k].run(t['points'], t['boxes'], t['lidar2img'], t['target'], t['weight'], float(frames * M))
    ms = time_graph(fn, reps=2 * n_sets)
    nbytes = frames * (16 * N + 28 * M + 4 * N * W + M * 156)
    return dict(cfg=cfg, name=c['name'], frames=frames, N=N, M=M, ms_per_step=round(ms, 5),
                frames_per_s=round(frames / ms * 1e3, 1), algorithmic_mb=round(nbytes / 1e6, 2),
                achieved_gbs=round(nbytes / ms / 1e6, 1), frac_of_hbm_peak=round(nbytes / ms / 1e6 / peak(), 4))


def matching_config(frames=464, M=512, Gt=8):
    rng = np.random.default_rng(4)
    boxes = np.concatenate([synth.make_boxes(rng, M) for _ in range(frames)], 0)
    fob = np.repeat(np.arange(frames, dtype=np.int32), M)
    rect = np.broadcast_to(synth.KITTI_RECT, (frames, 4, 4)).copy()
    trv = np.broadcast_to(synth.KITTI_TRV2C, (frames, 4, 4)).copy()
    p2 = np.broadcast_to(synth.KITTI_P2, (frames, 4, 4)).copy()
    hw = np.broadcast_to(np.float32(synth.KITTI_IMG_HW), (frames, 2)).copy()
    gt = np.concatenate([synth.make_targets(rng, boxes[f * M:(f + 1) * M], Gt) for f in range(frames)], 0).astype(np.float64)
    d = {k: torch.from_numpy(v).cuda() for k, v in dict(boxes=boxes, fob=fob, rect=rect, trv=trv, p2=p2, hw=hw, gt=gt).items()}
    do = torch.arange(0, frames * M + 1, M, dtype=torch.int32, device='cuda')
    go = torch.arange(0, frames * Gt + 1, Gt, dtype=torch.int32, device='cuda')

    def fn():
        r = G.convert_valid_bboxes_batch(d['boxes'], d['fob'], d['rect'], d['trv'], d['p2'], d['hw'], synth.KITTI_MATCH_RANGE)
        G.match_dt_to_gt(r['bbox'], do, d['gt'], go)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    return dict(cfg=4, name='pseudo_label_matching', frames=frames, M=M, gt_per_frame=Gt, ms_per_pass=round(ms, 4),
                frames_per_s=round(frames / ms * 1e3, 1), note='eager torch-level calls (projection kernel + match kernel + small torch ops); latency bound')


if __name__ == '__main__':
    for cfg, fr in ((1, 1), (3, 8), (5, 1)):
        print(json.dumps(step_config(cfg, fr)), flush=True)
    print(json.dumps(matching_config()), flush=True)
