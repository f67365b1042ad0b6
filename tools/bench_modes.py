#!/usr/bin/env python
"""The three membership entry points on config 2 (8 frames x 120 000 points x 256 boxes):
bit-packed rows (native), mmcv-layout int32 [B, M, T] (`points_in_boxes_all`) and int32 [B, M]
(`points_in_boxes_part`).  CUDA-graph replay, CUDA events; output bytes vs the HBM peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gga_b200 as G  # noqa: E402
from gga_b200 import synth  # noqa: E402


def main():
    cfg, F = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 8
    c = synth.CONFIGS[cfg]
    N, M = c['N'], c['M']
    W = G.row_words(M)
    L = G._lib.load()
    peak = 6537.6
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    pool = 3
    sets = []
    for k in range(pool):
        bt = synth.make_batch(cfg, k * F, F)
        sets.append((torch.from_numpy(bt['points']).cuda(), torch.from_numpy(bt['boxes']).cuda()))
    ws = [torch.zeros((int(L.gga_pib_workspace_bytes(F, N, M)),), dtype=torch.uint8, device='cuda') for _ in range(pool)]
    for name, fn, shape, out_bytes in (
            ('bits', L.gga_points_in_boxes_bits, (F, N, W), 4 * N * W),
            ('all', L.gga_points_in_boxes_all, (F, N, M), 4 * N * M),
            ('part', L.gga_points_in_boxes_part, (F, N), 4 * N)):
        outs = [torch.empty(shape, dtype=torch.int32, device='cuda') for _ in range(pool)]

        def call(k):
            p, b = sets[k % pool]
            rc = fn(p.data_ptr(), 4, b.data_ptr(), outs[k % pool].data_ptr(), F, N, M, ws[k % pool].data_ptr(),
                    ws[k % pool].numel(), torch.cuda.current_stream().cuda_stream)
            assert rc == 0
        for k in range(pool):
            call(k)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for k in range(2 * pool):
                call(k)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (10 * 2 * pool)
        nbytes = F * (16 * N + 28 * M + out_bytes)
        print(json.dumps(dict(cfg=cfg, mode=name, ms=round(ms, 4), algorithmic_mb=round(nbytes / 1e6, 1),
                              gbs=round(nbytes / ms / 1e6, 1), frac_of_hbm_peak=round(nbytes / ms / 1e6 / peak, 3))), flush=True)


if __name__ == '__main__':
    main()
