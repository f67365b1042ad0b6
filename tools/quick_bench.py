"""Developer timing sweep of the membership kernel (not the judged bench): CUDA-event time of
gga_points_in_boxes_bits for a config across cull-grid resolutions / CTAs per frame."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gga_b200 as G  # noqa: E402
from gga_b200 import synth  # noqa: E402


def time_ms(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', type=int, default=2)
    ap.add_argument('--frames', type=int, default=8)
    ap.add_argument('--pool', type=int, default=6)
    ap.add_argument('--unsorted', action='store_true')
    ap.add_argument('--grids', default='0,8,16,24,32,40,48,64')
    ap.add_argument('--ctas', default='0')
    ap.add_argument('--phase', type=int, default=0)
    ap.add_argument('--N', type=int, default=0)
    ap.add_argument('--M', type=int, default=0)
    a = ap.parse_args()
    c = synth.CONFIGS[a.cfg]
    pool = []
    for k in range(a.pool):
        bt = synth.make_batch(a.cfg, k * a.frames, a.frames, sort_azimuth=not a.unsorted, N=a.N or None, M=a.M or None)
        pool.append((torch.from_numpy(bt['points']).cuda(), torch.from_numpy(bt['boxes']).cuda()))
    N, M = a.N or c['N'], a.M or c['M']
    W = G.row_words(M)
    outs = [torch.empty((a.frames, N, W), dtype=torch.int32, device='cuda') for _ in range(a.pool)]
    bytes_step = a.frames * (16 * N + 28 * M + 4 * N * W)
    L = G._lib.load()
    st = torch.cuda.current_stream().cuda_stream
    res = []
    for g in [int(x) for x in a.grids.split(',')]:
        for ct in [int(x) for x in a.ctas.split(',')]:
            G.ops.set_tuning(g, ct)
            wss = [torch.zeros((int(L.gga_pib_workspace_bytes(a.frames, N, M)),), dtype=torch.uint8, device='cuda')
                   for _ in range(a.pool)]

            def call(k):
                p, b = pool[k % a.pool]
                o, ws = outs[k % a.pool], wss[k % a.pool]
                rc = L.gga_points_in_boxes_bits(p.data_ptr(), 4, b.data_ptr(), o.data_ptr(), a.frames, N, M,
                                                ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
                assert rc == 0
            for k in range(a.pool):
                call(k)
            torch.cuda.synchronize()
            L.gga_test_pib_phase(a.phase)
            reps = 4 * a.pool
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):     # GPU-bound timing: the launches are replayed by the driver
                for k in range(reps):
                    call(k)
            ms = time_ms(g_.replay, iters=10, warm=2) / reps
            L.gga_test_pib_phase(0)
            r = dict(cfg=a.cfg, N=N, M=M, phase=a.phase, grid=g, ctas=ct, ms=round(ms, 4), gbs=round(bytes_step / ms / 1e6, 1),
                     frames_per_s=round(a.frames / ms * 1e3, 1))
            print(json.dumps(r), flush=True)
            res.append(r)
    G.ops.set_tuning(0, 0)
    # reference point: plain device copy bandwidth of the same byte volume
    x = torch.empty(bytes_step // 8, dtype=torch.float32, device='cuda')
    y = torch.empty_like(x)
    ms = time_ms(lambda: y.copy_(x))
    print(json.dumps(dict(copy_ms=round(ms, 4), copy_gbs=round(bytes_step / ms / 1e6, 1))))


if __name__ == '__main__':
    main()
