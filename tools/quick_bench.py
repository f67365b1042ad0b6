"""Developer timing of the membership kernel alone (not the judged bench): CUDA-event time of
back-to-back gga_points_in_boxes_bits launches over rotating buffer sets (> 2x L2), captured in
one CUDA graph.  With the GGA_PROFILING build (tools/build_prof.py, GGA_B200_LIB=...) it can also
override the CTA size / ranges per frame and dump the per-CTA phase timeline."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gga_b200 as G  # noqa: E402
from gga_b200 import synth  # noqa: E402


def sm_clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    except Exception:
        return -1


SPIN_S = 0.0


def time_ms(fn, iters=20, warm=3):
    import time
    for _ in range(warm):
        fn()
    t0 = time.time()
    while time.time() - t0 < SPIN_S:      # keep the GPU busy so that the SM clock has ramped up
        for _ in range(50):
            fn()
        torch.cuda.synchronize()
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
    ap.add_argument('--frames', type=int, default=0)
    ap.add_argument('--unsorted', action='store_true')
    ap.add_argument('--mode', default='bits', choices=['bits', 'all', 'part'])
    ap.add_argument('--N', type=int, default=0)
    ap.add_argument('--M', type=int, default=0)
    ap.add_argument('--nt', default='0')
    ap.add_argument('--ranges', default='0')
    ap.add_argument('--variant', default='0')
    ap.add_argument('--trace', action='store_true')
    ap.add_argument('--spin', type=float, default=0.0)
    a = ap.parse_args()
    global SPIN_S
    SPIN_S = a.spin
    c = synth.CONFIGS[a.cfg]
    F = a.frames or c['frames_per_gpu']
    N, M = a.N or c['N'], a.M or c['M']
    W = G.row_words(M)
    L = G._lib.load()
    prof = hasattr(L, 'gga_prof_pib')
    out_bytes = {'bits': 4 * N * W, 'all': 4 * N * M, 'part': 4 * N}[a.mode]
    bytes_step = F * (16 * N + 28 * M + out_bytes)
    pool = max(3, int(2.2 * 126e6 / bytes_step) + 1)
    pool = min(pool, 24)
    host = [synth.make_batch(a.cfg, 7 * k, F, sort_azimuth=not a.unsorted, N=a.N or None, M=a.M or None) for k in range(2)]
    sets = []
    for k in range(pool):
        p = torch.from_numpy(host[k % 2]['points']).cuda()
        b = torch.from_numpy(host[k % 2]['boxes']).cuda()
        shape = {'bits': (F, N, W), 'all': (F, N, M), 'part': (F, N)}[a.mode]
        sets.append((p, b, torch.empty(shape, dtype=torch.int32, device='cuda')))
    fn = {'bits': L.gga_points_in_boxes_bits, 'all': L.gga_points_in_boxes_all, 'part': L.gga_points_in_boxes_part}[a.mode]

    def call(k):
        p, b, o = sets[k % pool]
        rc = fn(p.data_ptr(), 4, b.data_ptr(), o.data_ptr(), F, N, M, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, L.gga_last_error()

    for nt in [int(x) for x in a.nt.split(',')]:
      for rg in [int(x) for x in a.ranges.split(',')]:
        for var in [int(x) for x in a.variant.split(',')]:
            if prof:
                L.gga_prof_pib(nt, rg, var, None)
            elif nt or rg or var:
                continue
            for k in range(pool):
                call(k)
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                for k in range(pool):
                    call(k)
            ms = time_ms(g_.replay) / pool
            clk = sm_clock()
            r = dict(cfg=a.cfg, mode=a.mode, F=F, N=N, M=M, nt=nt, ranges=rg, variant=var, us=round(ms * 1e3, 2),
                     gbs=round(bytes_step / ms / 1e6, 1), frac=round(bytes_step / ms / 1e6 / 6537.6, 3), pool=pool, sm_mhz_after=clk,
                     sorted=not a.unsorted)
            print(json.dumps(r), flush=True)
            if a.trace and prof:
                nw = (nt or 1024) // 32
                tr = torch.zeros((148 * nw * 16,), dtype=torch.int64, device='cuda')
                L.gga_prof_pib(nt, rg, var, tr.data_ptr())
                g2 = torch.cuda.CUDAGraph()      # the trace pointer is a kernel parameter: capture again
                with torch.cuda.graph(g2):
                    for k in range(pool):
                        call(k)
                for _ in range(3):
                    g2.replay()                  # steady state: the stamps of the last launch of a replay survive
                torch.cuda.synchronize()
                L.gga_prof_pib(nt, rg, var, None)
                t = tr.cpu().numpy().reshape(148, nw, 16).astype(np.float64)
                t = t[t[:, 0, 0] > 0]
                t0 = t[:, :, 0].min()
                rel = (t - t0) / 1e3
                nbw = (min(M, 1024) + 31) // 32
                names = ['past wait', 'boxes requested', 'B1', 'rects published', 'cells filled', 'tables done', 'index done', 'sweep done']
                for cls, sel in (('box warps', slice(0, nbw)),):
                    sub = rel[:, sel, :]
                    if sub.size == 0:
                        continue
                    print(f'  -- {cls}')
                    for j, nm in enumerate(names):
                        col = sub[:, :, j].ravel()
                        col = col[col > -1e6]
                        print(f'  {nm:18s} min {col.min():7.2f}  median {np.median(col):7.2f}  max {col.max():7.2f} us')
    x = torch.empty(bytes_step // 8, dtype=torch.float32, device='cuda')
    y = torch.empty_like(x)
    ms = time_ms(lambda: y.copy_(x))
    print(json.dumps(dict(copy_same_bytes_us=round(ms * 1e3, 2), copy_gbs=round(bytes_step / ms / 1e6, 1))))


if __name__ == '__main__':
    main()
