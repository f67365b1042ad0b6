#!/usr/bin/env python
"""End-to-end (host buffers) timing of GeometryStep.run_host on config 2 for several stream
counts of the per-frame pipeline, beside the raw pinned-copy time of the same bytes over the link
(the bound of this path: 30.7 MB of masks go back to the host every step)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gga_b200 import synth  # noqa: E402
from gga_b200.step import GeometryStep  # noqa: E402


def main():
    c = synth.CONFIGS[2]
    F, N, M = c['frames_per_gpu'], c['N'], c['M']
    hb = synth.make_batch(2, 0, F)
    hin = {k: torch.from_numpy(np.ascontiguousarray(hb[k])).pin_memory() for k in ('points', 'boxes', 'lidar2img', 'target', 'weight')}
    args = (hin['points'], hin['boxes'], hin['lidar2img'], hin['target'], hin['weight'], float(F * M))
    # raw link: 30.7 MB D2H and 15.4 MB H2D, alone and together
    d = torch.empty(30_720_000 // 4, dtype=torch.float32, device='cuda')
    h = torch.empty_like(d, device='cpu').pin_memory()
    d2 = torch.empty(15_360_000 // 4, dtype=torch.float32, device='cuda')
    h2 = torch.empty_like(d2, device='cpu').pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def t(fn, reps=20):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    def both():
        with torch.cuda.stream(s1):
            h.copy_(d, non_blocking=True)
        with torch.cuda.stream(s2):
            d2.copy_(h2, non_blocking=True)
    print(json.dumps(dict(d2h_ms=round(t(lambda: h.copy_(d, non_blocking=True)), 4),
                          h2d_ms=round(t(lambda: d2.copy_(h2, non_blocking=True)), 4), duplex_ms=round(t(both), 4))), flush=True)
    ref = None
    for pieces in (1,):
        for ns in (2, 3, 4, 6):
            s = GeometryStep(F, N, M, 'cuda', kind='giou', mode='lidar_direct')
            out = s.run_host(*args, n_streams=ns)
            if ref is None:
                ref = (out[0].clone(), out[1], out[2].clone())
            else:
                assert torch.equal(out[0], ref[0]) and out[1] == ref[1] and torch.equal(out[2], ref[2])
            ms = t(lambda: s.run_host(*args, n_streams=ns), 30)
            print(json.dumps(dict(streams=ns, ms=round(ms, 4), frames_per_s=round(F / ms * 1e3, 1))), flush=True)
            s.close()

    # pipelined: `depth` contexts used in turn, batch k+1 .. k+depth-1 submitted before batch k is waited for
    for mode in (True, False, 'hits'):
        for depth in (2, 3, 4):
            for ns in (1, 2):
                ss = [GeometryStep(F, N, M, 'cuda', kind='giou', mode='lidar_direct') for _ in range(depth)]
                for s in ss:
                    s.run_host(*args, n_streams=ns, masks_to_host=mode)

                def loop(reps=30):
                    for k in range(reps):
                        if k >= depth:
                            ss[k % depth].wait_host()
                        ss[k % depth].submit_host(*args, n_streams=ns, masks_to_host=mode)
                    for k in range(reps, reps + depth):
                        ss[k % depth].wait_host()
                loop(6)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                loop(30)
                ms = (time.perf_counter() - t0) / 30 * 1e3
                print(json.dumps(dict(masks=str(mode), depth=depth, streams=ns, ms=round(ms, 4),
                                      frames_per_s=round(F / ms * 1e3, 1))), flush=True)
                for s in ss:
                    s.close()


if __name__ == '__main__':
    main()
