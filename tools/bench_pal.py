#!/usr/bin/env python
"""Secondary bench: Point-to-Box Alignment distances fwd + Jacobian.
GGA KITTI training shape: 8 frames x 500 object slots (centerpoint_head_gga.py max_objs), a few
real objects per frame with up to 6000 in-box points (kitti_converter_gga.py:410-413).
Prints one JSON line: GPU kernel time (CUDA events, graph replay), algorithmic bytes vs the
measured HBM peak, and the CPU oracle (the reference's per-object torch loop) on the same data."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gga_b200 as G  # noqa: E402
from oracle import losses as ol  # noqa: E402


def main():
    rng = np.random.default_rng(7)
    B, K, real = 8, 500, 40
    counts = np.zeros((B, K), np.int64)
    for b in range(B):
        counts[b, :real] = rng.choice([20, 80, 300, 1200, 6000], real, p=[0.3, 0.3, 0.25, 0.1, 0.05])
    n_obj = B * K
    off = np.concatenate([[0], np.cumsum(counts.reshape(-1))]).astype(np.int64)
    bev = np.stack([rng.uniform(0, 70, n_obj), rng.uniform(-40, 40, n_obj), rng.uniform(0.5, 5, n_obj),
                    rng.uniform(0.5, 2.5, n_obj), rng.uniform(-3.14, 3.14, n_obj)], 1).astype(np.float32)
    xy = np.concatenate([bev[i, :2] + rng.normal(0, 1.5, (off[i + 1] - off[i], 2)) for i in range(n_obj)], 0).astype(np.float32)
    P = xy.shape[0]
    dxy, doff, dbev = torch.from_numpy(xy).cuda(), torch.from_numpy(off.astype(np.int32)).cuda(), torch.from_numpy(bev).cuda()
    L = G._lib.load()
    dist = torch.empty((n_obj, 3), device='cuda')
    jac = torch.empty((n_obj, 3, 5), device='cuda')
    mx = int(counts.max())
    ws = torch.empty((int(L.gga_pal_workspace_bytes(n_obj, mx)),), dtype=torch.uint8, device='cuda')

    def call():
        assert L.gga_point_box_alignment(dxy.data_ptr(), doff.data_ptr(), dbev.data_ptr(), n_obj, mx, dist.data_ptr(),
                                         jac.data_ptr(), ws.data_ptr(), ws.numel(),
                                         torch.cuda.current_stream().cuda_stream) == 0
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            call()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    nbytes = 8 * P + 4 * (n_obj + 1) + 20 * n_obj + 12 * n_obj + 60 * n_obj
    # CPU: the reference's loop (oracle restatement, torch CPU) forward + backward on the same data
    tb = torch.from_numpy(bev).clone().requires_grad_(True)
    t0 = time.perf_counter()
    rmin, rx, ry = ol.point_box_distances(torch.from_numpy(xy), off, tb)
    (rmin.sum() + rx.sum() + ry.sum()).backward()
    cpu_s = time.perf_counter() - t0
    ok = np.allclose(dist.cpu().numpy(), torch.stack([rmin, rx, ry], 1).detach().numpy(), rtol=2e-5, atol=1e-4)
    peak = 6537.6
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    print(json.dumps({'op': 'point_box_alignment fwd+jacobian', 'frames': B, 'objects': n_obj, 'points': int(P),
                      'gpu_ms': round(ms, 5), 'frames_per_s': round(B / ms * 1e3, 1),
                      'achieved_gbs': round(nbytes / ms / 1e6, 1), 'frac_of_hbm_peak': round(nbytes / ms / 1e6 / peak, 4),
                      'cpu_oracle_s': round(cpu_s, 3), 'cpu_frames_per_s': round(B / cpu_s, 2),
                      'cpu_cores': os.cpu_count(), 'parity_ok': bool(ok)}))


if __name__ == '__main__':
    main()
