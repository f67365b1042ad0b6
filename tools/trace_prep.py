"""Per-CTA phase timeline of the index-build kernel (developer tool)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gga_b200 as G
from gga_b200 import synth
cfg, F = 2, 8
c = synth.CONFIGS[cfg]; N, M = c['N'], c['M']
L = G._lib.load()
sets = []
for k in range(4):
    bt = synth.make_batch(cfg, k * F, F)
    sets.append((torch.from_numpy(bt['points']).cuda(), torch.from_numpy(bt['boxes']).cuda(),
                 torch.empty((F, N, G.row_words(M)), dtype=torch.int32, device='cuda'),
                 torch.zeros((int(L.gga_pib_workspace_bytes(F, N, M)),), dtype=torch.uint8, device='cuda')))
trace = torch.zeros((4096, 16), dtype=torch.int64, device='cuda')
st = torch.cuda.current_stream().cuda_stream
def call(k):
    p, b, o, ws = sets[k % 4]
    assert L.gga_points_in_boxes_bits(p.data_ptr(), 4, b.data_ptr(), o.data_ptr(), F, N, M, ws.data_ptr(), ws.numel(), st) == 0
for k in range(8): call(k)
torch.cuda.synchronize()
L.gga_test_pib_trace_prep(trace.data_ptr())
call(8)
torch.cuda.synchronize()
L.gga_test_pib_trace_prep(None)
t = trace.cpu().numpy()
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = []
print('CTAs', len(t))
for k in range(16):
    m = t[:, k] > 0
    if m.sum() == 0: continue
    v = (t[m, k] - t0) / 1e3
    d = (t[m, k] - t[m, k - 1]) / 1e3 if k else v
    print(f'stamp {k:2d}: at med {np.median(v):6.2f} max {v.max():6.2f} us | phase med {np.median(d):5.2f} max {d.max():5.2f} us  ({m.sum()} CTAs)')
