"""Per-warp timeline of the streaming kernel (developer tool): globaltimer stamps at kernel
entry, before/after the grid-dependency wait, after the first lookup, after every batch."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gga_b200 as G
from gga_b200 import synth
cfg, F = 2, 8
c = synth.CONFIGS[cfg]; N, M = c['N'], c['M']
L = G._lib.load()
G.ops.set_tuning(0, int(sys.argv[1]) if len(sys.argv) > 1 else 0)
sets = []
for k in range(4):
    bt = synth.make_batch(cfg, k * F, F)
    sets.append((torch.from_numpy(bt['points']).cuda(), torch.from_numpy(bt['boxes']).cuda(),
                 torch.empty((F, N, G.row_words(M)), dtype=torch.int32, device='cuda'),
                 torch.zeros((int(L.gga_pib_workspace_bytes(F, N, M)),), dtype=torch.uint8, device='cuda')))
nwarps = 148 * 8 * 8
trace = torch.zeros((nwarps, 16), dtype=torch.int64, device='cuda')
st = torch.cuda.current_stream().cuda_stream
def call(k):
    p, b, o, ws = sets[k % 4]
    assert L.gga_points_in_boxes_bits(p.data_ptr(), 4, b.data_ptr(), o.data_ptr(), F, N, M, ws.data_ptr(), ws.numel(), st) == 0
for k in range(8): call(k)
torch.cuda.synchronize()
L.gga_test_pib_trace(trace.data_ptr())
call(8)
torch.cuda.synchronize()
L.gga_test_pib_trace(None)
t = trace.cpu().numpy()
live = t[:, 0] > 0
t = t[live]
t0 = t[:, 0].min()
print('warps traced', len(t), 'SMs', len(np.unique(t[:, 15])))
def stat(name, v):
    print(f'{name:34s} min {v.min()/1e3:7.2f}  p10 {np.percentile(v,10)/1e3:7.2f}  med {np.median(v)/1e3:7.2f}  p90 {np.percentile(v,90)/1e3:7.2f}  max {v.max()/1e3:7.2f} us')
stat('kernel entry (since first warp)', t[:, 0] - t0)
stat('before dependency wait', t[:, 1] - t0)
stat('after dependency wait', t[:, 2] - t0)
stat('after smem prep + first lookup', t[:, 3] - t0)
for k in range(4, 14):
    m = t[:, k] > 0
    if m.sum():
        stat(f'end of batch {k-4} ({m.sum()} warps)', t[m, k] - t0)
        if k > 4:
            stat(f'   duration of batch {k-4}', t[m, k] - t[m, k - 1])
stat('warp exit', t[:, 14] - t0)
stat('first batch duration', t[:, 4] - t[:, 3])
