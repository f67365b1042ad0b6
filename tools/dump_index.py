"""Developer tool: decode the per-frame box index left in the membership workspace."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gga_b200 as G
from gga_b200 import synth
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = synth.CONFIGS[cfg]; N, M = c['N'], c['M']; F = 8 if cfg != 5 else 1
L = G._lib.load()
bt = synth.make_batch(cfg, 0, F)
p, b = torch.from_numpy(bt['points']).cuda(), torch.from_numpy(bt['boxes']).cuda()
o = torch.empty((F, N, G.row_words(M)), dtype=torch.int32, device='cuda')
ws = torch.zeros((int(L.gga_pib_workspace_bytes(F, N, M)),), dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    assert L.gga_points_in_boxes_bits(p.data_ptr(), 4, b.data_ptr(), o.data_ptr(), F, N, M, ws.data_ptr(), ws.numel(), st) == 0
torch.cuda.synchronize()
w = ws.cpu().numpy()
def al(v, a=256): return (v + a - 1) // a * a
# mirror of ws_layout (membership.cu)
Gmax = int(round(min(64.0 * M, 4.0 * N) ** 0.5)); Gmax = max(4, min(192, Gmax))
gstride = al((Gmax + 2) ** 2, 4)
cap = max(8 * Gmax * Gmax, 16 * M + 64); cap = min(cap, (1 << 20) - 2) & ~7
o_hdr = 0; off = al(F * 48)
o_prep = off; off = al(off + F * M * 32)
o_grid = off; off = al(off + F * gstride * 4)
o_ids = off; off = al(off + F * cap * 2)
assert off == len(w), (off, len(w))
for f in range(min(F, 3)):
    hdr = w[o_hdr + 48 * f: o_hdr + 48 * f + 48]
    fl = hdr.view(np.float32); it = hdr.view(np.int32)
    Gf = it[8]
    print(f'frame {f}: G={Gf} n_rect={it[9]} n_inf={it[10]} g0=({fl[0]:.2f},{fl[1]:.2f}) inv=({fl[2]:.3f},{fl[3]:.3f}) cw=({fl[4]:.3f},{fl[5]:.3f})')
    g = w[o_grid + 4 * gstride * f: o_grid + 4 * gstride * f + 4 * (Gf + 2) ** 2].view(np.uint32).reshape(Gf + 2, Gf + 2)
    kind = g >> 30
    cnt = np.where(kind < 3, kind, (g >> 20) & 0x3ff)
    print('  cells: empty %.3f single %.3f double %.3f list %.4f ALL %d ; max list %d' % (
        (kind == 0).mean(), (kind == 1).mean(), (kind == 2).mean(), (kind == 3).mean(), (g == 0xffffffff).sum(), cnt.max()))
    print('  border rows max cnt', cnt[0].max(), cnt[-1].max(), cnt[:, 0].max(), cnt[:, -1].max())
    # candidates per point
    P = bt['points'][f]
    fx = np.float32(np.float32(P[:, 0] - fl[0]) * fl[2]) + np.float32(1); fy = np.float32(np.float32(P[:, 1] - fl[1]) * fl[3]) + np.float32(1)
    cx = np.clip(fx, 0, Gf + 1).astype(int); cy = np.clip(fy, 0, Gf + 1).astype(int)
    n = cnt[cy, cx]
    print('  candidates per point: mean %.3f  frac>0 %.3f  max %d ; hist' % (n.mean(), (n > 0).mean(), n.max()), np.bincount(n)[:10])
    mx = n[: N // 32 * 32].reshape(-1, 32).max(1)
    print('  warp-max per 32-batch: mean %.2f  hist' % mx.mean(), np.bincount(mx)[:12])
