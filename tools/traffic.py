"""Steady-state DRAM traffic of the membership kernel (profiles/traffic.json, read by bench.py's
`roofline.traffic`).  Run under ncu WITHOUT cache flushing so that the write-back of earlier launches'
rows is part of what a launch sees, over several rotations of buffer sets larger than 2x L2:

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none \
      -k regex:pib_sweep -s 28 -c 56 --csv --log-file gpurun_out/traffic.csv python tools/traffic.py --workload c2
  python tools/traffic.py --digest gpurun_out/traffic.csv --workload c2 > profiles/traffic.json
"""
import argparse
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def digest(path, workload):
    from gga_b200 import synth
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if 'Metric Name' in r)
    h = rows[hdr]
    im, iv, iu, ii = h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit'), h.index('ID')
    per = {}
    for r in rows[hdr + 1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(',', ''))
        u = r[iu].lower()
        v *= {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
        per.setdefault(r[ii], {})[r[im]] = v
    n = len(per)
    rd = sum(p.get('dram__bytes_read.sum', 0) for p in per.values()) / n
    wr = sum(p.get('dram__bytes_write.sum', 0) for p in per.values()) / n
    c = synth.CONFIGS[int(workload[1])]
    F, N, M = c['frames_per_gpu'], c['N'], c['M']
    W = 1 if M <= 32 else 2 if M <= 64 else 4 if M <= 128 else 8 * ((M + 255) // 256)
    alg = F * (16 * N + 28 * M + 4 * N * W)
    print(json.dumps({
        'workload': workload, 'membership_dram_bytes_per_launch': int(rd + wr), 'dram_read_bytes_per_launch': int(rd),
        'dram_write_bytes_per_launch': int(wr), 'algorithmic_bytes_per_launch': alg, 'ratio': round((rd + wr) / alg, 3),
        'launches_averaged': n,
        'how': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none over consecutive launches '
               'rotating through buffer sets > 2x L2 (steady state: the write-back of earlier launches is included)'}, indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='c2')
    ap.add_argument('--digest', default=None)
    a = ap.parse_args()
    if a.digest:
        return digest(a.digest, a.workload)
    import torch
    import gga_b200 as G
    from gga_b200 import synth
    cfg = int(a.workload[1])
    c = synth.CONFIGS[cfg]
    F, N, M = c['frames_per_gpu'], c['N'], c['M']
    W = G.row_words(M)
    L = G._lib.load()
    per_set = F * (16 * N + 28 * M + 4 * N * W)
    n_sets = max(3, int(2.2 * 126e6 / per_set) + 1)
    hb = synth.make_batch(cfg, 0, F)
    sets = [(torch.from_numpy(hb['points']).cuda(), torch.from_numpy(hb['boxes']).cuda(),
             torch.empty((F, N, W), dtype=torch.int32, device='cuda')) for _ in range(n_sets)]
    st = torch.cuda.current_stream().cuda_stream
    for rot in range(8):
        for p, b, o in sets:
            assert L.gga_points_in_boxes_bits(p.data_ptr(), 4, b.data_ptr(), o.data_ptr(), F, N, M, st) == 0
    torch.cuda.synchronize()
    print('launched', 8 * n_sets, 'membership kernels over', n_sets, 'buffer sets')


if __name__ == '__main__':
    main()
