#!/usr/bin/env python
"""Top source lines of each profiled kernel by warp-stall samples / executed instructions.
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [topN] [kernel-substring]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25; filt = sys.argv[3] if len(sys.argv) > 3 else ''
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source=cuda,sass'],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
def num(v):
    try:
        return int(float(v))
    except Exception:
        return 0
seen = set()
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == 'Function Name':
        name = r[1]; hdr = rows[i + 1]; i += 2
        lines = []
        while i < len(rows) and rows[i] and rows[i][0] not in ('File Path', 'Function Name'):
            if rows[i][0] != '':
                lines.append(rows[i])
            i += 1
        if name in seen or filt not in name:
            continue
        seen.add(name)
        si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
        stall_cols = [k for k, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        tot_s = sum(num(l[si]) for l in lines) or 1; tot_i = sum(num(l[ii]) for l in lines) or 1
        print(f'## {name[:100]}  samples={tot_s} warp-instr={tot_i}')
        for l in sorted(lines, key=lambda l: -num(l[(ii if "--by-ins" in sys.argv else si)]))[:top]:
            st = sorted(((num(l[k]), hdr[k][6:]) for k in stall_cols), reverse=True)[:3]
            sts = ' '.join(f'{n}:{v}' for v, n in st if v)
            print(f'{l[0]:>5s} smp {100*num(l[si])/tot_s:5.1f}% ins {100*num(l[ii])/tot_i:5.1f}% | {l[1].strip()[:95]:95s} | {sts}')
    else:
        i += 1
