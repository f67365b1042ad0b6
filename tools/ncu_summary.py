#!/usr/bin/env python
"""Digest ncu outputs (run here, no GPU needed) into small text summaries for profiles/.

    python tools/ncu_summary.py launches gpurun_out/r1a_launches.csv  > profiles/r1a_launches.txt
    python tools/ncu_summary.py full     gpurun_out/r1a_prof_pib.ncu-rep > profiles/r1a_pib_full.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor',
    'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__sass_thread_inst_executed_op_fp32_pred_on.sum',
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hdr]
    ik, im, iv = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value')
    agg = OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= iv or r[im] != 'gpu__time_duration.sum':
            continue
        name = r[ik].split('(')[0][:90]
        v = float(r[iv].replace(',', ''))
        a = agg.setdefault(name, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    unit = rows[hdr + 1][h.index('Metric Unit')] if 'Metric Unit' in h else '?'
    tot = sum(a[1] for a in agg.values())
    print(f'# per-kernel device time from {path} (ncu --metrics gpu__time_duration.sum, cold-cache, serialised)')
    print(f'# unit: {unit}; compare SHARES, not absolutes')
    print(f'{"kernel":92s} {"n":>5s} {"total":>12s} {"avg":>10s} {"min":>10s} {"max":>10s} {"share":>7s}')
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{name:92s} {a[0]:5d} {a[1]:12.1f} {a[1] / a[0]:10.2f} {a[2]:10.2f} {a[3]:10.2f} {100 * a[1] / tot:6.1f}%')


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f'# ncu --set full summary of {path}')
    for r in rows[2:]:
        print('## ' + r[hdr.index('Kernel Name')][:120])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'{k:90s} {r[i]:>18s} {units[i]}')
        print()


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
