"""bench.py contract checks that need no GPU: the CPU reference arm prints ONE JSON line with the
keys the driver reads, and our arm refuses to run without a CUDA device (no fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True, text=True,
                          timeout=600, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS='4'))


def test_reference_arm_line():
    r = run('--impl', 'reference', '--steps', '1', '--warmup', '1')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 1 and d['steps'] == 1 and d['warmup'] >= 1
    assert d['metric'] == 'geometry_loss_fwd_bwd_frames_per_s' and d['unit'] == 'frames/s'
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['dtype'] == 'f32' and d['data'] == 'synthetic' and 'workload' in d['config']
    assert d['value'] > 0 and d['ms_per_step'] > 0
    cb = d['cpu_baseline']
    assert cb['kind'] in ('port', 'reference') and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    e = d['e2e']
    assert e['value'] == d['value'] and e['unit'] == d['unit']
    assert e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0


def test_our_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    r = run('--steps', '1', '--warmup', '1')
    assert r.returncode != 0 and 'CUDA' in (r.stderr + r.stdout)
