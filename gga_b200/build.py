"""Builds gga_b200/_C/libgga_b200.so from gga_b200/csrc/*.cu with nvcc for sm_100a.

In-tree on purpose: the built .so is git-ignored but travels to the GPU box with the
repo snapshot, and the round-end driver records which in-tree .so files were loaded.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.path.join(HERE, '_C')
LIB = os.path.join(OUT_DIR, 'libgga_b200.so')
STAMP = os.path.join(OUT_DIR, 'libgga_b200.stamp')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '--fmad=false',            # the contract pins every rounding; contraction is opt-in per site
    '-Xcompiler', '-fPIC,-O2,-ffp-contract=off,-Wall', '-shared', '-cudart', 'static',
    '-Xptxas', '-v',
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + \
        sorted(glob.glob(os.path.join(HERE, '..', 'include', '*.h')))
    for f in files:
        h.update(f.encode())
        with open(f, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isfile(c) or c == 'nvcc'):
            return c
    return 'nvcc'


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    dg = _digest()
    if not force and os.path.isfile(LIB) and os.path.isfile(STAMP) and open(STAMP).read() == dg:
        return LIB
    env = dict(os.environ)
    env.setdefault('NVCC_CCBIN', '/usr/bin/g++')
    cmd = [nvcc_path(), '-ccbin', '/usr/bin/g++'] + NVCC_FLAGS + ['-o', LIB] + _sources()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(OUT_DIR, 'build.log')
    with open(log, 'w') as fh:
        fh.write(' '.join(cmd) + '\n' + res.stdout)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed (exit {res.returncode}); see {log}')
    with open(STAMP, 'w') as fh:
        fh.write(dg)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
