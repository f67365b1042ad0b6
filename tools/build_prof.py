"""Developer build of the library with -DGGA_PROFILING (per-CTA phase stamps, CTA-size and
ranges-per-frame overrides of the membership kernel): gga_b200/_C/libgga_b200_prof.so.
Select it with GGA_B200_LIB=<path>.  Never used by tests/, bench.py or the product path."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gga_b200 import build as B  # noqa: E402

if __name__ == '__main__':
    out = os.path.join(B.OUT_DIR, 'libgga_b200_prof.so')
    os.makedirs(B.OUT_DIR, exist_ok=True)
    cmd = [B.nvcc_path(), '-ccbin', '/usr/bin/g++'] + B.NVCC_FLAGS + ['-DGGA_PROFILING', '-o', out] + B._sources()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        sys.exit(1)
    print(out)
