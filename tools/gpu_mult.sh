#!/bin/bash
set -u
for occ in 5 6; do for m in 2 3; do
  echo "== nosp occ $occ mult $m"
  GGA_PIB_NOSP=1 GGA_PIB_OCC=$occ GGA_PIB_MULT=$m timeout 300 python tools/quick_bench.py --cfg 2 --frames 8 --grids 0 --ctas 0 2>/dev/null | head -1
  GGA_PIB_NOSP=1 GGA_PIB_OCC=$occ GGA_PIB_MULT=$m timeout 300 python tools/quick_bench.py --cfg 3 --frames 8 --grids 0 --ctas 0 2>/dev/null | head -1
done; done
echo "== defaults"
timeout 300 python tools/quick_bench.py --cfg 2 --frames 8 --grids 0 --ctas 0 2>/dev/null | head -1
timeout 300 python tools/quick_bench.py --cfg 3 --frames 8 --grids 0 --ctas 0 2>/dev/null | head -1
GGA_PIB_OCC=6 timeout 600 python -m pytest tests/test_gpu_membership.py -m gpu -x -q 2>&1 | tail -2
GGA_PIB_OCC=6 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('occ6',d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['kernel_ms'])"
