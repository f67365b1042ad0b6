#!/bin/bash
set -u
for fw in 8 4; do for m in 1 2 3; do
  echo "== fast warps $fw range mult $m"
  GGA_PIB_FAST_WARPS=$fw GGA_PIB_RANGE_MULT=$m timeout 300 python tools/quick_bench.py --cfg 2 --frames 8 --grids 0 --ctas 0 2>/dev/null | head -1
  GGA_PIB_FAST_WARPS=$fw GGA_PIB_RANGE_MULT=$m timeout 300 python tools/quick_bench.py --cfg 3 --N 50016 --frames 8 --grids 0 --ctas 0 2>/dev/null | head -1
done; done
GGA_PIB_FAST_WARPS=4 timeout 600 python -m pytest tests/test_gpu_membership.py -m gpu -x -q 2>&1 | tail -2
