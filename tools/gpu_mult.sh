#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_membership.py tests/test_gpu_step.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/quick_bench.py --cfg 2 --frames 8 --grids 0 --ctas 0 2>/dev/null | head -1
timeout 300 python tools/quick_bench.py --cfg 3 --frames 8 --grids 0 --ctas 0 2>/dev/null | head -1
timeout 300 python tools/quick_bench.py --cfg 3 --N 50016 --frames 8 --grids 0 --ctas 0 2>/dev/null | head -1
timeout 300 python tools/quick_bench.py --cfg 1 --frames 1 --grids 0 --ctas 0 2>/dev/null | head -1
