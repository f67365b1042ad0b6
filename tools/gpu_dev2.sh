#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_membership.py -m gpu -x -q > gpurun_out/dev_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/dev_pytest.log
python tools/quick_bench.py --cfg 2 --frames 8 --grids 0 --ctas 0,4,-5 | tail -8
python tools/quick_bench.py --cfg 2 --frames 8 --unsorted --grids 0 --ctas 0,-5 | tail -3
python tools/quick_bench.py --cfg 3 --frames 8 --grids 0,96 --ctas 0,-5 | tail -8
python tools/quick_bench.py --cfg 5 --frames 1 --grids 0 --ctas 0,-5 | tail -4
