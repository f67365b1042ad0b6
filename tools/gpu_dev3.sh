#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_membership.py -m gpu -x -q > gpurun_out/dev_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/dev_pytest.log
echo "== c5 lean(0) vs generic(-5)"
timeout 300 python tools/quick_bench.py --cfg 5 --frames 1 --pool 3 --grids 0 --ctas 0,-5 | tail -4
echo "== c2"
timeout 300 python tools/quick_bench.py --cfg 2 --frames 8 --grids 0 --ctas 0,-5 | tail -3
echo "== c3 shape with N=50016 (W=16, T=512)"
timeout 300 python tools/quick_bench.py --cfg 3 --frames 8 --N 50016 --grids 0 --ctas 0,-5 | tail -3
echo "== kitti 120000 x 768 (W=24)"
timeout 300 python tools/quick_bench.py --cfg 2 --frames 4 --M 768 --grids 0 --ctas 0,-5 | tail -3
