#!/bin/bash
# developer loop on the GPU box: membership parity tests + timing sweeps
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_membership.py -m gpu -x -q > gpurun_out/dev_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/dev_pytest.log
for cfg in 2 3 5; do
  fr=8; [ $cfg = 5 ] && fr=1
  timeout 300 python tools/quick_bench.py --cfg $cfg --frames $fr --grids ${GRIDS:-0,64,96,128,160,192} --ctas 0 2>&1 | tail -12
done
timeout 300 python tools/quick_bench.py --cfg 2 --frames 8 --unsorted --grids 0,96 2>&1 | tail -3
