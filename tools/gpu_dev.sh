#!/bin/bash
# One developer GPU session: membership parity tests, then isolated timings of the membership
# kernel (product build, then the GGA_PROFILING build with CTA-size / ranges sweeps + timeline).
set -u
TAG=${1:-dev}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
if [ -z "${SKIP_TESTS:-}" ]; then
timeout 900 python -m pytest tests/test_gpu_membership.py tests/test_gpu_step.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${TAG}_pytest.log
fi
Q="timeout 300 python tools/quick_bench.py"
{
$Q --cfg 2
$Q --cfg 2 --unsorted
$Q --cfg 1 --frames 8
$Q --cfg 3
$Q --cfg 5
$Q --cfg 2 --mode part
$Q --cfg 2 --mode all
export GGA_B200_LIB=$PWD/gga_b200/_C/libgga_b200_prof.so
$Q --cfg 2 --nt 1024 --variant 0 --trace
$Q --cfg 2 --nt 1024 --variant 1,8,9
} > gpurun_out/${TAG}_qb.log 2>&1
cat gpurun_out/${TAG}_qb.log
