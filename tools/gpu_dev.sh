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
$Q --cfg 2
$Q --cfg 1 --frames 8
$Q --cfg 3
$Q --cfg 5
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pib_sweep -s 20 -c 2 python tools/quick_bench.py --cfg 2 2>&1 | grep -E "inst_executed|time_duration|issue_active"
} > gpurun_out/${TAG}_qb.log 2>&1
cat gpurun_out/${TAG}_qb.log
