#!/bin/bash
# Full GPU session: all -m gpu tests, smoke, the reference's own test with the CUDA op injected
# (needs a staged, git-ignored copy of the reference files under _refstage/), bench on all workloads.
set -u
TAG=${1:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
if [ -d _refstage/mmdet3d ]; then
  export GGA_REFERENCE_ROOT=$PWD/_refstage
  timeout 600 python tools/run_reference_box3d_test.py -v > gpurun_out/${TAG}_reference_test_box3d.log 2>&1; echo "reference test rc=$?"
  tail -6 gpurun_out/${TAG}_reference_test_box3d.log
  timeout 600 python -m pytest tests/test_gpu_membership.py tests/test_oracle_membership.py -m "gpu or refonly" -k "reference_own or refonly or wrappers" -v > gpurun_out/${TAG}_refonly.log 2>&1; echo "refonly rc=$?"
  tail -5 gpurun_out/${TAG}_refonly.log
  unset GGA_REFERENCE_ROOT
fi
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 2000 --warmup 20 --no-cpu-baseline > gpurun_out/${TAG}_bench_2000.json 2> gpurun_out/${TAG}_bench_2000.err; echo "bench2000 rc=$?"; cut -c1-250 gpurun_out/${TAG}_bench_2000.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
for WL in c1 c3 c4 c5; do
  timeout 600 python bench.py --steps 200 --warmup 5 --workload $WL > gpurun_out/${TAG}_bench_$WL.json 2> gpurun_out/${TAG}_bench_$WL.err; echo "$WL rc=$?"
  cut -c1-220 gpurun_out/${TAG}_bench_$WL.json; tail -2 gpurun_out/${TAG}_bench_$WL.err
done
