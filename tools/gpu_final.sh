#!/bin/bash
# Final sanity of a session: full GPU suite, smoke, both bench arms (no ncu).
set -u
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
timeout 300 python tools/bench_configs.py 2>/dev/null | tee gpurun_out/${TAG}_other_configs.jsonl | cut -c1-260
