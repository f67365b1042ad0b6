#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_targets.py tests/test_gpu_format.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/bench_next_rows.py 2>&1 | tail -4 | tee gpurun_out/next_rows.jsonl
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_default.err | tee gpurun_out/bench_default.json | cut -c1-400
