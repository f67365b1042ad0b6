#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_targets.py -m gpu -x -q 2>&1 | tail -15
for l in ${LANES:-4 6 8}; do
  timeout 300 python bench.py --lanes $l --no-cpu-baseline 2>gpurun_out/lanes_$l.err | tee gpurun_out/lanes_$l.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes',$l,d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['clocks'])"
done
