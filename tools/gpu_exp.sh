#!/bin/bash
# Experiment session: step + membership tests, isolated timings (product build, then the profiling
# build's variants and timeline), the host-buffer pipeline sweep, one short bench line.
set -u
TAG=${1:-exp}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_membership.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_pytest.log
Q="timeout 300 python tools/quick_bench.py"
{
$Q --cfg 2
$Q --cfg 2
$Q --cfg 1 --frames 8
$Q --cfg 3
$Q --cfg 5
export GGA_B200_LIB=$PWD/gga_b200/_C/libgga_b200_prof.so
$Q --cfg 2 --variant ${TRACEVAR:-0} --trace
unset GGA_B200_LIB
} > gpurun_out/${TAG}_qb.log 2>&1
grep -v copy_same gpurun_out/${TAG}_qb.log | cut -c1-200
if [ -n "${E2E:-}" ]; then timeout 280 python tools/bench_e2e.py > gpurun_out/${TAG}_e2e.jsonl 2>&1; cat gpurun_out/${TAG}_e2e.jsonl; fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${TAG}_bench.json') if l.startswith('{')][0])
print('value', d['value'], 'ms', d['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'])
print('e2e', json.dumps(d['e2e'])[:900])
PY
tail -3 gpurun_out/${TAG}_bench.err
