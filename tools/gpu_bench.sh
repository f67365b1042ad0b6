#!/bin/bash
# bench.py on one GPU: default workload with a few lane counts, then the other workloads
set -u
TAG=${1:-dev}
mkdir -p gpurun_out
for L in 1 2 3; do
  timeout 600 python bench.py --steps 2000 --warmup 20 --lanes $L --no-cpu-baseline > gpurun_out/${TAG}_bench_l$L.json 2> gpurun_out/${TAG}_bench_l$L.err; echo "lanes $L rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench_l$L.json'))
    print({k:d[k] for k in ('value','ms_per_step','lanes')}, d['sequential_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/${TAG}_bench_l$L.err').read()[-1500:])
PY
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_driver.json 2> gpurun_out/${TAG}_bench_driver.err; echo "driver-like rc=$?"; cut -c1-600 gpurun_out/${TAG}_bench_driver.json
for WL in c1 c3 c4 c5; do
  timeout 600 python bench.py --steps 200 --warmup 5 --workload $WL --no-cpu-baseline > gpurun_out/${TAG}_bench_$WL.json 2> gpurun_out/${TAG}_bench_$WL.err; echo "$WL rc=$?"
  cut -c1-300 gpurun_out/${TAG}_bench_$WL.json; tail -3 gpurun_out/${TAG}_bench_$WL.err
done
