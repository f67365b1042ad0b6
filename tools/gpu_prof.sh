#!/bin/bash
# Profiling session for profiles/: ncu launch list of the bench command, full capture of the membership
# kernel, steady-state DRAM traffic.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pib_sweep -s 30 -c 3 -f -o gpurun_out/${TAG}_prof_pib \
    python tools/quick_bench.py --cfg 2 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none \
    -k regex:pib_sweep -s 28 -c 56 --csv --log-file gpurun_out/${TAG}_traffic.csv python tools/traffic.py --workload c2 \
    > gpurun_out/${TAG}_traffic.log 2>&1; echo "traffic rc=$?"
python tools/traffic.py --digest gpurun_out/${TAG}_traffic.csv --workload c2 | tee gpurun_out/${TAG}_traffic.json
