#!/bin/bash
# ncu capture of the membership kernel on one config: full sections + source counters
set -u
CFG=${1:-2}; TAG=${2:-dev}; shift 2
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pib_sweep -s 10 -c 3 -f -o gpurun_out/${TAG}_prof \
    python tools/quick_bench.py --cfg $CFG "$@" > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/${TAG}_ncu.log
