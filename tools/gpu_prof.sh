#!/bin/bash
# ncu capture of the membership kernels on one config: launch list + full sections
set -u
CFG=${1:-2}; FR=${2:-8}; TAG=${3:-dev}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/quick_bench.py --cfg $CFG --frames $FR --grids 0 --ctas 0 > gpurun_out/${TAG}_l.log 2>&1; echo "launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pib -s 10 -c 4 -f -o gpurun_out/${TAG}_prof \
    python tools/quick_bench.py --cfg $CFG --frames $FR --grids 0 --ctas 0 > gpurun_out/${TAG}_f.log 2>&1; echo "full rc=$?"
