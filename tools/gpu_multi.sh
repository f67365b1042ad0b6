#!/bin/bash
# bench.py on N GPUs of one box (torchrun, one rank per GPU): default workload + c4 (all-gather) + c5 split
set -u
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
run() {  # name, args...
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N "$@" > gpurun_out/${TAG}_n${N}_$name.json 2> gpurun_out/${TAG}_n${N}_$name.err
  echo "$name rc=$?"; grep '^{' gpurun_out/${TAG}_n${N}_$name.json | cut -c1-330; grep -i "error\|Traceback" gpurun_out/${TAG}_n${N}_$name.err | head -3
}
RUNS=${RUNS:-"ref c2_driver c2 c4 c5_split c5_rep"}
for r in $RUNS; do
  case $r in
    ref)       run ref --impl reference --steps 3 --warmup 1 ;;
    c2_driver) run c2_driver --steps 20 --warmup 5 ;;
    c2)        run c2 --steps 2000 --warmup 20 --no-cpu-baseline ;;
    c4)        run c4 --steps 200 --warmup 5 --workload c4 --no-cpu-baseline ;;
    c5_split)  run c5_split --steps 200 --warmup 5 --workload c5 --partition split --no-cpu-baseline ;;
    c5_rep)    run c5_rep --steps 200 --warmup 5 --workload c5 --no-cpu-baseline ;;
  esac
done
