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
run ref --impl reference --steps 3 --warmup 1
run c2_driver --steps 20 --warmup 5
run c2 --steps 2000 --warmup 20
if [ -z "${ONLY_C2:-}" ]; then
run c4 --steps 200 --warmup 5 --workload c4
run c5_split --steps 200 --warmup 5 --workload c5 --partition split
run c5_rep --steps 200 --warmup 5 --workload c5
fi
