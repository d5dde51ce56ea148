#!/bin/bash
mkdir -p gpurun_out
run() {
  echo "-- TAIL=$1 FUSE=$2 grid=$3"
  AMGXB_TAIL_ROWS=$1 AMGXB_DIST_FUSE=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29744 bench.py --gpus 4 --grid $3 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('  its', d['config']['iterations_per_step'], d['config']['solve_status'], 'global its/s', round(d['config']['global_iterations_per_sec'],1))"
}
run 0 0 96
run 0 1 96
run 131072 0 96
run 131072 1 96
run 2000 0 96
