#!/bin/bash
mkdir -p gpurun_out
echo "== dist tests"; timeout 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
for T in 131072; do
  echo "-- tail rows $T"
  AMGXB_TAIL_ROWS=$T timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('value', round(d['value'],1), 'global its/s', round(d['config']['global_iterations_per_sec'],1), 'its', d['config']['iterations_per_step'], 'ms/step', round(d['ms_per_step'],1), 'launches', d.get('gpu_launches'))"
done
