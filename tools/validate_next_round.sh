#!/bin/bash
# First GPU call of the next round: everything that was written after round 1 ran out of GPU minutes.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/validate_next_round.sh'
mkdir -p gpurun_out
echo "== unvalidated: DENSE_LU_SOLVER, W/F cycles"
AMGXB_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests/test_gpu_dense_lu.py tests/test_gpu_cycles.py -q -m gpu 2>&1 | tail -6
echo "== full gpu suite"
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== replicated tail (partitioned aggregates), 2 GPUs: iterations must equal the tail-off run"
for T in 0 131072; do
  AMGXB_TAIL_ROWS=$T timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --grid 128 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('  tail $T: its', d['config']['iterations_per_step'], d['config']['solve_status'], 'global its/s', round(d['config']['global_iterations_per_sec'],1))"
done
echo "== 4 ranks need --gpus 4: AMGXB_TAIL_ROWS=131072 vs 0 at --grid 96 must both give 51 iterations"
