#!/bin/bash
# Multi-GPU part of the next round's validation (charged N x box time, so only what needs two ranks):
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/validate_next_round_2gpu.sh'
mkdir -p gpurun_out
echo "== validated 2-GPU suite"
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -4
echo "== unvalidated: partition vectors, comm maps, read_system_distributed, replace_coefficients, DENSE_LU on partitions, replicated tail"
AMGXB_RUN_UNVALIDATED=1 timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/unvalidated_2gpu.log
echo "== replicated tail (partitioned aggregates), 2 GPUs: iterations must equal the tail-off run"
for T in 0 131072; do
  AMGXB_TAIL_ROWS=$T timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --grid 128 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('  tail $T: its', d['config']['iterations_per_step'], d['config']['solve_status'], 'global its/s', round(d['config']['global_iterations_per_sec'],1))"
done
echo "== 4 ranks need --gpus 4: AMGXB_TAIL_ROWS=131072 vs 0 at --grid 96 must both give 51 iterations"
echo "== config 3 (FGMRES + classical AMG) on 1 and 2 GPUs, same global 128^3 problem: same iteration count expected"
timeout 300 python tools/bench_classical_dist.py 128 2>&1 | grep '^{' | tee gpurun_out/cla_dist_1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tools/bench_classical_dist.py 128 2>&1 | grep '^{' | tee gpurun_out/cla_dist_2.json
