#!/bin/bash
# round-2 GPU call 5 (2 GPUs): fused one-kernel exchange on unsplit levels, coded streams on row-partitioned matrices
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== 2-GPU parity tests, peer-memory path (all sections)"
AMGXB_P2P_VERBOSE=1 AMGXB_RUN_UNVALIDATED=1 timeout 1200 python -m pytest tests/test_gpu_dist.py -q -m gpu -k "2-" -rf > gpurun_out/r2/dist2b_p2p.log 2>&1; tail -8 gpurun_out/r2/dist2b_p2p.log | cut -c1-300
echo "== same, NCCL path, plain kernels"
AMGXB_P2P=0 AMGXB_COLENC=0 AMGXB_RUN_UNVALIDATED=1 timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu -k "2-0" -rf > gpurun_out/r2/dist2b_nccl.log 2>&1; tail -4 gpurun_out/r2/dist2b_nccl.log | cut -c1-300
echo "== bench N=2"
run() { # P2P TAIL SPLIT COLENC
  AMGXB_P2P=$1 AMGXB_TAIL_ROWS=$2 AMGXB_SPLIT_ROWS=$3 AMGXB_COLENC=$4 timeout 600 $TR --master-port 2973$1 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench2b_$1_$2_$3_$4.json 2> gpurun_out/r2/bench2b_$1_$2_$3_$4.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2/bench2b_$1_$2_$3_$4.json") if l.startswith('{')][-1])
    p=d.get('parity') or {}
    print("  P2P=$1 TAIL=$2 SPLIT=$3 COLENC=$4: value", round(d['value'],1), "global it/s", round(d['config']['global_iterations_per_sec'],1), "iters", d['config']['iterations_per_step'], d['config']['solve_status'], "launches/it", round(d['gpu_launches']/3/d['config']['iterations_per_step']), "parity green", p.get('green'), p.get('error'))
except Exception as e: print("  P2P=$1 TAIL=$2 SPLIT=$3 COLENC=$4 failed", e)
PY
}
run 1 0 1048576 3
run 1 0 1048576 0
run 1 131072 1048576 3
run 1 0 0 3
run 0 0 1048576 3
run 0 131072 1048576 3
echo "== strong 256^3 over 2"
timeout 600 $TR --master-port 29741 bench.py --gpus 2 --strong --grid 256 --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>gpurun_out/r2/bench2b_strong.err | tee gpurun_out/r2/bench2b_strong.json | cut -c1-300
