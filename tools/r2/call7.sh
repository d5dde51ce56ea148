#!/bin/bash
# round-2 GPU call 7 (2 GPUs): single-kernel exchange (+ side-stream overlap on split levels), receiver-driven send maps, latency micro-benchmark
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== 2-GPU parity tests, peer-memory path (all sections)"
AMGXB_P2P_VERBOSE=1 AMGXB_RUN_UNVALIDATED=1 timeout 1200 python -m pytest tests/test_gpu_dist.py -q -m gpu -k "2-" -rf > gpurun_out/r2/dist2c_p2p.log 2>&1; tail -8 gpurun_out/r2/dist2c_p2p.log | cut -c1-300
echo "== same, NCCL path"
AMGXB_P2P=0 AMGXB_RUN_UNVALIDATED=1 timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu -k "2-0" -rf > gpurun_out/r2/dist2c_nccl.log 2>&1; tail -4 gpurun_out/r2/dist2c_nccl.log | cut -c1-300
echo "== latency of small distributed operations"
for P in 1 0; do AMGXB_P2P=$P timeout 300 $TR --master-port 2975$P tools/r2/dist_latency.py 2>/dev/null | grep '^{' | tee gpurun_out/r2/dist_latency_$P.json; done
echo "== bench N=2"
run() { # P2P TAIL
  AMGXB_P2P=$1 AMGXB_TAIL_ROWS=$2 timeout 600 $TR --master-port 2973$1 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench2c_$1_$2.json 2> gpurun_out/r2/bench2c_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2/bench2c_$1_$2.json") if l.startswith('{')][-1])
    p=d.get('parity') or {}
    print("  P2P=$1 TAIL=$2: value", round(d['value'],1), "global it/s", round(d['config']['global_iterations_per_sec'],1), "iters", d['config']['iterations_per_step'], d['config']['solve_status'], "launches/it", round(d['gpu_launches']/3/d['config']['iterations_per_step']), "parity green", p.get('green'), p.get('error'))
except Exception as e: print("  P2P=$1 TAIL=$2 failed", e)
PY
}
run 1 0
run 0 0
run 1 131072
run 0 131072
run 1 1000000
