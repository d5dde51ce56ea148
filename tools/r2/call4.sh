#!/bin/bash
# round-2 GPU call 4 (2 GPUs): peer-memory exchange / all-reduce vs NCCL path: parity tests (incl. the never-run multi-GPU sections), bench N=2
mkdir -p gpurun_out/r2
nvidia-smi topo -m > gpurun_out/r2/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== 2-GPU parity tests, peer-memory path"
AMGXB_P2P_VERBOSE=1 AMGXB_RUN_UNVALIDATED=1 timeout 1200 python -m pytest tests/test_gpu_dist.py -q -m gpu -k "2-" -rf > gpurun_out/r2/dist2_p2p.log 2>&1; tail -30 gpurun_out/r2/dist2_p2p.log
echo "== 2-GPU parity tests, NCCL path (regression)"
AMGXB_P2P=0 AMGXB_RUN_UNVALIDATED=1 timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu -k "2-0" -rf > gpurun_out/r2/dist2_nccl.log 2>&1; tail -5 gpurun_out/r2/dist2_nccl.log
echo "== bench N=2"
for V in "1 0" "0 0" "1 131072" "1 2000000"; do
  set -- $V
  AMGXB_P2P=$1 AMGXB_TAIL_ROWS=$2 timeout 600 $TR --master-port 2972$1 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench2_p2p$1_tail$2.json 2> gpurun_out/r2/bench2_p2p$1_tail$2.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2/bench2_p2p$1_tail$2.json") if l.startswith('{')][-1])
    print("  P2P=$1 TAIL=$2: value", round(d['value'],1), "global it/s", round(d['config']['global_iterations_per_sec'],1), "iters", d['config']['iterations_per_step'], d['config']['solve_status'], "launches", d['gpu_launches'], "parity", json.dumps(d.get('parity'))[:600])
except Exception as e: print("  P2P=$1 TAIL=$2 failed", e)
PY
done
echo "== strong 256^3 over 2"
timeout 600 $TR --master-port 29741 bench.py --gpus 2 --strong --grid 256 --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>gpurun_out/r2/bench2_strong.err | tee gpurun_out/r2/bench2_strong.json | cut -c1-400
