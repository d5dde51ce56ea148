#!/bin/bash
# round-2 GPU call 30 (2 GPUs): the final tree on the row-partitioned path: 2-GPU parity suite (tail 131072 only) + one bench line with the parity object
mkdir -p gpurun_out/r2/final
F=gpurun_out/r2/final
timeout 600 python -m pytest -q -m gpu --timeout=500 tests/test_gpu_dist.py -k "2-131072 or 2-600" 2>&1 | tail -3 | cut -c1-300
AMGXB_P2P=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29791 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>$F/scale_weak2_p2p_final.err | grep '^{' > $F/scale_weak2_p2p_final.json
python - <<PY
import json
try:
    d=json.loads(open("$F/scale_weak2_p2p_final.json").read().strip().splitlines()[-1])
    p=d.get('parity') or {}
    print("weak2 p2p: value", round(d['value'],1), "global it/s", round(d['config']['global_iterations_per_sec'],1), "iters", d['config']['iterations_per_step'], d['config']['solve_status'], "e2e", round(d['e2e']['value'],1), "parity", p.get('green'), (p.get('amg') or {}).get('iterations_distributed'), (p.get('amg') or {}).get('iterations_single_rank'), p.get('error'))
except Exception as e: print("failed", e); print(open("$F/scale_weak2_p2p_final.err").read()[-1500:])
PY
