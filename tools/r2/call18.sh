#!/bin/bash
# round-2 GPU call 18 (4 GPUs): where should the replicated tail start when N grows (threshold in global rows: 131072 / 262144 / 524288)
mkdir -p gpurun_out/r2
for T in 131072 262144 524288; do
AMGXB_TAIL_ROWS=$T timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | grep '^{' > gpurun_out/r2/bench4_tail$T.json
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2/bench4_tail$T.json").read().strip().splitlines()[-1])
    print("  N=4 tail=$T: value", round(d['value'],1), "global it/s", round(d['config']['global_iterations_per_sec'],1), "iters", d['config']['iterations_per_step'], "launches/it", round(d['gpu_launches']/3/d['config']['iterations_per_step']))
except Exception as e: print("  tail=$T failed", e)
PY
done
