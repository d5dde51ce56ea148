#!/bin/bash
# round-2 GPU call 15 (2 GPUs): exchange-kernel grid size, split threshold, tail default; final N=2 lines
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest -q -m gpu --timeout=600 tests/test_gpu_dist.py -k "2-" 2>&1 | tail -2 | cut -c1-300
run() { # P2P CTAS SPLIT TAIL
  AMGXB_P2P=$1 AMGXB_P2P_CTAS=$2 AMGXB_SPLIT_ROWS=$3 AMGXB_TAIL_ROWS=$4 timeout 600 $TR --master-port 2977$1 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/r2/bench2e_$1_$2_$3_$4.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2/bench2e_$1_$2_$3_$4.json").read().strip().splitlines()[-1])
    print("  p2p=$1 ctas=$2 split=$3 tail=$4: value", round(d['value'],1), "global it/s", round(d['config']['global_iterations_per_sec'],1), "iters", d['config']['iterations_per_step'], "launches/it", round(d['gpu_launches']/3/d['config']['iterations_per_step']), "parity", (d.get('parity') or {}).get('green'))
except Exception as e: print("  p2p=$1 ctas=$2 split=$3 tail=$4 failed", e)
PY
}
run 1 8 1048576 131072
run 1 2 1048576 131072
run 1 32 1048576 131072
run 1 8 1073741824 131072
run 1 8 4194304 131072
run 0 8 1048576 131072
run 1 8 1048576 262144
echo "== strong 256^3 over 2 GPUs and --impl reference arm"
timeout 600 $TR --master-port 29781 bench.py --gpus 2 --strong --grid 256 --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  strong 256^3 N=2: it/s', round(d['value'],1), d['scaling'])"
