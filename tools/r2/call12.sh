#!/bin/bash
# round-2 GPU call 12 (2 GPUs): where does an iteration's time go at N = 2 (phase timing, graphs off) + coding statistics of the partitioned levels
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest -q -m gpu --timeout=300 tests/test_gpu_parity.py tests/test_gpu_classical.py tests/test_gpu_dist.py -k "not 4-" 2>&1 | tail -2 | cut -c1-300
for T in 0 131072; do
AMGXB_PHASE_TIMING=1 AMGXB_COLENC_VERBOSE=1 AMGXB_TAIL_ROWS=$T timeout 600 $TR --master-port 29761 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --no-parity > /dev/null 2> gpurun_out/r2/phase_n2_tail$T.txt
echo "== N=2 tail=$T"; grep "colenc level" gpurun_out/r2/phase_n2_tail$T.txt | head -12 | cut -c1-260; grep -A58 "phase timing\]" gpurun_out/r2/phase_n2_tail$T.txt | head -60 | cut -c1-160
done
echo "== N=2 tail=131072 NCCL path"
AMGXB_P2P=0 AMGXB_PHASE_TIMING=1 AMGXB_TAIL_ROWS=131072 timeout 600 $TR --master-port 29763 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --no-parity > /dev/null 2> gpurun_out/r2/phase_n2_nccl_tail.txt
grep -A58 "phase timing\]" gpurun_out/r2/phase_n2_nccl_tail.txt | head -60 | cut -c1-160
echo "== N=1 on the same box, phase timing"
AMGXB_PHASE_TIMING=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-reference-gpu > /dev/null 2> gpurun_out/r2/phase_n1_same_box.txt; grep -A12 "phase timing\]" gpurun_out/r2/phase_n1_same_box.txt | head -14 | cut -c1-160
echo "== bench N=2 (graphs on) for the record"
for V in "1 0" "1 131072" "0 131072"; do set -- $V; T=$2
AMGXB_P2P=$1 AMGXB_TAIL_ROWS=$T timeout 600 $TR --master-port 29762 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/r2/bench2d_p$1_tail$T.json
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2/bench2d_p$1_tail$T.json").read().strip().splitlines()[-1])
    print("  p2p=$1 tail=$T: value", round(d['value'],1), "global it/s", round(d['config']['global_iterations_per_sec'],1), "iters", d['config']['iterations_per_step'], "launches/it", round(d['gpu_launches']/3/d['config']['iterations_per_step']), "parity", (d.get('parity') or {}).get('green'))
except Exception as e: print("  tail=$T failed", e)
PY
done
