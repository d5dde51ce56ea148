#!/bin/bash
# round-2 GPU call 17 (8 GPUs): 4-GPU parity tests, weak scaling N=8 on both exchange paths, config 4 as stated (512^3 strong over 8), config 5 over 8
mkdir -p gpurun_out/r2/final
F=gpurun_out/r2/final
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== 4-GPU parity tests (peer-memory path): whole hierarchy distributed, and replicated tail"
timeout 900 python -m pytest -q -m gpu --timeout=600 tests/test_gpu_dist.py -k "4-0 or 4-131072" 2>&1 | tail -3 | cut -c1-300
run() { # N P2P extra-args tag
  N=$1; P=$2; TAG=$4
  AMGXB_P2P=$P timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2979$P bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline $3 2>$F/scale_$TAG.err | grep '^{' > $F/scale_$TAG.json
  python - <<PY
import json
try:
    d=json.loads(open("$F/scale_$TAG.json").read().strip().splitlines()[-1])
    p=d.get('parity') or {}
    print("  $TAG: value", round(d['value'],1), "global it/s", round(d['config']['global_iterations_per_sec'],1), "iters", d['config']['iterations_per_step'], d['config']['solve_status'], "launches/it", round(d['gpu_launches']/3/d['config']['iterations_per_step']), "e2e", round(d['e2e']['value'],1), "parity", p.get('green'), (p.get('amg') or {}).get('iterations_distributed'), (p.get('amg') or {}).get('iterations_single_rank'), p.get('error'))
except Exception as e: print("  $TAG failed", e)
PY
}
echo "== weak scaling (256^3 per GPU)"
run 8 1 "" weak8_p2p
run 8 0 "" weak8_nccl
run 4 1 "" weak4_p2p
echo "== config 4 as stated: 512^3 strong over 8 GPUs (1-GPU figure: final/bench_poisson512.json)"
run 8 1 "--strong --grid 512 --no-parity" strong512_8_p2p
echo "== config 5 over 8 GPUs: 4x4 blocks, 160x160x(160*N) block rows = 32.8 M, dDFI, AMG + DILU"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29795 tools/bench_block_dist.py dDFI 160 160 AMG 2>/dev/null | grep '^{' | tee $F/block_scale_8.json | cut -c1-400
