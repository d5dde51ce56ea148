#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_scale.sh "1 2 4 8"
for N in 1 2 4 8; do
  for M in dDFI; do
  if [ "$N" = "1" ]; then
    timeout 600 python tools/bench_block_dist.py $M 128 80 PCG 2>&1 | tail -1 | tee gpurun_out/block_scale_$N.json | cut -c1-420
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) tools/bench_block_dist.py $M 128 80 PCG 2>&1 | grep '^{' | tail -1 | tee gpurun_out/block_scale_$N.json | cut -c1-420
  fi
  done
done
