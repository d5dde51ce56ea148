#!/bin/bash
# usage: gpu_block_scale.sh "1 2 4 8"
mkdir -p gpurun_out
echo "== dist tests"; timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_block_dilu.py -x -q -m gpu 2>&1 | tail -5
for N in $1; do
  for M in dDFI dDDI; do
  if [ "$N" = "1" ]; then
    timeout 600 python tools/bench_block_dist.py $M 128 80 PCG 2>&1 | tail -1
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) tools/bench_block_dist.py $M 128 80 PCG 2>&1 | grep -v Warning | tail -1
  fi
  done
done
timeout 600 python tools/bench_block_dist.py dDFI 128 80 AMG 2>&1 | tail -1
