#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for M in dDFI dDDI; do timeout 600 python tools/bench_block_dist.py $M 128 80 PCG 2>&1 | tail -1; done
timeout 600 python tools/bench_block_dist.py dDFI 128 80 AMG 2>&1 | tail -1
timeout 600 python tools/bench_block_dist.py dDFI 160 160 AMG 2>&1 | tail -1
