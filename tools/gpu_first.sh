#!/bin/bash
# first GPU contact: smoke, golden generation with the reference, kernel timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -20
echo "== golden"; timeout 900 python tests/golden/make_golden.py 2>&1 | tail -40
echo "== bench";  timeout 600 python tools/quick_bench.py 128 256 2>&1 | tail -30
