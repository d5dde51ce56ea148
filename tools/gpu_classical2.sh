#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== bench classical"; timeout 2400 python tools/bench_classical.py 256 512 2>&1 | tail -8
