#!/bin/bash
mkdir -p gpurun_out
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -8
echo "== pytest gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -40
echo "== bench";  timeout 600 python tools/quick_bench.py 128 256 2>&1 | tail -30
