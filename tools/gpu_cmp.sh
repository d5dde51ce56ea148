#!/bin/bash
mkdir -p gpurun_out
echo "== defused"; AMGXB_FUSE_PROLONG=0 timeout 300 python tools/profile_solve.py 256 20 2>&1 | head -1
echo "== fused";   AMGXB_FUSE_PROLONG=1 timeout 300 python tools/profile_solve.py 256 20 2>&1 | head -1
echo "== reference GPU build, 256^3, shipped config"
timeout 600 oracle/_ref/ref_dump poisson:256 amgx_b200/configs/PCG_AGGREGATION_JACOBI.json gpurun_out/ref256.bin dDDI 3 2>&1 | tail -12
echo "== reference 128^3"
timeout 600 oracle/_ref/ref_dump poisson:128 amgx_b200/configs/PCG_AGGREGATION_JACOBI.json gpurun_out/ref128.bin dDDI 3 2>&1 | tail -3
rm -f gpurun_out/ref256.bin gpurun_out/ref128.bin
echo "== pytest"; timeout 900 python -m pytest tests -x -q -m gpu --durations=5 2>&1 | tail -12
