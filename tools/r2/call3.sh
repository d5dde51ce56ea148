#!/bin/bash
# round-2 GPU call 3 (1 GPU): tile-kernel occupancy / unroll sweep, encoded kernel with the compact layout
mkdir -p gpurun_out/r2
PT="python -m pytest -q -m gpu --timeout=300"
OUT=gpurun_out/r2/sweep.jsonl; : > $OUT
echo "== plain kernel sweep 256"
for cfg in "4 2 4" "4 2 8" "2 4 4" "2 4 8" "3 2 8" "2 3 8" "2 5 8"; do
  set -- $cfg
  AMGXB_TILE_STAGES=$1 AMGXB_TILE_CTAS=$2 AMGXB_TILE_UNROLL=$3 timeout 300 python tools/r2/sweep_kernel.py 256 2>>gpurun_out/r2/sweep.err | tee -a $OUT
done
echo "== plain kernel sweep 512"
for cfg in "4 2 4" "4 2 8" "2 4 8" "2 5 8"; do
  set -- $cfg
  AMGXB_TILE_STAGES=$1 AMGXB_TILE_CTAS=$2 AMGXB_TILE_UNROLL=$3 timeout 400 python tools/r2/sweep_kernel.py 512 2>>gpurun_out/r2/sweep.err | tee -a $OUT
done
echo "== encoded kernel parity (compact layout)"
for E in 1 3; do
  AMGXB_COLENC=$E timeout 600 $PT tests/test_gpu_parity.py tests/test_gpu_resetup.py tests/test_gpu_classical.py -x > gpurun_out/r2/colenc2_parity_$E.log 2>&1; tail -3 gpurun_out/r2/colenc2_parity_$E.log
done
echo "== encoded kernel sweep"
for cfg in "3 0 0" "3 0 3" "3 0 2" "3 2 0" "1 0 0"; do
  set -- $cfg
  AMGXB_COLENC=$1 AMGXB_ENC_STAGES=$2 AMGXB_ENC_CTAS=$3 AMGXB_COLENC_VERBOSE=1 timeout 400 python tools/r2/sweep_kernel.py 256 solve 2>gpurun_out/r2/enc_sweep_$1_$2_$3.err | tee -a $OUT
done
AMGXB_COLENC=3 timeout 500 python tools/r2/sweep_kernel.py 512 solve 2>>gpurun_out/r2/sweep.err | tee -a $OUT
AMGXB_TILE_STAGES=2 AMGXB_TILE_CTAS=4 AMGXB_TILE_UNROLL=8 timeout 400 python tools/r2/sweep_kernel.py 256 solve 2>>gpurun_out/r2/sweep.err | tee -a $OUT
