#!/bin/bash
# round-2 GPU call 2: full gpu suite (gate removed), encoded streams parity + bench, perf switches, reference GPU timing
mkdir -p gpurun_out/r2
PT="python -m pytest -q -m gpu --timeout=300"
echo "== full suite"
timeout 900 $PT tests -rf > gpurun_out/r2/suite2.log 2>&1; tail -25 gpurun_out/r2/suite2.log
echo "== colenc parity"
for E in 1 3; do
  AMGXB_COLENC=$E AMGXB_COLENC_VERBOSE=1 timeout 600 $PT tests/test_gpu_parity.py tests/test_gpu_resetup.py -x > gpurun_out/r2/colenc_parity_$E.log 2>&1; tail -4 gpurun_out/r2/colenc_parity_$E.log
done
echo "== colenc bench"
for E in 0 1 3; do
  AMGXB_COLENC=$E AMGXB_COLENC_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2/colenc_bench_$E.err > gpurun_out/r2/colenc_bench_$E.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2/colenc_bench_$E.json").read().strip().splitlines()[-1]); r=d['roofline']
    print('  COLENC=$E: its/s', round(d['value'],1), 'SpMV ms', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3), 'Jacobi ms', round(r['fused_jacobi_sweep']['ms_per_launch'],4), 'iters', d['config']['iterations_per_step'])
except Exception as e: print('  COLENC=$E failed', e)
PY
done
echo "== 512^3 bench"
timeout 900 python bench.py --grid 512 --steps 2 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2/bench512.err > gpurun_out/r2/bench512.json; tail -c 900 gpurun_out/r2/bench512.json
echo "== reference GPU build, same config, 256^3 (3 reps)"
REFDUMP_NO_LEVELS=1 timeout 600 oracle/_ref/ref_dump poisson:256 amgx_b200/configs/PCG_AGGREGATION_JACOBI.json /tmp/ref256.bin dDDI 3 2>&1 | tail -2 | tee gpurun_out/r2/ref_gpu_256.log
echo "== perf switches"
timeout 900 python tools/bench_configs.py block_pg 2>&1 | grep '^{' | tee gpurun_out/r2/block_coloring.json | cut -c1-300
for G in 0 1; do
  AMGXB_GRAPH_COARSE=$G timeout 900 python tools/bench_classical.py 128 256 2>&1 | grep '"engine": "ours"' | tee gpurun_out/r2/graph_coarse_$G.json | cut -c1-300
done
