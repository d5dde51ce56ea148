#!/bin/bash
# First GPU calls of the next round: everything that was written after round 1 ran out of GPU minutes, in stages so that one call stays
# short and a hang costs one stage only (each pytest run carries a per-test timeout).  ONE GPU:
#   gpurun --timeout 600 -- 'bash tools/validate_next_round.sh regress'     validated suite + smoke + default bench (the edits since the last
#                                                                            device run touched validated files: this must be green first)
#   gpurun --timeout 900 -- 'bash tools/validate_next_round.sh unvalidated' opt-in tests of the new components against the CPU oracle
#   gpurun --timeout 900 -- 'bash tools/validate_next_round.sh goldens'     reference binary on tests/golden/cases_round2.py -> gpurun_out/golden/r2_*.npz
#   gpurun --timeout 900 -- 'bash tools/validate_next_round.sh colenc'      compressed matrix streams: parity, value-change hooks, bench 0 / 1 / 3
#   gpurun --timeout 900 -- 'bash tools/validate_next_round.sh perf'        coarse-level graph for config 3, PARALLEL_GREEDY colours for config 5
#   (no argument: all stages in that order; needs --timeout 3000)
# Two GPUs: tools/validate_next_round_2gpu.sh
mkdir -p gpurun_out
STAGE=${1:-all}
PT="python -m pytest -q -m gpu --timeout=300"

if [ "$STAGE" = regress ] || [ "$STAGE" = all ]; then
  echo "== validated gpu suite"
  timeout 900 $PT tests -x 2>&1 | tail -6 | tee gpurun_out/regress_suite.log
  echo "== smoke + default bench"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3 | tee gpurun_out/regress_smoke.log
  timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/regress_bench.err | grep '^{' | tee gpurun_out/regress_bench.json | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('  its/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'SpMV frac', round(r['frac'],3), 'iteration', r.get('iteration'))"
fi

if [ "$STAGE" = unvalidated ] || [ "$STAGE" = all ]; then
  echo "== unvalidated: DENSE_LU, W/F/CG/CGF cycles, error_scaling, CG/PCGF/PBICGSTAB/GMRES, MULTICOLOR_GS, CHEBYSHEV(_POLY), resetup (aggregation + classical), HMIS, D1, SIZE_4"
  AMGXB_RUN_UNVALIDATED=1 timeout 1500 $PT tests/test_gpu_dense_lu.py tests/test_gpu_cycles.py tests/test_gpu_krylov.py tests/test_gpu_smoothers.py \
      tests/test_gpu_resetup.py tests/test_gpu_classical.py tests/test_gpu_edge_cases.py tests/test_capi_graceful_failure.py tests/test_golden_round2.py 2>&1 | tail -40 | tee gpurun_out/unvalidated.log
fi

if [ "$STAGE" = goldens ] || [ "$STAGE" = all ]; then
  echo "== reference goldens for the round-2 cases"
  timeout 900 python tests/golden/make_golden.py r2 2>&1 | tail -60 | tee gpurun_out/make_golden_r2.log
fi

if [ "$STAGE" = colenc ] || [ "$STAGE" = all ]; then
  echo "== experimental compressed matrix streams (csrc/k_spmv_enc.cu): AMGXB_COLENC=1 columns, 3 columns + value dictionaries"
  for E in 1 3; do
    AMGXB_COLENC=$E AMGXB_COLENC_VERBOSE=1 timeout 600 $PT tests/test_gpu_parity.py -x 2>&1 | tail -15 | tee gpurun_out/colenc_parity_$E.log
  done
  # value codes must follow in-place value changes: replace_coefficients + resetup, DIAGONAL_SYMMETRIC scaling
  AMGXB_COLENC=3 AMGXB_RUN_UNVALIDATED=1 timeout 600 $PT tests/test_gpu_resetup.py tests/test_golden_round2.py -k "resetup or diagsym or replace" 2>&1 | tail -8 | tee gpurun_out/colenc_values_changed.log
  for E in 0 1 3; do
    AMGXB_COLENC=$E timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/colenc_bench_$E.err | grep '^{' | tee gpurun_out/colenc_bench_$E.json | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('  COLENC=$E: its/s', round(d['value'],1), 'SpMV ms', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3), 'Jacobi ms', round(r['fused_jacobi_sweep']['ms_per_launch'],4))"
  done
fi

if [ "$STAGE" = perf ] || [ "$STAGE" = all ]; then
  echo "== config 5 (block 4x4 DILU): MIN_MAX vs PARALLEL_GREEDY colouring"
  timeout 900 python tools/bench_configs.py block_pg 2>&1 | grep '^{' | tee gpurun_out/block_coloring.json | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('  ', d['case'], 'colors', d['colors_L0'], 'iters', d['iters'], 'its/s', round(d['iters_per_s'],1), d['status'])"
  echo "== config 3: coarse levels of the preconditioner V-cycle replayed as one CUDA graph (AMGXB_GRAPH_COARSE=1): parity, then 128^3 / 256^3 with and without"
  AMGXB_GRAPH_COARSE=1 timeout 600 $PT tests/test_gpu_classical.py -x 2>&1 | tail -4 | tee gpurun_out/graph_coarse_parity.log
  for G in 0 1; do
    AMGXB_GRAPH_COARSE=$G timeout 900 python tools/bench_classical.py 128 256 2>&1 | grep '"engine": "ours"' | tee gpurun_out/graph_coarse_$G.json | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('  GRAPH_COARSE=$G nx', d['nx'], 'iters', d['iters'], 'its/s', round(d['iters_per_s'],1), 'launches', d['launches'])"
  done
fi
