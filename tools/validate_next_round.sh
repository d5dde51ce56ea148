#!/bin/bash
# First GPU call of the next round: everything that was written after round 1 ran out of GPU minutes.
#   gpurun --timeout 2400 -- 'bash tools/validate_next_round.sh'                  (ONE GPU: everything but the multi-GPU sections)
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/validate_next_round_2gpu.sh'    (charged twice: only what needs two ranks)
# 1. regular gpu suite (must stay green), 2. the opt-in tests of the unvalidated components, 3. reference goldens for them
# (tests/golden/cases_round2.py through oracle/_ref/ref_dump) -> copy gpurun_out/golden/r2_*.npz to tests/golden/ and commit,
# 4. the experimental compressed column stream and the colouring comparison.
mkdir -p gpurun_out
echo "== full gpu suite (validated components)"
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== unvalidated: DENSE_LU, W/F/CG/CGF cycles, error_scaling, CG/PCGF/PBICGSTAB/GMRES, MULTICOLOR_GS, CHEBYSHEV(_POLY), resetup, HMIS, partition vectors, comm maps, replicated tail"
AMGXB_RUN_UNVALIDATED=1 timeout 1200 python -m pytest tests/test_gpu_dense_lu.py tests/test_gpu_cycles.py tests/test_gpu_krylov.py tests/test_gpu_smoothers.py \
    tests/test_gpu_resetup.py tests/test_golden_round2.py tests/test_gpu_classical.py -q -m gpu 2>&1 | tail -30 | tee gpurun_out/unvalidated.log
echo "== reference goldens for the round-2 cases"
timeout 900 python tests/golden/make_golden.py r2 2>&1 | tail -40 | tee gpurun_out/make_golden_r2.log
echo "== experimental compressed matrix streams (csrc/k_spmv_enc.cu): AMGXB_COLENC=1 columns, 3 columns + value dictionaries; parity suite, then the bench"
for E in 1 3; do
  AMGXB_COLENC=$E AMGXB_COLENC_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/colenc_parity_$E.log
done
# value codes must follow in-place value changes: replace_coefficients + resetup, DIAGONAL_SYMMETRIC scaling
AMGXB_COLENC=3 AMGXB_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests/test_gpu_resetup.py tests/test_golden_round2.py -q -m gpu -k "resetup or diagsym or replace" 2>&1 | tail -8 | tee gpurun_out/colenc_values_changed.log
for E in 0 1 3; do
  AMGXB_COLENC=$E timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/colenc_bench_$E.err | grep '^{' | tee gpurun_out/colenc_bench_$E.json | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('  COLENC=$E: its/s', round(d['value'],1), 'SpMV ms', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3), 'Jacobi ms', round(r['fused_jacobi_sweep']['ms_per_launch'],4))"
done
echo "== config 5 (block 4x4 DILU): MIN_MAX vs PARALLEL_GREEDY colouring"
timeout 900 python tools/bench_configs.py block_pg 2>&1 | grep '^{' | tee gpurun_out/block_coloring.json | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('  ', d['case'], 'colors', d['colors_L0'], 'iters', d['iters'], 'its/s', round(d['iters_per_s'],1), d['status'])"
echo "== config 3: coarse levels of the preconditioner V-cycle replayed as one CUDA graph (AMGXB_GRAPH_COARSE=1): parity, then 128^3 / 256^3 with and without"
AMGXB_GRAPH_COARSE=1 timeout 600 python -m pytest tests/test_gpu_classical.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/graph_coarse_parity.log
for G in 0 1; do
  AMGXB_GRAPH_COARSE=$G timeout 900 python tools/bench_classical.py 128 256 2>&1 | grep '"engine": "ours"' | tee gpurun_out/graph_coarse_$G.json | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('  GRAPH_COARSE=$G nx', d['nx'], 'iters', d['iters'], 'its/s', round(d['iters_per_s'],1), 'launches', d['launches'])"
done
