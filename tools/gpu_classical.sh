#!/bin/bash
mkdir -p gpurun_out
echo "== dev_classical"; timeout 900 python tools/dev_classical.py 2>&1 | tail -60
echo "== goldens"; timeout 900 python tests/golden/make_golden.py poisson12_fgmres_classical_aggr poisson16x12x9_fgmres_classical_aggr poisson12_sorted_fgmres_classical_d2 banded3000_fgmres_classical_d2_trunc 2>&1 | tail -12
