#!/bin/bash
# round-2 GPU call 1: regression of the validated suite, the opt-in (never run on a device) suites with full logs, reference goldens for the r2 cases
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2/gpu.txt
PT="python -m pytest -q -m gpu --timeout=300"
echo "== goldens"
timeout 900 python tests/golden/make_golden.py r2 > gpurun_out/r2/make_golden_r2.log 2>&1; tail -5 gpurun_out/r2/make_golden_r2.log
echo "== regress"
timeout 900 $PT tests -x > gpurun_out/r2/regress_suite.log 2>&1; tail -5 gpurun_out/r2/regress_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2/smoke.log 2>&1; tail -2 gpurun_out/r2/smoke.log
echo "== unvalidated"
AMGXB_RUN_UNVALIDATED=1 timeout 1500 $PT tests -rf > gpurun_out/r2/unvalidated.log 2>&1; tail -60 gpurun_out/r2/unvalidated.log
echo "== bench"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2/bench0.json 2> gpurun_out/r2/bench0.err; tail -c 1500 gpurun_out/r2/bench0.json
