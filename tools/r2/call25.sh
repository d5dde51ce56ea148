#!/bin/bash
# round-2 GPU call 25 (1 GPU): window kernel variants on the banded workload (ring 16384 x 2 stages / 8192 x 3 stages; 8- / 12-entry steps)
mkdir -p gpurun_out/r2/win
F=gpurun_out/r2/win
timeout 600 python -m pytest -q -m gpu --timeout=300 tests/test_gpu_window.py -x 2>&1 | tail -2 | cut -c1-300
AMGXB_WINDOW_RING=8192 timeout 600 python -m pytest -q -m gpu --timeout=300 tests/test_gpu_window.py -x 2>&1 | tail -2 | cut -c1-300
run() { # tag env...
  TAG=$1; shift
  env "$@" timeout 600 python bench.py --workload banded --steps 3 --warmup 2 --no-cpu-baseline --no-reference-gpu > $F/banded_$TAG.json 2> $F/banded_$TAG.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$F/banded_$TAG.json") if l.startswith('{')][-1]); r=d.get('roofline') or {}
    print('$TAG it/s', round(d['value'],1), 'iters', d['config'].get('iterations_per_step'), '| jacobi ms', round(r.get('ms_per_launch'),4), 'frac', round(r.get('frac'),3), '| spmv ms', round((r.get('spmv') or {}).get('ms_per_launch'),4), 'frac', round((r.get('spmv') or {}).get('frac'),3))
except Exception as e: print('$TAG failed', e); print(open("$F/banded_$TAG.err").read()[-800:])
PY
}
run wide1_r16 AMGXB_WINDOW_WIDE=1
run wide0_r16 AMGXB_WINDOW_WIDE=0
run wide1_r8 AMGXB_WINDOW_WIDE=1 AMGXB_WINDOW_RING=8192
run wide0_r8 AMGXB_WINDOW_WIDE=0 AMGXB_WINDOW_RING=8192
