#!/bin/bash
# round-2 GPU call 31 (1 GPU): long rows leave the sliced-ELL copy (hub rows of aggregated levels): parity, then the banded workload
mkdir -p gpurun_out/r2/win
F=gpurun_out/r2/win
timeout 500 python -m pytest -q -m gpu --timeout=300 tests/test_gpu_window.py -x 2>&1 | tail -3 | cut -c1-300
AMGXB_WINDOW_VERBOSE=1 timeout 500 python bench.py --workload banded --steps 3 --warmup 2 --no-cpu-baseline --no-reference-gpu > $F/banded_long.json 2> $F/banded_long.err
grep "window level" $F/banded_long.err | cut -c1-330
python - <<PY
import json
try:
    d=json.loads([l for l in open("$F/banded_long.json") if l.startswith('{')][-1]); r=d.get('roofline') or {}
    print('banded it/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'iters', d['config'].get('iterations_per_step'), d['config'].get('solve_status'), '| jacobi ms', round(r.get('ms_per_launch'),4), 'frac', round(r.get('frac'),3), '| spmv ms', round((r.get('spmv') or {}).get('ms_per_launch'),4), 'final rel', d.get('final_relative_residual'))
except Exception as e: print('failed', e); print(open("$F/banded_long.err").read()[-1500:])
PY
AMGXB_WINDOW=0 timeout 500 python bench.py --workload banded --steps 3 --warmup 2 --no-cpu-baseline --no-reference-gpu 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('window off: it/s', round(d['value'],1), 'iters', d['config'].get('iterations_per_step'), 'final rel', d.get('final_relative_residual'))"
