#!/bin/bash
# round-2 GPU call 21 (1 GPU): the sliding-window kernel for banded irregular matrices: parity, then the banded workload with / without
mkdir -p gpurun_out/r2/win gpurun_out/r2/ncu
F=gpurun_out/r2/win
PT="python -m pytest -q -m gpu --timeout=300"
AMGXB_WINDOW_VERBOSE=1 timeout 600 $PT tests/test_gpu_window.py -x -s 2>&1 | grep -v "^$" | tail -12 | cut -c1-300
timeout 600 $PT tests/test_gpu_parity.py tests/test_gpu_resetup.py -x 2>&1 | tail -2 | cut -c1-300
echo "== banded workload"
for W in 1; do
  AMGXB_WINDOW=$W AMGXB_WINDOW_VERBOSE=1 timeout 600 python bench.py --workload banded --steps 3 --warmup 2 --no-cpu-baseline > $F/banded_win$W.json 2> $F/banded_win$W.err
  grep "window level" $F/banded_win$W.err | head -4 | cut -c1-250
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$F/banded_win$W.json") if l.startswith('{')][-1]); r=d.get('roofline') or {}
    print('banded window=$W it/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'iters', d['config'].get('iterations_per_step'), d['config'].get('solve_status'), '| jacobi ms', r.get('ms_per_launch'), 'frac', r.get('frac'), '| spmv ms', (r.get('spmv') or {}).get('ms_per_launch'), 'frac', (r.get('spmv') or {}).get('frac'), '| refgpu', (d.get('reference_gpu') or {}).get('value'))
    print('   ', r.get('kernel'))
except Exception as e: print('failed', e); print(open("$F/banded_win$W.err").read()[-1500:])
PY
done
echo "== ncu window kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"csr_window_kernel" -s 4 -c 2 -f -o gpurun_out/r2/ncu/window_banded python bench.py --workload banded --steps 1 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2/ncu/window_banded.log 2>&1
ls -la gpurun_out/r2/ncu | grep window
