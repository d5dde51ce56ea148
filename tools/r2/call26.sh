#!/bin/bash
# round-2 GPU call 26 (1 GPU): streaming form of the window kernel (per-warp chunk queues): parity, then the banded workload, both forms
mkdir -p gpurun_out/r2/win gpurun_out/r2/ncu
F=gpurun_out/r2/win
timeout 300 python -m pytest -q -m gpu --timeout=120 tests/test_gpu_window.py -x 2>&1 | tail -3 | cut -c1-300
run() { # tag env...
  TAG=$1; shift
  env "$@" timeout 400 python bench.py --workload banded --steps 3 --warmup 2 --no-cpu-baseline --no-reference-gpu > $F/banded_$TAG.json 2> $F/banded_$TAG.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$F/banded_$TAG.json") if l.startswith('{')][-1]); r=d.get('roofline') or {}
    print('$TAG it/s', round(d['value'],1), 'iters', d['config'].get('iterations_per_step'), d['config'].get('solve_status'), '| jacobi ms', round(r.get('ms_per_launch'),4), 'frac', round(r.get('frac'),3), '| spmv ms', round((r.get('spmv') or {}).get('ms_per_launch'),4), 'frac', round((r.get('spmv') or {}).get('frac'),3))
except Exception as e: print('$TAG failed', e); print(open("$F/banded_$TAG.err").read()[-800:])
PY
}
run stream1 AMGXB_WINDOW_STREAM=1
#run stream0 AMGXB_WINDOW_STREAM=0
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"csr_stream_kernel" -s 4 -c 1 -f -o gpurun_out/r2/ncu/stream_banded python bench.py --workload banded --steps 1 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2/ncu/stream_banded.log 2>&1
ls -la gpurun_out/r2/ncu | grep stream
